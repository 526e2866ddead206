// The narrow 3x3 output head of the generator (ReLU -> Conv2d(256 -> 3, k3, p1) -> Tanh, models/networks.py:227-232)
// re-expressed so that its three passes read / write the 256-channel tensor exactly ONCE each on the tensor cores:
//   forward : z27[p][tap*Co + co] = sum_c x[p][c] W[co][c][tap]      (1x1 GEMM, K = 256, N = 27 -> 32)
//             y[p][co] = tanh(b[co] + sum_tap z27[p + off(tap)][tap*Co + co])            (head_shift_add)
//   backward: dzs[q][tap*Co + co] = dz[q - off(tap)][co]                                  (head_shift_gather)
//             dW[co][c][tap] = sum_q x[q][c] dzs[q][tap*Co + co]     (1x1 weight-gradient GEMM over pixels)
//             dx[q][c]       = sum_j dzs[q][j] W[co(j)][c][tap(j)]   (1x1 GEMM, K = 27 -> 32, N = 256)
// (a 3-channel-wide implicit GEMM would re-fetch the 256-channel operand once per tap: 9x the bytes).
// off(tap) = (kh - 1, kw - 1); out-of-image neighbours contribute zero (padding 1).
#include "common.cuh"

namespace ptk {

constexpr int kHeadCols = 32;   // 9 taps x Co (<= 3) columns, zero padded to one 128-byte row

// torch weight [Co][Cin][3][3] -> wk[j][c] (forward B operand, K = c contiguous) and wd[c][j] (dgrad B operand)
__global__ void head_pack_weights_kernel(const float* __restrict__ w, int Co, int Cin, float* __restrict__ wk,
                                         float* __restrict__ wd) {
  const int total = kHeadCols * Cin;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int j = i / Cin, c = i - j * Cin;
    float v = 0.f;
    if (j < 9 * Co) {
      const int tap = j / Co, co = j - tap * Co;
      v = __ldg(w + ((int64_t)co * Cin + c) * 9 + tap);
    }
    wk[(int64_t)j * Cin + c] = v;
    wd[(int64_t)c * kHeadCols + j] = v;
  }
}

// y[p][co] = act(bias[co] + sum_tap z[p + off(tap)][tap*Co + co]); written NCHW and/or into an NHWC slice
__global__ void __launch_bounds__(256)
head_shift_add_kernel(const float* __restrict__ z, const float* __restrict__ bias, int Co, int H, int W, int act,
                      float* __restrict__ y_nchw, float* __restrict__ y_nhwc, int ldy) {
  const int n = blockIdx.z;
  const int ox = blockIdx.x * 32 + (threadIdx.x & 31), oy = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (ox >= W || oy >= H) return;
  const float* zb = z + (int64_t)n * H * W * kHeadCols;
  float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const int iy = oy + kh - 1;
    if (iy < 0 || iy >= H) continue;
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int ix = ox + kw - 1;
      if (ix < 0 || ix >= W) continue;
      const float* zp = zb + ((int64_t)iy * W + ix) * kHeadCols + (kh * 3 + kw) * Co;
      for (int co = 0; co < Co; ++co) acc[co] += __ldg(zp + co);
    }
  }
  for (int co = 0; co < Co; ++co) {
    const float v = apply_act(acc[co] + (bias ? __ldg(bias + co) : 0.f), act);
    if (y_nchw) y_nchw[(((int64_t)n * Co + co) * H + oy) * W + ox] = v;
    if (y_nhwc) y_nhwc[(((int64_t)n * H + oy) * W + ox) * ldy + co] = v;
  }
}

// dzs[q][tap*Co + co] = dz[q - off(tap)][co]   (dz: [N,H,W,ldz] with the Co gradients in the first channels)
__global__ void __launch_bounds__(256)
head_shift_gather_kernel(const float* __restrict__ dz, int ldz, int Co, int H, int W, float* __restrict__ dzs) {
  const int n = blockIdx.z;
  const int qx = blockIdx.x * 32 + (threadIdx.x & 31), qy = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (qx >= W || qy >= H) return;
  const float* db = dz + (int64_t)n * H * W * ldz;
  float row[kHeadCols];
#pragma unroll
  for (int j = 0; j < kHeadCols; ++j) row[j] = 0.f;
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const int py = qy - (kh - 1);
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int px = qx - (kw - 1);
      if (py < 0 || py >= H || px < 0 || px >= W) continue;
      const float* dp = db + ((int64_t)py * W + px) * ldz;
#pragma unroll
      for (int co = 0; co < 3; ++co)
        if (co < Co) row[(kh * 3 + kw) * Co + co] = __ldg(dp + co);
    }
  }
  float4* out = reinterpret_cast<float4*>(dzs + (((int64_t)n * H + qy) * W + qx) * kHeadCols);
#pragma unroll
  for (int j = 0; j < kHeadCols / 4; ++j) out[j] = make_float4(row[4 * j], row[4 * j + 1], row[4 * j + 2], row[4 * j + 3]);
}

// grad[co][c][tap] (+)= sum_parts dwT[part][c][tap*Co + co]
__global__ void head_wgrad_scatter_kernel(const float* __restrict__ dwT, int nparts, int64_t part_stride, int Co, int Cin,
                                          float* __restrict__ grad, int accumulate) {
  const int total = Co * Cin * 9;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int tap = i % 9, c = (i / 9) % Cin, co = i / (9 * Cin);
    const float* src = dwT + (int64_t)c * kHeadCols + tap * Co + co;
    float v = 0.f;
    for (int q = 0; q < nparts; ++q) v += __ldg(src + q * part_stride);
    grad[i] = accumulate ? grad[i] + v : v;
  }
}

}  // namespace ptk

using namespace ptk;

extern "C" int ptk_head_pack_weights(const float* w, int Co, int Cin, float* wk, float* wd, void* stream) {
  PTK_REQUIRE(w && wk && wd && Co >= 1 && Co <= 3 && Cin > 0, "head_pack_weights: Cout must be 1..3");
  head_pack_weights_kernel<<<(kHeadCols * Cin + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, Co, Cin, wk, wd);
  PTK_LAUNCH_CHECK("head_pack_weights_kernel");
  return 0;
}

extern "C" int ptk_head_shift_add(const float* z, const float* bias, int N, int Co, int H, int W, int act, float* y_nchw,
                                  float* y_nhwc, int ldy, void* stream) {
  PTK_REQUIRE(z && (y_nchw || y_nhwc) && Co >= 1 && Co <= 3 && N > 0 && N <= 65535, "head_shift_add: bad arguments");
  dim3 grid((W + 31) / 32, (H + 7) / 8, N);
  head_shift_add_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(z, bias, Co, H, W, act, y_nchw, y_nhwc, ldy);
  PTK_LAUNCH_CHECK("head_shift_add_kernel");
  return 0;
}

extern "C" int ptk_head_shift_gather(const float* dz, int ldz, int N, int Co, int H, int W, float* dzs, void* stream) {
  PTK_REQUIRE(dz && dzs && Co >= 1 && Co <= 3 && ldz >= Co && N > 0 && N <= 65535, "head_shift_gather: bad arguments");
  PTK_REQUIRE((reinterpret_cast<uintptr_t>(dzs) & 15) == 0, "head_shift_gather: dzs must be 16-byte aligned");
  dim3 grid((W + 31) / 32, (H + 7) / 8, N);
  head_shift_gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dz, ldz, Co, H, W, dzs);
  PTK_LAUNCH_CHECK("head_shift_gather_kernel");
  return 0;
}

extern "C" int ptk_head_wgrad_scatter(const float* dwT, int nparts, int64_t part_stride, int Co, int Cin, float* grad,
                                      int accumulate, void* stream) {
  PTK_REQUIRE(dwT && grad && nparts >= 1 && Co >= 1 && Co <= 3 && Cin > 0, "head_wgrad_scatter: bad arguments");
  head_wgrad_scatter_kernel<<<(Co * Cin * 9 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(dwT, nparts, part_stride, Co, Cin, grad, accumulate);
  PTK_LAUNCH_CHECK("head_wgrad_scatter_kernel");
  return 0;
}
