// Layout kernels: NCHW <-> NHWC channel-slice copies (the reference's torch.cat / get_imgpose call sites,
// utils/pose_utils.py:227-233, models/networks.py:271, models/pose_gan.py:86,133-136) and weight repacking
// between the torch checkpoint layout and the GEMM layouts.
#include "common.cuh"

namespace ptk {

// src NCHW plane-major -> dst NHWC slice.  A CTA moves 256 consecutive pixels x up to 32 channels: 1 KB coalesced reads
// per plane into shared memory, then one pixel-major pass whose consecutive threads write consecutive channels.
constexpr int kNhwcPix = 256, kNhwcCh = 32;
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float* __restrict__ src, int C_src, int c_src0, float* __restrict__ dst,
                                    int ld_dst, int c_dst0, int C, int64_t HW, int act) {
  __shared__ float tile[kNhwcCh][kNhwcPix + 1];
  const int n = blockIdx.z;
  const int64_t p0 = (int64_t)blockIdx.x * kNhwcPix;
  const int c0 = blockIdx.y * kNhwcCh;
  const int nc = min(kNhwcCh, C - c0);
  const int np = (int)min((int64_t)kNhwcPix, HW - p0);
  const float* sb = src + ((int64_t)n * C_src + c_src0 + c0) * HW + p0;
  for (int c = 0; c < nc; ++c)
    if ((int)threadIdx.x < np) tile[c][threadIdx.x] = __ldg(sb + (int64_t)c * HW + threadIdx.x);
  __syncthreads();
  float* db = dst + ((int64_t)n * HW + p0) * ld_dst + c_dst0 + c0;
  for (int e = threadIdx.x; e < np * nc; e += 256) {
    const int p = e / nc, c = e - p * nc;
    db[(int64_t)p * ld_dst + c] = apply_act(tile[c][p], act);
  }
}

// Assembles NHWC rows from up to four NCHW channel ranges in ONE pass and writes the WHOLE range [c_dst0, c_dst0 + c_total) of
// every pixel (channels not covered by a segment are written as zeros): 16-byte stores of full 32-byte sectors.  The
// slice-by-slice nchw_to_nhwc above leaves partially written sectors (21 of 32 channels, 3 of 64, ...), each of which costs an
// HBM read-modify-write: ncu showed 556 MB of DRAM reads per training step for 11 such launches (2 TB/s effective).
constexpr int kGatherPix = 128;
struct GatherSegs { const float* src[4]; int Cs[4], c0[4], C[4], cd[4]; int n; };
__global__ void __launch_bounds__(256)
gather_nhwc_kernel(const __grid_constant__ GatherSegs g, float* __restrict__ dst, int ld_dst, int c_dst0, int c_total, int64_t HW) {
  extern __shared__ float s_tile[];                // [c_total][kGatherPix + 1]
  pdl_trigger();
  const int n = blockIdx.y;
  const int64_t p0 = (int64_t)blockIdx.x * kGatherPix;
  const int np = (int)min((int64_t)kGatherPix, HW - p0);
  for (int e = threadIdx.x; e < c_total * (kGatherPix + 1); e += 256) s_tile[e] = 0.f;
  __syncthreads();
  if ((HW & 3) == 0 && np == kGatherPix) {
    // a warp fetches one channel row of the tile per instruction (128 pixels = 32 x 16 bytes), eight rows in flight per CTA
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int q = 0; q < g.n; ++q) {
      const float* sb = g.src[q] + ((int64_t)n * g.Cs[q] + g.c0[q]) * HW + p0 + 4 * lane;
#pragma unroll 4
      for (int c = w; c < g.C[q]; c += 8) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(sb + (int64_t)c * HW));
        float* t = s_tile + (g.cd[q] + c) * (kGatherPix + 1) + 4 * lane;
        t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
      }
    }
  } else {
    const int px = threadIdx.x & (kGatherPix - 1), half = threadIdx.x >> 7;
    for (int q = 0; q < g.n; ++q) {
      const float* sb = g.src[q] + ((int64_t)n * g.Cs[q] + g.c0[q]) * HW + p0;
      if (px < np)
        for (int c = half; c < g.C[q]; c += 2) s_tile[(g.cd[q] + c) * (kGatherPix + 1) + px] = __ldg(sb + (int64_t)c * HW + px);
    }
  }
  __syncthreads();
  const int c4 = c_total >> 2;
  float* db = dst + ((int64_t)n * HW + p0) * ld_dst + c_dst0;
  for (int e = threadIdx.x; e < np * c4; e += 256) {
    const int p = e / c4, q = e - p * c4;
    const float* t = s_tile + (4 * q) * (kGatherPix + 1) + p;
    *reinterpret_cast<float4*>(db + (int64_t)p * ld_dst + 4 * q) =
        make_float4(t[0], t[kGatherPix + 1], t[2 * (kGatherPix + 1)], t[3 * (kGatherPix + 1)]);
  }
}

__global__ void nhwc_to_nchw_kernel(const float* __restrict__ src, int ld_src, int c_src0, float* __restrict__ dst,
                                    int C, int64_t HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int i = ty; i < 32; i += 8) {
    const int64_t p = p0 + i;
    const int c = c0 + tx;
    tile[i][tx] = (c < C && p < HW) ? src[((int64_t)n * HW + p) * ld_src + c_src0 + c] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i;
    const int64_t p = p0 + tx;
    if (c < C && p < HW) dst[((int64_t)n * C + c) * HW + p] = tile[tx][i];
  }
}

// src [A][B][taps] -> dst[tap][a][b_pad] (transpose=0) or dst[tap][b][a_pad] (transpose=1); zero padding.
__global__ void pack_weight_kernel(const float* __restrict__ src, float* __restrict__ dst, int A, int B, int taps,
                                   int A_pad, int B_pad, int transpose) {
  const int64_t rows = transpose ? B_pad : A_pad, cols = transpose ? A_pad : B_pad;
  const int64_t total = (int64_t)taps * rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = i % cols;
    const int64_t r = (i / cols) % rows;
    const int t = (int)(i / (cols * rows));
    const int64_t a = transpose ? c : r, b = transpose ? r : c;
    dst[i] = (a < A && b < B) ? src[(a * B + b) * taps + t] : 0.f;
  }
}

// Both GEMM layouts of one weight tensor in one pass over the checkpoint layout src[A][B][taps]:
//   dst0[t][a][b] (row stride cols0)  and  dst1[t][b][a] (row stride cols1).
// A tile of 16 a x 32 b x taps is read with fully coalesced loads, transposed through shared memory and written in
// 128-byte (dst0) / 64-byte (dst1) runs.  Padding rows/columns are never written (buffers are zero-initialised).
__global__ void __launch_bounds__(256)
pack_weight_dual_kernel(const float* __restrict__ src, float* __restrict__ dst0, float* __restrict__ dst1, int A, int B,
                        int taps, int rows0, int cols0, int rows1, int cols1) {
  constexpr int TA = 16, TB = 32;
  __shared__ float tile[TA][TB][17];
  const int a0 = blockIdx.y * TA, b0 = blockIdx.x * TB;
  const int tid = threadIdx.x;
  const int nb = min(TB, B - b0);
  // load: for each a the (nb * taps) floats starting at src[(a*B + b0)*taps] are contiguous
  for (int a = 0; a < TA; ++a) {
    if (a0 + a >= A) break;
    const float* row = src + ((int64_t)(a0 + a) * B + b0) * taps;
    for (int i = tid; i < nb * taps; i += 256) tile[a][i / taps][i % taps] = __ldg(row + i);
  }
  __syncthreads();
  // dst0[t][a][b]: consecutive threads -> consecutive b
  for (int i = tid; i < taps * TA * TB; i += 256) {
    const int b = i % TB, a = (i / TB) % TA, t = i / (TB * TA);
    if (a0 + a < A && b < nb) dst0[((int64_t)t * rows0 + a0 + a) * cols0 + b0 + b] = tile[a][b][t];
  }
  // dst1[t][b][a]: consecutive threads -> consecutive a
  for (int i = tid; i < taps * TA * TB; i += 256) {
    const int a = i % TA, b = (i / TA) % TB, t = i / (TB * TA);
    if (a0 + a < A && b < nb) dst1[((int64_t)t * rows1 + b0 + b) * cols1 + a0 + a] = tile[a][b][t];
  }
}

__global__ void unpack_weight_grad_kernel(const float* __restrict__ src, float* __restrict__ grad, int A, int B,
                                          int taps, int B_pad, int accumulate) {
  const int64_t total = (int64_t)A * B * taps;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i % taps);
    const int64_t b = (i / taps) % B;
    const int64_t a = i / ((int64_t)taps * B);
    const float v = src[((int64_t)t * A + a) * B_pad + b];
    grad[i] = accumulate ? grad[i] + v : v;
  }
}

// tiled inverse of pack_weight_dual's dst0: grad[a][b][t] (+)= src[t][a][b_pad]; coalesced on both sides
// (nparts partial buffers part_stride floats apart are summed in a fixed order: the split-K reduction of the tensor-core
// wgrad, deterministic)
__global__ void __launch_bounds__(256)
unpack_weight_grad_tiled_kernel(const float* __restrict__ src, int nparts, int64_t part_stride, float* __restrict__ grad,
                                int A, int B, int taps, int B_pad, int accumulate) {
  constexpr int TA = 16, TB = 32;
  __shared__ float tile[TA][TB][17];
  const int a0 = blockIdx.y * TA, b0 = blockIdx.x * TB;
  const int tid = threadIdx.x;
  const int nb = min(TB, B - b0);
  for (int i = tid; i < taps * TA * TB; i += 256) {
    const int b = i % TB, a = (i / TB) % TA, t = i / (TB * TA);
    if (a0 + a < A && b < nb) {
      const float* sp = src + ((int64_t)t * A + a0 + a) * B_pad + b0 + b;
      float v = __ldg(sp);
      for (int q = 1; q < nparts; ++q) v += __ldg(sp + q * part_stride);
      tile[a][b][t] = v;
    }
  }
  __syncthreads();
  for (int a = 0; a < TA; ++a) {
    if (a0 + a >= A) break;
    float* row = grad + ((int64_t)(a0 + a) * B + b0) * taps;
    for (int i = tid; i < nb * taps; i += 256) {
      const float v = tile[a][i / taps][i % taps];
      row[i] = accumulate ? row[i] + v : v;
    }
  }
}

// src [taps][A][Bp] -> dst [taps][Bp][A]; A and Bp are multiples of 32 (GEMM-layout master weights: the other operand
// layout is a plain per-tap matrix transpose).  32 x 32 tiles through shared memory, both sides fully coalesced.
__global__ void __launch_bounds__(256)
transpose_weight_kernel(const float* __restrict__ src, float* __restrict__ dst, int A, int Bp) {
  __shared__ float tile[32][33];
  const int t = blockIdx.z, b0 = blockIdx.x * 32, a0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* s = src + (int64_t)t * A * Bp;
  float* d = dst + (int64_t)t * A * Bp;
#pragma unroll
  for (int i = 0; i < 4; ++i) tile[ty + 8 * i][tx] = __ldg(s + (int64_t)(a0 + ty + 8 * i) * Bp + b0 + tx);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) d[(int64_t)(b0 + ty + 8 * i) * A + a0 + tx] = tile[tx][ty + 8 * i];
}

// dst[i] (+)= sum_q src[q * stride + i]   (fixed order: the deterministic split-K reduction)
__global__ void __launch_bounds__(256)
sum_parts_kernel(const float4* __restrict__ src, int nparts, int64_t stride4, float4* __restrict__ dst, int64_t n4, int accumulate) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = __ldg(src + i);
    for (int q = 1; q < nparts; ++q) {
      const float4 u = __ldg(src + q * stride4 + i);
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    if (accumulate) { const float4 o = dst[i]; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
    dst[i] = v;
  }
}

__global__ void fill_kernel(float* __restrict__ dst, int64_t n, float v) {
  pdl_trigger();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = v;
}

static inline int grid_for(int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace ptk

using namespace ptk;

extern "C" int ptk_nchw_to_nhwc(const float* src, int C_src, int c_src0, float* dst, int ld_dst, int c_dst0,
                                int N, int C, int H, int W, int act, void* stream) {
  PTK_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && N <= 65535, "nchw_to_nhwc: bad extents");
  const int64_t HW = (int64_t)H * W;
  dim3 grid((unsigned)((HW + kNhwcPix - 1) / kNhwcPix), (C + kNhwcCh - 1) / kNhwcCh, N);
  nchw_to_nhwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, C_src, c_src0, dst, ld_dst, c_dst0, C, HW, act);
  PTK_LAUNCH_CHECK("nchw_to_nhwc_kernel");
  return 0;
}

extern "C" int ptk_gather_nhwc(const float* const* srcs, const int* src_channels, const int* c_src0, const int* C, const int* c_dst,
                               int nseg, float* dst, int ld_dst, int c_dst0, int c_total, int N, int H, int W, void* stream) {
  PTK_REQUIRE(srcs && dst && nseg >= 1 && nseg <= 4 && N > 0 && N <= 65535 && H > 0 && W > 0, "gather_nhwc: 1..4 segments, N in [1,65535]");
  PTK_REQUIRE(c_total > 0 && c_total % 4 == 0 && c_total <= 96 && c_dst0 % 4 == 0 && ld_dst % 4 == 0 && c_dst0 + c_total <= ld_dst &&
              (reinterpret_cast<uintptr_t>(dst) & 15) == 0, "gather_nhwc: the written channel range must be float4-aligned and inside the row");
  ptk::GatherSegs g;
  memset(&g, 0, sizeof(g));
  g.n = nseg;
  for (int q = 0; q < nseg; ++q) {
    PTK_REQUIRE(srcs[q] && C[q] > 0 && c_src0[q] >= 0 && c_src0[q] + C[q] <= src_channels[q] && c_dst[q] >= 0 && c_dst[q] + C[q] <= c_total,
                "gather_nhwc: segment %d outside its source or the written range", q);
    g.src[q] = srcs[q]; g.Cs[q] = src_channels[q]; g.c0[q] = c_src0[q]; g.C[q] = C[q]; g.cd[q] = c_dst[q];
  }
  const int64_t HW = (int64_t)H * W;
  const size_t smem = (size_t)c_total * (ptk::kGatherPix + 1) * sizeof(float);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(ptk::gather_nhwc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * (ptk::kGatherPix + 1) * 4); attr = true; }
  ptk::gather_nhwc_kernel<<<dim3((unsigned)((HW + ptk::kGatherPix - 1) / ptk::kGatherPix), (unsigned)N), 256, smem, (cudaStream_t)stream>>>(
      g, dst, ld_dst, c_dst0, c_total, HW);
  PTK_LAUNCH_CHECK("gather_nhwc_kernel");
  return 0;
}

extern "C" int ptk_nhwc_to_nchw(const float* src, int ld_src, int c_src0, float* dst, int N, int C, int H, int W,
                                void* stream) {
  PTK_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && N <= 65535, "nhwc_to_nchw: bad extents");
  const int64_t HW = (int64_t)H * W;
  dim3 grid((unsigned)((HW + 31) / 32), (C + 31) / 32, N), block(32, 8);
  nhwc_to_nchw_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(src, ld_src, c_src0, dst, C, HW);
  PTK_LAUNCH_CHECK("nhwc_to_nchw_kernel");
  return 0;
}

extern "C" int ptk_pack_weight(const float* src, float* dst, int A, int B, int taps, int A_pad, int B_pad,
                               int transpose, void* stream) {
  PTK_REQUIRE(A_pad >= A && B_pad >= B && taps > 0, "pack_weight: bad extents");
  const int64_t total = (int64_t)taps * A_pad * B_pad;
  pack_weight_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, A, B, taps, A_pad, B_pad, transpose);
  PTK_LAUNCH_CHECK("pack_weight_kernel");
  return 0;
}

extern "C" int ptk_pack_weight_dual(const float* src, float* dst0, float* dst1, int A, int B, int taps, int rows0,
                                    int cols0, int rows1, int cols1, void* stream) {
  PTK_REQUIRE(rows0 >= A && cols0 >= B && rows1 >= B && cols1 >= A && taps > 0 && taps <= 16, "pack_weight_dual: bad extents");
  dim3 grid((B + 31) / 32, (A + 15) / 16);
  PTK_REQUIRE(grid.y <= 65535, "pack_weight_dual: A too large");
  pack_weight_dual_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst0, dst1, A, B, taps, rows0, cols0, rows1, cols1);
  PTK_LAUNCH_CHECK("pack_weight_dual_kernel");
  return 0;
}

extern "C" int ptk_unpack_weight_grad_parts(const float* src, int nparts, int64_t part_stride, float* grad, int A, int B,
                                            int taps, int B_pad, int accumulate, void* stream) {
  PTK_REQUIRE(nparts >= 1 && taps > 0 && taps <= 16 && (A + 15) / 16 <= 65535, "unpack_weight_grad_parts: bad extents");
  dim3 grid((B + 31) / 32, (A + 15) / 16);
  unpack_weight_grad_tiled_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, nparts, part_stride, grad, A, B, taps, B_pad, accumulate);
  PTK_LAUNCH_CHECK("unpack_weight_grad_tiled_kernel");
  return 0;
}

extern "C" int ptk_unpack_weight_grad(const float* src, float* grad, int A, int B, int taps, int B_pad,
                                      int accumulate, void* stream) {
  const int64_t total = (int64_t)A * B * taps;
  if (taps <= 16 && (A + 15) / 16 <= 65535) {
    dim3 grid((B + 31) / 32, (A + 15) / 16);
    unpack_weight_grad_tiled_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, 1, 0, grad, A, B, taps, B_pad, accumulate);
  } else {
    unpack_weight_grad_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(src, grad, A, B, taps, B_pad, accumulate);
  }
  PTK_LAUNCH_CHECK("unpack_weight_grad_kernel");
  return 0;
}

extern "C" int ptk_transpose_weight(const float* src, float* dst, int taps, int A, int Bp, void* stream) {
  PTK_REQUIRE(src && dst && taps > 0 && A > 0 && Bp > 0 && A % 32 == 0 && Bp % 32 == 0 && A / 32 <= 65535 && taps <= 65535,
              "transpose_weight: A and Bp must be multiples of 32");
  dim3 grid(Bp / 32, A / 32, taps);
  transpose_weight_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, A, Bp);
  PTK_LAUNCH_CHECK("transpose_weight_kernel");
  return 0;
}

extern "C" int ptk_sum_parts(const float* src, int nparts, int64_t part_stride, float* dst, int64_t n, int accumulate,
                             void* stream) {
  PTK_REQUIRE(src && dst && nparts >= 1 && n > 0 && n % 4 == 0 && part_stride % 4 == 0, "sum_parts: n and part_stride must be multiples of 4");
  PTK_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0, "sum_parts: pointers must be 16-byte aligned");
  sum_parts_kernel<<<grid_for(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(src), nparts, part_stride / 4,
                                                                         reinterpret_cast<float4*>(dst), n / 4, accumulate);
  PTK_LAUNCH_CHECK("sum_parts_kernel");
  return 0;
}

extern "C" int ptk_fill(float* dst, int64_t n, float value, void* stream) {
  if (n <= 0) return 0;
  fill_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(dst, n, value);
  PTK_LAUNCH_CHECK("fill_kernel");
  return 0;
}

// ---------------------------------------------------------------- VGG prefix helpers (content_loss_layer deeper than block1_conv2)
namespace ptk {

// utils/pose_utils.py:324-331: the NCHW buffer is re-viewed as NHWC without a permute, so the element with flat per-sample
// index i is normalised with mean[i % 3], std[i % 3].  Output: NHWC with `ld` channels (3 used, the rest left untouched).
__global__ void vgg_preprocess_kernel(const float* __restrict__ x, float* __restrict__ out, int ld, int64_t HW, int64_t total, int backward) {
  pdl_trigger();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / (3 * HW), flat = i - n * 3 * HW;      // NCHW order: flat = c*HW + p
    const int c = (int)(flat / HW);
    const int64_t p = flat - (int64_t)c * HW;
    const int r = (int)(flat % 3);
    const float mean = r == 0 ? 0.485f : (r == 1 ? 0.456f : 0.406f);
    const float sd = r == 0 ? 0.229f : (r == 1 ? 0.224f : 0.225f);
    if (!backward) out[(n * HW + p) * ld + c] = (x[i] - mean) / sd;
    else out[i] = x[(n * HW + p) * ld + c] / sd;               // x = gradient w.r.t. the NHWC preprocessed tensor, out = NCHW
  }
}

// nn.MaxPool2d(2, 2) on NHWC (floor mode).  Backward routes the gradient to the first maximum of the window (row-major).
__global__ void maxpool2_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y, int ldy, int H, int W, int C, int64_t total) {
  pdl_trigger();
  const int OH = H / 2, OW = W / 2, C4 = C >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) << 2;
    int64_t t = i / C4;
    const int ox = (int)(t % OW); t /= OW;
    const int oy = (int)(t % OH);
    const int64_t n = t / OH;
    const float* p = x + ((n * H + 2 * oy) * W + 2 * ox) * (int64_t)ldx + c;
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + ldx));
    const float4 d = __ldg(reinterpret_cast<const float4*>(p + (int64_t)W * ldx)), e = __ldg(reinterpret_cast<const float4*>(p + (int64_t)W * ldx + ldx));
    *reinterpret_cast<float4*>(y + ((n * OH + oy) * OW + ox) * (int64_t)ldy + c) =
        make_float4(fmaxf(fmaxf(a.x, b.x), fmaxf(d.x, e.x)), fmaxf(fmaxf(a.y, b.y), fmaxf(d.y, e.y)),
                    fmaxf(fmaxf(a.z, b.z), fmaxf(d.z, e.z)), fmaxf(fmaxf(a.w, b.w), fmaxf(d.w, e.w)));
  }
}

__device__ __forceinline__ void route4(float g, float a, float b, float d, float e, float& ga, float& gb, float& gd, float& ge) {
  const float m = fmaxf(fmaxf(a, b), fmaxf(d, e));
  ga = gb = gd = ge = 0.f;
  if (a == m) ga = g; else if (b == m) gb = g; else if (d == m) gd = g; else ge = g;
}

__global__ void maxpool2_bwd_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ x, int ldx, float* __restrict__ dx,
                                    int lddx, int H, int W, int C, int64_t total) {
  pdl_trigger();
  const int OH = H / 2, OW = W / 2, C4 = C >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) << 2;
    int64_t t = i / C4;
    const int ox = (int)(t % OW); t /= OW;
    const int oy = (int)(t % OH);
    const int64_t n = t / OH;
    const int64_t pix = (n * H + 2 * oy) * W + 2 * ox;
    const float* p = x + pix * ldx + c;
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + ldx));
    const float4 d = __ldg(reinterpret_cast<const float4*>(p + (int64_t)W * ldx)), e = __ldg(reinterpret_cast<const float4*>(p + (int64_t)W * ldx + ldx));
    const float4 g = __ldg(reinterpret_cast<const float4*>(dy + ((n * OH + oy) * OW + ox) * (int64_t)lddy + c));
    float4 ga, gb, gd, ge;
    route4(g.x, a.x, b.x, d.x, e.x, ga.x, gb.x, gd.x, ge.x);
    route4(g.y, a.y, b.y, d.y, e.y, ga.y, gb.y, gd.y, ge.y);
    route4(g.z, a.z, b.z, d.z, e.z, ga.z, gb.z, gd.z, ge.z);
    route4(g.w, a.w, b.w, d.w, e.w, ga.w, gb.w, gd.w, ge.w);
    float* q = dx + pix * lddx + c;
    *reinterpret_cast<float4*>(q) = ga;
    *reinterpret_cast<float4*>(q + lddx) = gb;
    *reinterpret_cast<float4*>(q + (int64_t)W * lddx) = gd;
    *reinterpret_cast<float4*>(q + (int64_t)W * lddx + lddx) = ge;
  }
}

// dy *= (y > 0) in place (nn.ReLU backward from the saved OUTPUT), NHWC float4 lanes
__global__ void relu_bwd_kernel(const float* __restrict__ y, int ldy, float* __restrict__ dy, int lddy, int C, int64_t total) {
  pdl_trigger();
  const int C4 = C >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / C4;
    const int c = (int)(i - p * C4) << 2;
    const float4 v = __ldg(reinterpret_cast<const float4*>(y + p * ldy + c));
    float4* gp = reinterpret_cast<float4*>(dy + p * lddy + c);
    float4 g = *gp;
    g.x = v.x > 0.f ? g.x : 0.f; g.y = v.y > 0.f ? g.y : 0.f; g.z = v.z > 0.f ? g.z : 0.f; g.w = v.w > 0.f ? g.w : 0.f;
    *gp = g;
  }
}

}  // namespace ptk

extern "C" int ptk_relu_backward(const float* y, int ldy, float* dy, int lddy, int64_t pixels, int C, void* stream) {
  PTK_REQUIRE(y && dy && pixels > 0 && C > 0 && C % 4 == 0 && ldy % 4 == 0 && lddy % 4 == 0, "relu_backward: bad arguments");
  const int64_t total = pixels * (C / 4);
  ptk::relu_bwd_kernel<<<ptk::grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(y, ldy, dy, lddy, C, total);
  PTK_LAUNCH_CHECK("relu_bwd_kernel");
  return 0;
}

extern "C" int ptk_vgg_preprocess(const float* x, float* out, int ld, int N, int H, int W, int backward, void* stream) {
  const int64_t HW = (int64_t)H * W, total = (int64_t)N * 3 * HW;
  PTK_REQUIRE(x && out && total > 0 && ld >= 3, "vgg_preprocess: bad arguments");
  ptk::vgg_preprocess_kernel<<<ptk::grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x, out, ld, HW, total, backward);
  PTK_LAUNCH_CHECK("vgg_preprocess_kernel");
  return 0;
}

extern "C" int ptk_maxpool2_forward(const float* x, int ldx, float* y, int ldy, int N, int H, int W, int C, void* stream) {
  PTK_REQUIRE(x && y && N > 0 && H >= 2 && W >= 2 && C > 0 && C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0, "maxpool2: bad arguments");
  const int64_t total = (int64_t)N * (H / 2) * (W / 2) * (C / 4);
  ptk::maxpool2_kernel<<<ptk::grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, y, ldy, H, W, C, total);
  PTK_LAUNCH_CHECK("maxpool2_kernel");
  return 0;
}

extern "C" int ptk_maxpool2_backward(const float* dy, int lddy, const float* x, int ldx, float* dx, int lddx, int N, int H, int W, int C,
                                     void* stream) {
  PTK_REQUIRE(dy && x && dx && N > 0 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0 && C > 0 && C % 4 == 0 && ldx % 4 == 0 &&
              lddy % 4 == 0 && lddx % 4 == 0, "maxpool2_backward: even extents and 4-float strides required");
  const int64_t total = (int64_t)N * (H / 2) * (W / 2) * (C / 4);
  ptk::maxpool2_bwd_kernel<<<ptk::grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(dy, lddy, x, ldx, dx, lddx, H, W, C, total);
  PTK_LAUNCH_CHECK("maxpool2_bwd_kernel");
  return 0;
}
