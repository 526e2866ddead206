// Loss kernels of the deformable-GAN step (models/pose_gan.py:90-108,140-165,173-199 of the reference):
// adversarial log-losses on the PatchGAN logits (fused sigmoid, one launch instead of 2N tiny ones),
// L1, and the nearest-neighbour "perceptual" loss fused with the VGG-19 conv1_1+ReLU feature extractor
// (utils/pose_utils.py:320-338) so the 64-channel features and the reference's [N,64,H,W,25] unfold
// (3.4 GB at 256^2, N=8) never exist in HBM.
#include "common.cuh"

namespace ptk {

// ---------------------------------------------------------------- adversarial
__global__ void adv_loss_kernel(const float* __restrict__ logits, int rows, int J, int n_true, float scale,
                                float* __restrict__ loss, float* __restrict__ dlogits, int ldd) {
  float lt = 0.f, lf = 0.f;
  const int total = rows * J;
  const float invJ = 1.f / (float)J;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / J;
    const float z = logits[i];
    const float p = 1.f / (1.f + expf(-z));
    const float dp = p * (1.f - p);
    float g;
    if (r < n_true) {
      const float q = p + 1e-7f;          // -mean log(out + 1e-7)        pose_gan.py:93-98,148-151
      lt -= logf(q);
      g = -dp / q;
    } else {
      const float q = (1.f - p) + 1e-7f;  // -mean log(1 - out + 1e-7)    pose_gan.py:158-160
      lf -= logf(q);
      g = dp / q;
    }
    if (dlogits) dlogits[(int64_t)i * ldd] = g * scale * invJ;
  }
  lt = warp_sum(lt); lf = warp_sum(lf);
  __shared__ float st[32], sf[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { st[wid] = lt; sf[wid] = lf; }
  __syncthreads();
  if (wid == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    lt = lane < nw ? st[lane] : 0.f; lf = lane < nw ? sf[lane] : 0.f;
    lt = warp_sum(lt); lf = warp_sum(lf);
    if (lane == 0) { atomicAdd(loss, lt * scale * invJ); atomicAdd(loss + 1, lf * scale * invJ); }
  }
}

// ---------------------------------------------------------------- L1 (torch.nn.L1Loss, pose_gan.py:66,105)
__global__ void l1_loss_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, float scale,
                               float* __restrict__ loss, float* __restrict__ grad) {
  float s = 0.f;
  const float gs = scale / (float)n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = a[i] - b[i];
    s += fabsf(d);
    if (grad) grad[i] = d > 0.f ? gs : (d < 0.f ? -gs : 0.f);
  }
  s = warp_sum(s);
  __shared__ float sh[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sh[wid] = s;
  __syncthreads();
  if (wid == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    s = lane < nw ? sh[lane] : 0.f;
    s = warp_sum(s);
    if (lane == 0) atomicAdd(loss, s * gs);
  }
}

// ---------------------------------------------------------------- fused VGG conv1_1 + NN loss
constexpr int kTile = 16;
constexpr int kMaxArea = 5;
constexpr int kFeat = 64;
constexpr int kChunk = 16;
constexpr float kPadValue = -10000.f;  // pose_gan.py:176

// conv1_1 weights [64][3][3][3] and bias [64] live in constant memory for the duration of a launch so that every FFMA
// of the feature extractor takes its weight as a constant operand (no shared-memory load per multiply-add).
__constant__ float c_vgg_w[kFeat * 27];
__constant__ float c_vgg_b[kFeat];

// utils/pose_utils.py:324-331: the NCHW buffer is re-viewed as NHWC, so element with flat per-sample index
// i is normalised with mean[i % 3], std[i % 3].
__device__ __forceinline__ float vgg_pre(float v, int64_t flat) {
  const int r = (int)(flat % 3);
  const float mean = r == 0 ? 0.485f : (r == 1 ? 0.456f : 0.406f);
  const float sd = r == 0 ? 0.229f : (r == 1 ? 0.224f : 0.225f);
  return (v - mean) / sd;
}

// Load a (R x R) x 3 preprocessed patch whose top-left image coordinate is (y0, x0); zero outside the image
// (conv padding=1 pads the PREPROCESSED tensor with zeros).
__device__ __forceinline__ void load_patch(const float* __restrict__ img, int n, int H, int W, int y0, int x0, int R,
                                           float* __restrict__ dst) {
  const int64_t HW = (int64_t)H * W;
  for (int i = threadIdx.x; i < 3 * R * R; i += blockDim.x) {
    const int c = i / (R * R);
    const int r = i - c * R * R;
    const int yy = y0 + r / R, xx = x0 + r % R;
    float v = 0.f;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
      const int64_t flat = c * HW + (int64_t)yy * W + xx;
      v = vgg_pre(__ldg(img + (int64_t)n * 3 * HW + flat), flat);
    }
    dst[i] = v;
  }
}

// relu(conv1_1) for 16 consecutive channels [c0, c0+16) at patch position (py, px) = centre; patch is [3][R][R].
template <int C0>
__device__ __forceinline__ void feat16(const float* __restrict__ patch, int R, int py, int px, float* __restrict__ out) {
  float in[27];
#pragma unroll
  for (int ci = 0; ci < 3; ++ci)
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) in[(ci * 3 + ky) * 3 + kx] = patch[(ci * R + py - 1 + ky) * R + px - 1 + kx];
#pragma unroll
  for (int c = 0; c < kChunk; ++c) {
    float a = c_vgg_b[C0 + c];
#pragma unroll
    for (int t = 0; t < 27; ++t) a = fmaf(in[t], c_vgg_w[(C0 + c) * 27 + t], a);
    out[c] = a > 0.f ? a : 0.f;
  }
}

struct NnCtx {
  const float* s_pin;   // pred input patch   [3][RPI][RPI]
  const float* s_gin;   // target input patch [3][RGI][RGI]
  int RPI, RGI, P, H, W, ty0, tx0, area;
};

// One 16-channel chunk of the forward: GT features over the (16+2P)^2 halo region -> shared memory (all threads),
// then the 16x16 pixel threads accumulate the area^2 shifted L1 distances.
template <int C0>
__device__ __forceinline__ void nn_fwd_chunk(const NnCtx& c, float4* __restrict__ s_gf, int RG, bool live, int ly, int lx,
                                             float* __restrict__ d) {
  for (int pos = threadIdx.x; pos < RG * RG; pos += blockDim.x) {
    const int ry = pos / RG, rx = pos - ry * RG;
    const int iy = c.ty0 - c.P + ry, ix = c.tx0 - c.P + rx;
    float f[kChunk];
    if (iy >= 0 && iy < c.H && ix >= 0 && ix < c.W) {
      feat16<C0>(c.s_gin, c.RGI, ry + 1, rx + 1, f);
    } else {
#pragma unroll
      for (int q = 0; q < kChunk; ++q) f[q] = kPadValue;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) s_gf[q * RG * RG + pos] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
  }
  __syncthreads();
  if (live) {
    float p[kChunk];
    feat16<C0>(c.s_pin, c.RPI, ly + 1, lx + 1, p);
#pragma unroll
    for (int si = 0; si < kMaxArea; ++si) {
      if (si >= c.area) break;
#pragma unroll
      for (int sj = 0; sj < kMaxArea; ++sj) {
        if (sj >= c.area) break;
        const int pos = (ly + si) * RG + lx + sj;
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 g = s_gf[q * RG * RG + pos];
          acc += fabsf(g.x - p[4 * q]) + fabsf(g.y - p[4 * q + 1]) + fabsf(g.z - p[4 * q + 2]) + fabsf(g.w - p[4 * q + 3]);
        }
        d[si * kMaxArea + sj] += acc;
      }
    }
  }
  __syncthreads();
}

constexpr int kNnFwdThreads = 416;   // >= (16 + 4)^2 halo positions in one round; threads 0..255 own the tile pixels
constexpr int kNnBwdThreads = 352;   // >= 18^2 positions
// Both kernels are compiled for TWO resident CTAs per SM (72 / 80 registers instead of 105 / 150, at the price of 0.5 / 1.3 KB
// of spilled accumulators): they are latency-bound on dependent FFMA chains with 11-13 warps per CTA, and the second CTA is
// worth more than the spills cost -- forward 0.501 -> 0.303 ms, backward 0.511 -> 0.388 ms at 8 x 256 x 256, 5 x 5 window
// (tools/bench_nnloss.py).

__global__ void __launch_bounds__(kNnFwdThreads, 2)
nnloss_forward_kernel(const float* __restrict__ pred, const float* __restrict__ target, int H, int W, int area,
                      float scale_over_count, float* __restrict__ loss, uint8_t* __restrict__ argmin) {
  const int P = area / 2;
  const int RG = kTile + 2 * P;        // GT feature region
  const int RGI = RG + 2;              // GT input patch
  const int RPI = kTile + 2;           // pred input patch
  extern __shared__ float smem[];
  float* s_pin = smem;                   // 3*RPI*RPI
  float* s_gin = s_pin + 3 * RPI * RPI;  // 3*RGI*RGI
  float4* s_gf = reinterpret_cast<float4*>(s_gin + ((3 * RGI * RGI + 3) & ~3));  // [4][RG*RG] float4
  const int n = blockIdx.z;
  const int ty0 = blockIdx.y * kTile, tx0 = blockIdx.x * kTile;
  load_patch(pred, n, H, W, ty0 - 1, tx0 - 1, RPI, s_pin);
  load_patch(target, n, H, W, ty0 - P - 1, tx0 - P - 1, RGI, s_gin);
  __syncthreads();
  const int ly = threadIdx.x / kTile, lx = threadIdx.x % kTile;
  const int gy = ty0 + ly, gx = tx0 + lx;
  const bool live = threadIdx.x < kTile * kTile && gy < H && gx < W;
  NnCtx c{s_pin, s_gin, RPI, RGI, P, H, W, ty0, tx0, area};
  float d[kMaxArea * kMaxArea];
#pragma unroll
  for (int s = 0; s < kMaxArea * kMaxArea; ++s) d[s] = 0.f;
  nn_fwd_chunk<0>(c, s_gf, RG, live, ly, lx, d);
  nn_fwd_chunk<16>(c, s_gf, RG, live, ly, lx, d);
  nn_fwd_chunk<32>(c, s_gf, RG, live, ly, lx, d);
  nn_fwd_chunk<48>(c, s_gf, RG, live, ly, lx, d);
  float best = 0.f;
  if (live) {
    best = INFINITY;
    int arg = 0;
    for (int si = 0; si < area; ++si)
      for (int sj = 0; sj < area; ++sj) {
        const float v = d[si * kMaxArea + sj];
        if (v < best) { best = v; arg = si * area + sj; }   // torch.min: first minimum (pose_gan.py:195)
      }
    argmin[((int64_t)n * H + gy) * W + gx] = (uint8_t)arg;
  }
  best = warp_sum(best);
  __shared__ float sh[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sh[wid] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < (int)((blockDim.x + 31) >> 5); ++i) s += sh[i];
    atomicAdd(loss, s * scale_over_count);
  }
}

// One 16-channel chunk of the backward: feature gradients G[c](q) on the 18x18 positions around the tile -> shared
// memory, then every tile pixel gathers d xp[ci](r) = sum_{ky,kx,c} G[c](r - (ky-1,kx-1)) * w[c][ci][ky][kx].
template <int C0>
__device__ __forceinline__ void nn_bwd_chunk(const NnCtx& c, const uint8_t* __restrict__ argmin, int n, float* __restrict__ s_G,
                                             int RQ, float scale_over_count, bool pix, int ly, int lx, float* __restrict__ acc) {
  for (int pos = threadIdx.x; pos < RQ * RQ; pos += blockDim.x) {
    const int ry = pos / RQ, rx = pos - ry * RQ;
    const int qy = c.ty0 - 1 + ry, qx = c.tx0 - 1 + rx;
    float gq[kChunk];
#pragma unroll
    for (int q = 0; q < kChunk; ++q) gq[q] = 0.f;
    if (qy >= 0 && qy < c.H && qx >= 0 && qx < c.W) {
      const int arg = argmin[((int64_t)n * c.H + qy) * c.W + qx];
      const int si = arg / c.area, sj = arg - si * c.area;
      float p[kChunk], t[kChunk];
      feat16<C0>(c.s_pin, c.RPI, ry + 1, rx + 1, p);
      const int gyy = qy + si - c.P, gxx = qx + sj - c.P;
      if (gyy >= 0 && gyy < c.H && gxx >= 0 && gxx < c.W) {
        feat16<C0>(c.s_gin, c.RGI, ry + si + 1, rx + sj + 1, t);
      } else {
#pragma unroll
        for (int q = 0; q < kChunk; ++q) t[q] = kPadValue;
      }
#pragma unroll
      for (int q = 0; q < kChunk; ++q) {
        // d|gt - p|/dp = -sign(gt - p); relu'(feat) = [p > 0]
        const float diff = t[q] - p[q];
        const float sg = diff > 0.f ? -1.f : (diff < 0.f ? 1.f : 0.f);
        gq[q] = p[q] > 0.f ? sg * scale_over_count : 0.f;
      }
    }
#pragma unroll
    for (int q = 0; q < kChunk; ++q) s_G[pos * kChunk + q] = gq[q];
  }
  __syncthreads();
  if (pix) {
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int pos = (ly + 1 - (ky - 1)) * RQ + (lx + 1 - (kx - 1));
        const float4* gp = reinterpret_cast<const float4*>(s_G + pos * kChunk);
#pragma unroll
        for (int q4 = 0; q4 < kChunk / 4; ++q4) {
          const float4 gv = gp[q4];
          const float g[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int ch = C0 + q4 * 4 + j;
            acc[0] = fmaf(g[j], c_vgg_w[ch * 27 + ky * 3 + kx], acc[0]);
            acc[1] = fmaf(g[j], c_vgg_w[ch * 27 + 9 + ky * 3 + kx], acc[1]);
            acc[2] = fmaf(g[j], c_vgg_w[ch * 27 + 18 + ky * 3 + kx], acc[2]);
          }
        }
      }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kNnBwdThreads, 2)
nnloss_backward_kernel(const float* __restrict__ pred, const float* __restrict__ target, const uint8_t* __restrict__ argmin,
                       int H, int W, int area, float scale_over_count, float* __restrict__ dpred) {
  const int P = area / 2;
  const int RQ = kTile + 2;            // positions q whose feature gradient reaches this tile
  const int RPI = RQ + 2;              // pred input patch
  const int RGI = RQ + 2 * P + 2;      // GT input patch
  extern __shared__ float smem[];
  float* s_pin = smem;
  float* s_gin = s_pin + 3 * RPI * RPI;
  float* s_G = s_gin + ((3 * RGI * RGI + 3) & ~3);  // [RQ*RQ][16], 16-byte aligned rows
  const int n = blockIdx.z;
  const int ty0 = blockIdx.y * kTile, tx0 = blockIdx.x * kTile;
  load_patch(pred, n, H, W, ty0 - 2, tx0 - 2, RPI, s_pin);
  load_patch(target, n, H, W, ty0 - 2 - P, tx0 - 2 - P, RGI, s_gin);
  __syncthreads();
  const int ly = threadIdx.x / kTile, lx = threadIdx.x % kTile;
  const bool pix = threadIdx.x < kTile * kTile;
  NnCtx c{s_pin, s_gin, RPI, RGI, P, H, W, ty0, tx0, area};
  float acc[3] = {0.f, 0.f, 0.f};
  nn_bwd_chunk<0>(c, argmin, n, s_G, RQ, scale_over_count, pix, ly, lx, acc);
  nn_bwd_chunk<16>(c, argmin, n, s_G, RQ, scale_over_count, pix, ly, lx, acc);
  nn_bwd_chunk<32>(c, argmin, n, s_G, RQ, scale_over_count, pix, ly, lx, acc);
  nn_bwd_chunk<48>(c, argmin, n, s_G, RQ, scale_over_count, pix, ly, lx, acc);
  const int gy = ty0 + ly, gx = tx0 + lx;
  if (pix && gy < H && gx < W) {
    const int64_t HW = (int64_t)H * W;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
      const int64_t flat = ci * HW + (int64_t)gy * W + gx;
      const int r = (int)(flat % 3);
      const float sd = r == 0 ? 0.229f : (r == 1 ? 0.224f : 0.225f);
      dpred[(int64_t)n * 3 * HW + flat] = acc[ci] / sd;
    }
  }
}

// ---------------------------------------------------------------- NN loss on materialised features
// DeformablePose_GAN.nn_loss(predicted, ground_truth, nh, nw) (models/pose_gan.py:173-199) as a public method: NCHW
// feature tensors of any channel count.  One thread per pixel (coalesced along x), channels streamed.
template <bool BWD>
__global__ void __launch_bounds__(256)
nnloss_feat_kernel(const float* __restrict__ pred, const float* __restrict__ gt, int C, int H, int W, int area,
                   float scale_over_count, float* __restrict__ loss, uint8_t* __restrict__ argmin, float* __restrict__ dpred) {
  const int n = blockIdx.y;
  const int64_t HW = (int64_t)H * W;
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = p < HW;
  const int y = live ? (int)(p / W) : 0, x = live ? (int)(p - (int64_t)y * W) : 0;
  const int P = area / 2;
  const float* pb = pred + (int64_t)n * C * HW + p;
  const float* gb = gt + (int64_t)n * C * HW;
  float best = 0.f;
  if (live && !BWD) {
    float d[kMaxArea * kMaxArea];
#pragma unroll
    for (int s = 0; s < kMaxArea * kMaxArea; ++s) d[s] = 0.f;
    for (int c = 0; c < C; ++c) {
      const float pv = __ldg(pb + c * HW);
#pragma unroll
      for (int si = 0; si < kMaxArea; ++si) {
        if (si >= area) break;
        const int yy = y + si - P;
#pragma unroll
        for (int sj = 0; sj < kMaxArea; ++sj) {
          if (sj >= area) break;
          const int xx = x + sj - P;
          const float g = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(gb + c * HW + (int64_t)yy * W + xx) : kPadValue;
          d[si * kMaxArea + sj] += fabsf(g - pv);
        }
      }
    }
    best = INFINITY;
    int arg = 0;
    for (int si = 0; si < area; ++si)
      for (int sj = 0; sj < area; ++sj) {
        const float v = d[si * kMaxArea + sj];
        if (v < best) { best = v; arg = si * area + sj; }   // torch.min: first minimum (pose_gan.py:195)
      }
    argmin[(int64_t)n * HW + p] = (uint8_t)arg;
  }
  if (live && BWD) {
    const int arg = argmin[(int64_t)n * HW + p];
    const int si = arg / area, sj = arg - si * area;
    const int yy = y + si - P, xx = x + sj - P;
    const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
    for (int c = 0; c < C; ++c) {
      const float pv = __ldg(pb + c * HW);
      const float g = in ? __ldg(gb + c * HW + (int64_t)yy * W + xx) : kPadValue;
      const float diff = g - pv;
      dpred[(int64_t)n * C * HW + c * HW + p] = diff > 0.f ? -scale_over_count : (diff < 0.f ? scale_over_count : 0.f);
    }
  }
  if (!BWD) {
    best = warp_sum(best);
    __shared__ float sh[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sh[wid] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int i = 0; i < 8; ++i) s += sh[i];
      atomicAdd(loss, s * scale_over_count);
    }
  }
}

static int upload_vgg(const float* vgg_w, const float* vgg_b, cudaStream_t st) {
  cudaError_t e = cudaMemcpyToSymbolAsync(c_vgg_w, vgg_w, sizeof(float) * kFeat * 27, 0, cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyToSymbolAsync(c_vgg_b, vgg_b, sizeof(float) * kFeat, 0, cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) return fail(3, "nnloss: cudaMemcpyToSymbolAsync: %s", cudaGetErrorString(e));
  return 0;
}

__global__ void tanh_bwd_combine_kernel(const float* __restrict__ g_nchw, const float* __restrict__ g_nhwc, int ldg,
                                        const float* __restrict__ out_nchw, float* __restrict__ dz, int ld, int C,
                                        int64_t HW, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    // i enumerates NCHW order (coalesced reads of the planes)
    const int64_t p = i % HW;
    const int c = (int)((i / HW) % C);
    const int64_t n = i / (HW * C);
    float g = g_nchw ? g_nchw[i] : 0.f;
    if (g_nhwc) g += g_nhwc[(n * HW + p) * ldg + c];
    const float o = out_nchw[i];
    dz[(n * HW + p) * ld + c] = g * (1.f - o * o);
  }
}

}  // namespace ptk

using namespace ptk;

extern "C" int ptk_adv_loss(const float* logits, int rows, int J, int n_true, float scale, float* loss,
                            float* dlogits, int ldd, void* stream) {
  PTK_REQUIRE(rows > 0 && J > 0, "adv_loss: bad extents");
  int blocks = (rows * J + 255) / 256;
  if (blocks > 64) blocks = 64;
  adv_loss_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(logits, rows, J, n_true, scale, loss, dlogits, ldd);
  PTK_LAUNCH_CHECK("adv_loss_kernel");
  return 0;
}

extern "C" int ptk_l1_loss(const float* a, const float* b, int64_t n, float scale, float* loss, float* grad,
                           void* stream) {
  PTK_REQUIRE(n > 0, "l1_loss: empty input");
  int64_t blocks = (n + 255) / 256;
  if (blocks > (int64_t)num_sms() * 8) blocks = (int64_t)num_sms() * 8;
  l1_loss_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a, b, n, scale, loss, grad);
  PTK_LAUNCH_CHECK("l1_loss_kernel");
  return 0;
}

extern "C" int ptk_nnloss_forward(const float* pred, const float* target, const float* vgg_w, const float* vgg_b,
                                  int N, int H, int W, int area, float scale, float* loss, uint8_t* argmin,
                                  void* stream) {
  PTK_REQUIRE(area >= 1 && area <= kMaxArea && (area & 1), "nnloss: area must be odd and <= %d", kMaxArea);
  PTK_REQUIRE(N > 0 && N <= 65535, "nnloss: bad batch");
  const int P = area / 2, RG = kTile + 2 * P, RGI = RG + 2, RPI = kTile + 2;
  const size_t smem = sizeof(float) * (size_t)(3 * RPI * RPI + ((3 * RGI * RGI + 3) & ~3) + 4 * 4 * RG * RG) + 16;
  int rc = upload_vgg(vgg_w, vgg_b, (cudaStream_t)stream);
  if (rc) return rc;
  cudaFuncSetAttribute(nnloss_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((W + kTile - 1) / kTile, (H + kTile - 1) / kTile, N);
  nnloss_forward_kernel<<<grid, kNnFwdThreads, smem, (cudaStream_t)stream>>>(pred, target, H, W, area,
                                                                            scale / ((float)N * H * W), loss, argmin);
  PTK_LAUNCH_CHECK("nnloss_forward_kernel");
  return 0;
}

extern "C" int ptk_nnloss_backward(const float* pred, const float* target, const float* vgg_w, const float* vgg_b,
                                   const uint8_t* argmin, int N, int H, int W, int area, float scale, float* dpred,
                                   void* stream) {
  PTK_REQUIRE(area >= 1 && area <= kMaxArea && (area & 1), "nnloss: area must be odd and <= %d", kMaxArea);
  PTK_REQUIRE(N > 0 && N <= 65535, "nnloss: bad batch");
  const int P = area / 2, RQ = kTile + 2, RPI = RQ + 2, RGI = RQ + 2 * P + 2;
  const size_t smem = sizeof(float) * (size_t)(3 * RPI * RPI + ((3 * RGI * RGI + 3) & ~3) + RQ * RQ * kChunk) + 16;
  int rc = upload_vgg(vgg_w, vgg_b, (cudaStream_t)stream);
  if (rc) return rc;
  cudaFuncSetAttribute(nnloss_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((W + kTile - 1) / kTile, (H + kTile - 1) / kTile, N);
  nnloss_backward_kernel<<<grid, kNnBwdThreads, smem, (cudaStream_t)stream>>>(pred, target, argmin, H, W, area,
                                                                             scale / ((float)N * H * W), dpred);
  PTK_LAUNCH_CHECK("nnloss_backward_kernel");
  return 0;
}

extern "C" int ptk_nnloss_features_forward(const float* pred, const float* gt, int N, int C, int H, int W, int area,
                                           float scale, float* loss, uint8_t* argmin, void* stream) {
  PTK_REQUIRE(area >= 1 && area <= kMaxArea && (area & 1), "nnloss_features: area must be odd and <= %d", kMaxArea);
  PTK_REQUIRE(N > 0 && N <= 65535 && C > 0 && H > 0 && W > 0, "nnloss_features: bad extents");
  dim3 grid((unsigned)(((int64_t)H * W + 255) / 256), (unsigned)N);
  nnloss_feat_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(pred, gt, C, H, W, area, scale / ((float)N * H * W), loss,
                                                                   argmin, nullptr);
  PTK_LAUNCH_CHECK("nnloss_feat_kernel<fwd>");
  return 0;
}

extern "C" int ptk_nnloss_features_backward(const float* pred, const float* gt, const uint8_t* argmin, int N, int C, int H,
                                            int W, int area, float scale, float* dpred, void* stream) {
  PTK_REQUIRE(area >= 1 && area <= kMaxArea && (area & 1), "nnloss_features: area must be odd and <= %d", kMaxArea);
  PTK_REQUIRE(N > 0 && N <= 65535 && C > 0 && H > 0 && W > 0, "nnloss_features: bad extents");
  dim3 grid((unsigned)(((int64_t)H * W + 255) / 256), (unsigned)N);
  nnloss_feat_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(pred, gt, C, H, W, area, scale / ((float)N * H * W), nullptr,
                                                                  const_cast<uint8_t*>(argmin), dpred);
  PTK_LAUNCH_CHECK("nnloss_feat_kernel<bwd>");
  return 0;
}

extern "C" int ptk_tanh_bwd_combine(const float* g_nchw, const float* g_nhwc, int ldg, const float* out_nchw,
                                    float* dz, int ld, int N, int C, int H, int W, void* stream) {
  const int64_t HW = (int64_t)H * W, total = (int64_t)N * C * HW;
  PTK_REQUIRE(total > 0, "tanh_bwd_combine: empty");
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)num_sms() * 8) blocks = (int64_t)num_sms() * 8;
  tanh_bwd_combine_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(g_nchw, g_nhwc, ldg, out_nchw, dz, ld, C, HW, total);
  PTK_LAUNCH_CHECK("tanh_bwd_combine_kernel");
  return 0;
}
