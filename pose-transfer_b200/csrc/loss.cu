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
__device__ __forceinline__ void feat16(const float* __restrict__ patch, int R, int py, int px,
                                       const float* __restrict__ s_w, const float* __restrict__ s_b, int c0,
                                       float* __restrict__ out) {
  float in[27];
#pragma unroll
  for (int ci = 0; ci < 3; ++ci)
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) in[(ci * 3 + ky) * 3 + kx] = patch[(ci * R + py - 1 + ky) * R + px - 1 + kx];
#pragma unroll
  for (int c = 0; c < kChunk; ++c) {
    float a = s_b[c0 + c];
    const float* wp = s_w + (c0 + c) * 27;
#pragma unroll
    for (int t = 0; t < 27; ++t) a = fmaf(in[t], wp[t], a);
    out[c] = a > 0.f ? a : 0.f;
  }
}

__global__ void __launch_bounds__(256)
nnloss_forward_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                      const float* __restrict__ vgg_w, const float* __restrict__ vgg_b, int H, int W, int area,
                      float scale_over_count, float* __restrict__ loss, uint8_t* __restrict__ argmin) {
  const int P = area / 2;
  const int RG = kTile + 2 * P;        // GT feature region
  const int RGI = RG + 2;              // GT input patch
  const int RPI = kTile + 2;           // pred input patch
  extern __shared__ float smem[];
  float* s_w = smem;                   // 64*27
  float* s_b = s_w + kFeat * 27;       // 64
  float* s_pin = s_b + kFeat;          // 3*RPI*RPI
  float* s_gin = s_pin + 3 * RPI * RPI;  // 3*RGI*RGI
  float4* s_gf = reinterpret_cast<float4*>(s_gin + ((3 * RGI * RGI + 3) & ~3));  // [4][RG*RG] float4
  const int n = blockIdx.z;
  const int ty0 = blockIdx.y * kTile, tx0 = blockIdx.x * kTile;
  for (int i = threadIdx.x; i < kFeat * 27; i += blockDim.x) s_w[i] = vgg_w[i];
  for (int i = threadIdx.x; i < kFeat; i += blockDim.x) s_b[i] = vgg_b[i];
  load_patch(pred, n, H, W, ty0 - 1, tx0 - 1, RPI, s_pin);
  load_patch(target, n, H, W, ty0 - P - 1, tx0 - P - 1, RGI, s_gin);
  __syncthreads();
  const int ly = threadIdx.x / kTile, lx = threadIdx.x % kTile;
  const int gy = ty0 + ly, gx = tx0 + lx;
  const bool live = gy < H && gx < W;
  float d[kMaxArea * kMaxArea];
#pragma unroll
  for (int s = 0; s < kMaxArea * kMaxArea; ++s) d[s] = 0.f;
  for (int c0 = 0; c0 < kFeat; c0 += kChunk) {
    // GT features of this channel chunk over the halo region (value -10000 outside the image)
    for (int pos = threadIdx.x; pos < RG * RG; pos += blockDim.x) {
      const int ry = pos / RG, rx = pos - ry * RG;
      const int iy = ty0 - P + ry, ix = tx0 - P + rx;
      float f[kChunk];
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
        feat16(s_gin, RGI, ry + 1, rx + 1, s_w, s_b, c0, f);
      } else {
#pragma unroll
        for (int c = 0; c < kChunk; ++c) f[c] = kPadValue;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) s_gf[q * RG * RG + pos] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
    }
    __syncthreads();
    if (live) {
      float p[kChunk];
      feat16(s_pin, RPI, ly + 1, lx + 1, s_w, s_b, c0, p);
#pragma unroll
      for (int si = 0; si < kMaxArea; ++si) {
        if (si >= area) break;
#pragma unroll
        for (int sj = 0; sj < kMaxArea; ++sj) {
          if (sj >= area) break;
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

__global__ void __launch_bounds__(256)
nnloss_backward_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                       const float* __restrict__ vgg_w, const float* __restrict__ vgg_b,
                       const uint8_t* __restrict__ argmin, int H, int W, int area, float scale_over_count,
                       float* __restrict__ dpred) {
  const int P = area / 2;
  const int RQ = kTile + 2;            // positions q whose feature gradient reaches this tile
  const int RPI = RQ + 2;              // pred input patch
  const int RGI = RQ + 2 * P + 2;      // GT input patch
  extern __shared__ float smem[];
  float* s_w = smem;
  float* s_b = s_w + kFeat * 27;
  float* s_pin = s_b + kFeat;
  float* s_gin = s_pin + 3 * RPI * RPI;
  float* s_G = s_gin + 3 * RGI * RGI;  // [RQ*RQ][16]
  const int n = blockIdx.z;
  const int ty0 = blockIdx.y * kTile, tx0 = blockIdx.x * kTile;
  for (int i = threadIdx.x; i < kFeat * 27; i += blockDim.x) s_w[i] = vgg_w[i];
  for (int i = threadIdx.x; i < kFeat; i += blockDim.x) s_b[i] = vgg_b[i];
  load_patch(pred, n, H, W, ty0 - 2, tx0 - 2, RPI, s_pin);
  load_patch(target, n, H, W, ty0 - 2 - P, tx0 - 2 - P, RGI, s_gin);
  __syncthreads();
  const int ly = threadIdx.x / kTile, lx = threadIdx.x % kTile;
  float acc[3] = {0.f, 0.f, 0.f};
  for (int c0 = 0; c0 < kFeat; c0 += kChunk) {
    for (int pos = threadIdx.x; pos < RQ * RQ; pos += blockDim.x) {
      const int ry = pos / RQ, rx = pos - ry * RQ;
      const int qy = ty0 - 1 + ry, qx = tx0 - 1 + rx;
      float gq[kChunk];
#pragma unroll
      for (int c = 0; c < kChunk; ++c) gq[c] = 0.f;
      if (qy >= 0 && qy < H && qx >= 0 && qx < W) {
        const int arg = argmin[((int64_t)n * H + qy) * W + qx];
        const int si = arg / area, sj = arg - si * area;
        float p[kChunk], t[kChunk];
        feat16(s_pin, RPI, ry + 1, rx + 1, s_w, s_b, c0, p);
        const int gyy = qy + si - P, gxx = qx + sj - P;
        if (gyy >= 0 && gyy < H && gxx >= 0 && gxx < W) {
          feat16(s_gin, RGI, ry + si + 1, rx + sj + 1, s_w, s_b, c0, t);
        } else {
#pragma unroll
          for (int c = 0; c < kChunk; ++c) t[c] = kPadValue;
        }
#pragma unroll
        for (int c = 0; c < kChunk; ++c) {
          // d|gt - p|/dp = -sign(gt - p); relu'(feat) = [p > 0]
          const float diff = t[c] - p[c];
          const float sg = diff > 0.f ? -1.f : (diff < 0.f ? 1.f : 0.f);
          gq[c] = p[c] > 0.f ? sg * scale_over_count : 0.f;
        }
      }
#pragma unroll
      for (int c = 0; c < kChunk; ++c) s_G[pos * kChunk + c] = gq[c];
    }
    __syncthreads();
    // d xp[ci](r) = sum_{ky,kx,c} G[c](r - (ky-1,kx-1)) * w[c][ci][ky][kx]
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int pos = (ly + 1 - (ky - 1)) * RQ + (lx + 1 - (kx - 1));
        const float* gp = s_G + pos * kChunk;
#pragma unroll
        for (int c = 0; c < kChunk; ++c) {
          const float gv = gp[c];
          const float* wp = s_w + (c0 + c) * 27 + ky * 3 + kx;
          acc[0] = fmaf(gv, wp[0], acc[0]);
          acc[1] = fmaf(gv, wp[9], acc[1]);
          acc[2] = fmaf(gv, wp[18], acc[2]);
        }
      }
    __syncthreads();
  }
  const int gy = ty0 + ly, gx = tx0 + lx;
  if (gy < H && gx < W) {
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
  const size_t smem = sizeof(float) * (size_t)(kFeat * 27 + kFeat + 3 * RPI * RPI + ((3 * RGI * RGI + 3) & ~3) + 4 * 4 * RG * RG) + 16;
  cudaFuncSetAttribute(nnloss_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((W + kTile - 1) / kTile, (H + kTile - 1) / kTile, N);
  nnloss_forward_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(pred, target, vgg_w, vgg_b, H, W, area,
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
  const size_t smem = sizeof(float) * (size_t)(kFeat * 27 + kFeat + 3 * RPI * RPI + 3 * RGI * RGI + RQ * RQ * kChunk);
  cudaFuncSetAttribute(nnloss_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((W + kTile - 1) / kTile, (H + kTile - 1) / kTile, N);
  nnloss_backward_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(pred, target, vgg_w, vgg_b, argmin, H, W, area,
                                                                   scale / ((float)N * H * W), dpred);
  PTK_LAUNCH_CHECK("nnloss_backward_kernel");
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
