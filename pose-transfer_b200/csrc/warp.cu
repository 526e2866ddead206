// Fused per-body-part affine warp of an encoder skip tensor (AffineTransformLayer + AffineLayer,
// utils/pose_transform.py:16-92 of the reference).  The reference materialises K=10 replicas of the
// feature map, an affine grid, the sampled replicas and the masked product before a max over parts, and
// round-trips the masks through the CPU (cv2.resize) once per level.  Here: one pass, coordinates in
// registers, masks from a device-side pyramid, zero-mask parts skipped, output written (optionally through
// the consumer's ReLU) straight into its slice of the decoder's concat buffer.  HBM-bound gather.
#include "common.cuh"

namespace ptk {

struct Theta { float a, b, tx, c, d, ty; };

// AffineTransformLayer.forward :72-76 followed by AffineLayer.normalize_transforms :48-58 (sequential
// in-place updates: tx' sees the rescaled b, ty' the rescaled c; tx is scaled with the HEIGHT ratio).
__device__ __forceinline__ Theta normalized_theta(const float* __restrict__ wp, int h, int w, int H0, int W0) {
  Theta t;
  const float mul_x = (float)H0 / (float)h, mul_y = (float)W0 / (float)w;
  t.a = wp[0];
  t.b = wp[1] * (float)w / (float)h;
  t.tx = (wp[2] / mul_x) * 2.f / (float)h + t.a + t.b - 1.f;
  t.c = wp[3] * (float)h / (float)w;
  t.d = wp[4];
  t.ty = (wp[5] / mul_y) * 2.f / (float)w + t.c + t.d - 1.f;
  return t;
}

struct Footprint { int x0, y0; float w00, w01, w10, w11; bool any; };

// F.affine_grid + grid_sample(bilinear, zeros) coordinates for output pixel (i, j) (pose_transform.py:37-39)
__device__ __forceinline__ Footprint footprint(const Theta& t, int i, int j, int h, int w, int align_corners) {
  float gx, gy;
  if (align_corners) {
    gx = w > 1 ? (float)j * 2.f / (float)(w - 1) - 1.f : 0.f;
    gy = h > 1 ? (float)i * 2.f / (float)(h - 1) - 1.f : 0.f;
  } else {
    gx = (2.f * (float)j + 1.f) / (float)w - 1.f;
    gy = (2.f * (float)i + 1.f) / (float)h - 1.f;
  }
  const float sx = t.a * gx + t.b * gy + t.tx;
  const float sy = t.c * gx + t.d * gy + t.ty;
  float px, py;
  if (align_corners) {
    px = (sx + 1.f) * 0.5f * (float)(w - 1);
    py = (sy + 1.f) * 0.5f * (float)(h - 1);
  } else {
    px = ((sx + 1.f) * (float)w - 1.f) * 0.5f;
    py = ((sy + 1.f) * (float)h - 1.f) * 0.5f;
  }
  Footprint f;
  const float fx0 = floorf(px), fy0 = floorf(py);
  const float fx = px - fx0, fy = py - fy0;
  // keep the integer conversion safe for the 1000-pixel "missing part" translations
  f.x0 = (int)fminf(fmaxf(fx0, -2.f), (float)w);
  f.y0 = (int)fminf(fmaxf(fy0, -2.f), (float)h);
  const bool xin0 = f.x0 >= 0 && f.x0 < w, xin1 = f.x0 + 1 >= 0 && f.x0 + 1 < w;
  const bool yin0 = f.y0 >= 0 && f.y0 < h, yin1 = f.y0 + 1 >= 0 && f.y0 + 1 < h;
  f.w00 = (yin0 && xin0) ? (1.f - fy) * (1.f - fx) : 0.f;
  f.w01 = (yin0 && xin1) ? (1.f - fy) * fx : 0.f;
  f.w10 = (yin1 && xin0) ? fy * (1.f - fx) : 0.f;
  f.w11 = (yin1 && xin1) ? fy * fx : 0.f;
  f.any = (xin0 || xin1) && (yin0 || yin1);
  return f;
}

constexpr int kMaxParts = 16;

// grid (chunks, N); thread -> (pixel, 4 channels)
__global__ void __launch_bounds__(256)
warp_forward_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ warps,
                    const float* __restrict__ mask_lvl, float* __restrict__ y, int ldy, uint8_t* __restrict__ argk,
                    int C, int h, int w, int K, int H0, int W0, int align_corners, int act) {
  __shared__ Theta s_theta[kMaxParts];
  const int n = blockIdx.y;
  if (threadIdx.x < K) s_theta[threadIdx.x] = normalized_theta(warps + ((int64_t)n * K + threadIdx.x) * 8, h, w, H0, W0);
  __syncthreads();
  const int C4 = C >> 2;
  const int64_t HW = (int64_t)h * w;
  const int64_t total = HW * C4;
  const float* xb = x + (int64_t)n * HW * ldx;
  const float* mb = mask_lvl + (int64_t)n * HW * K;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = idx / C4;
    const int c = (int)(idx - p * C4) << 2;
    const int i = (int)(p / w), j = (int)(p - (int64_t)i * w);
    float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    uchar4 arg = make_uchar4(255, 255, 255, 255);
    for (int k = 0; k < K; ++k) {
      const float m = __ldg(mb + p * K + k);
      float4 cand = make_float4(0.f, 0.f, 0.f, 0.f);
      unsigned char tag = 255;
      if (m != 0.f) {
        const Footprint f = footprint(s_theta[k], i, j, h, w, align_corners);
        if (f.any) {
          tag = (unsigned char)k;
          const float* r0 = xb + ((int64_t)f.y0 * w + f.x0) * ldx + c;
          const float* r1 = r0 + (int64_t)w * ldx;
          if (f.w00 != 0.f) { const float4 v = __ldg(reinterpret_cast<const float4*>(r0)); cand.x += f.w00 * v.x; cand.y += f.w00 * v.y; cand.z += f.w00 * v.z; cand.w += f.w00 * v.w; }
          if (f.w01 != 0.f) { const float4 v = __ldg(reinterpret_cast<const float4*>(r0 + ldx)); cand.x += f.w01 * v.x; cand.y += f.w01 * v.y; cand.z += f.w01 * v.z; cand.w += f.w01 * v.w; }
          if (f.w10 != 0.f) { const float4 v = __ldg(reinterpret_cast<const float4*>(r1)); cand.x += f.w10 * v.x; cand.y += f.w10 * v.y; cand.z += f.w10 * v.z; cand.w += f.w10 * v.w; }
          if (f.w11 != 0.f) { const float4 v = __ldg(reinterpret_cast<const float4*>(r1 + ldx)); cand.x += f.w11 * v.x; cand.y += f.w11 * v.y; cand.z += f.w11 * v.z; cand.w += f.w11 * v.w; }
          cand.x *= m; cand.y *= m; cand.z *= m; cand.w *= m;
        }
      }
      // torch.max(dim=1): first maximum wins (pose_transform.py:89)
      if (cand.x > best.x) { best.x = cand.x; arg.x = tag; }
      if (cand.y > best.y) { best.y = cand.y; arg.y = tag; }
      if (cand.z > best.z) { best.z = cand.z; arg.z = tag; }
      if (cand.w > best.w) { best.w = cand.w; arg.w = tag; }
    }
    const float4 o = make_float4(apply_act(best.x, act), apply_act(best.y, act), apply_act(best.z, act), apply_act(best.w, act));
    *reinterpret_cast<float4*>(y + ((int64_t)n * HW + p) * ldy + c) = o;
    *reinterpret_cast<uchar4*>(argk + ((int64_t)n * HW + p) * C + c) = arg;
  }
}

__global__ void __launch_bounds__(256)
warp_backward_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ y, int ldy, int act,
                     const float* __restrict__ warps, const float* __restrict__ mask_lvl,
                     const uint8_t* __restrict__ argk, float* __restrict__ dx, int C, int h, int w, int K, int H0,
                     int W0, int align_corners) {
  __shared__ Theta s_theta[kMaxParts];
  const int n = blockIdx.y;
  if (threadIdx.x < K) s_theta[threadIdx.x] = normalized_theta(warps + ((int64_t)n * K + threadIdx.x) * 8, h, w, H0, W0);
  __syncthreads();
  const int C4 = C >> 2;
  const int64_t HW = (int64_t)h * w;
  const int64_t total = HW * C4;
  const float* mb = mask_lvl + (int64_t)n * HW * K;
  float* dxb = dx + (int64_t)n * HW * C;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = idx / C4;
    const int c = (int)(idx - p * C4) << 2;
    const uchar4 arg4 = *reinterpret_cast<const uchar4*>(argk + ((int64_t)n * HW + p) * C + c);
    if (arg4.x == 255 && arg4.y == 255 && arg4.z == 255 && arg4.w == 255) continue;
    const int i = (int)(p / w), j = (int)(p - (int64_t)i * w);
    float4 g = __ldg(reinterpret_cast<const float4*>(dy + ((int64_t)n * HW + p) * lddy + c));
    if (act != PTK_ACT_NONE) {
      const float4 yv = __ldg(reinterpret_cast<const float4*>(y + ((int64_t)n * HW + p) * ldy + c));
      g.x *= act_grad_from_output(yv.x, act); g.y *= act_grad_from_output(yv.y, act);
      g.z *= act_grad_from_output(yv.z, act); g.w *= act_grad_from_output(yv.w, act);
    }
    const float gs[4] = {g.x, g.y, g.z, g.w};
    const unsigned char as[4] = {arg4.x, arg4.y, arg4.z, arg4.w};
    int prev = -1;
    Footprint f;
    float m = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = as[q];
      if (k == 255 || gs[q] == 0.f) continue;
      if (k != prev) {
        f = footprint(s_theta[k], i, j, h, w, align_corners);
        m = __ldg(mb + p * K + k);
        prev = k;
      }
      const float gm = gs[q] * m;
      float* r0 = dxb + ((int64_t)f.y0 * w + f.x0) * C + c + q;
      float* r1 = r0 + (int64_t)w * C;
      if (f.w00 != 0.f) atomicAdd(r0, gm * f.w00);
      if (f.w01 != 0.f) atomicAdd(r0 + C, gm * f.w01);
      if (f.w10 != 0.f) atomicAdd(r1, gm * f.w10);
      if (f.w11 != 0.f) atomicAdd(r1 + C, gm * f.w11);
    }
  }
}


// ---------------------------------------------------------------- cooperative kernels (C >= 64)
// A group of G = min(32, C/4) lanes owns one output pixel (a warp handles 32/G pixels at a time).  The per-part
// geometry (mask value, bilinear footprint) is computed ONCE per pixel by lane k of the group and broadcast with
// shuffles, so the per-element instruction stream is just the gather itself; parts whose mask is zero at the pixel
// (or whose footprint is outside the image) are skipped -- they all contribute the same candidate "0, no gradient".
struct PartGeom { float m; int x0, y0; float w00, w01, w10, w11; };

__device__ __forceinline__ PartGeom shfl_geom(const PartGeom& g, int src) {
  PartGeom r;
  r.m = __shfl_sync(0xffffffffu, g.m, src);
  r.x0 = __shfl_sync(0xffffffffu, g.x0, src);
  r.y0 = __shfl_sync(0xffffffffu, g.y0, src);
  r.w00 = __shfl_sync(0xffffffffu, g.w00, src);
  r.w01 = __shfl_sync(0xffffffffu, g.w01, src);
  r.w10 = __shfl_sync(0xffffffffu, g.w10, src);
  r.w11 = __shfl_sync(0xffffffffu, g.w11, src);
  return r;
}

__device__ __forceinline__ void fma4(float4& a, float w, const float4& v) {
  a.x = fmaf(w, v.x, a.x); a.y = fmaf(w, v.y, a.y); a.z = fmaf(w, v.z, a.z); a.w = fmaf(w, v.w, a.w);
}

// Packed per-(pixel, part) geometry that travels through the shuffles: 4 registers.
struct PackedGeom { float m, fx, fy; int xy; };   // xy = (y0 + 2) << 16 | (x0 + 2)   (x0, y0 in [-2, 32765])

__device__ __forceinline__ PackedGeom shfl_packed(const PackedGeom& g, int src) {
  PackedGeom r;
  r.m = __shfl_sync(0xffffffffu, g.m, src);
  r.fx = __shfl_sync(0xffffffffu, g.fx, src);
  r.fy = __shfl_sync(0xffffffffu, g.fy, src);
  r.xy = __shfl_sync(0xffffffffu, g.xy, src);
  return r;
}

// Forward, G lanes per pixel, 8 channels (two float4) per lane and chunk: C = 64 -> G = 8 (4 pixels per warp),
// C = 128 -> G = 16, C % 256 == 0 -> G = 32.  Lane gl of a group evaluates parts gl, gl + G, ... of its pixel.
template <int G>
__global__ void __launch_bounds__(256)
warp_forward_coop_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ warps,
                         const float* __restrict__ mask_lvl, float* __restrict__ y, int ldy, uint8_t* __restrict__ argk,
                         int C, int h, int w, int K, int H0, int W0, int act) {
  __shared__ Theta s_theta[kMaxParts];
  const int n = blockIdx.y;
  if (threadIdx.x < K) s_theta[threadIdx.x] = normalized_theta(warps + ((int64_t)n * K + threadIdx.x) * 8, h, w, H0, W0);
  __syncthreads();
  constexpr int PPW = 32 / G;                        // pixels per warp iteration
  constexpr int SLOTS = (kMaxParts + G - 1) / G;     // parts evaluated per lane
  const int lane = threadIdx.x & 31;
  const int gl = lane % G, gbase = lane - gl, grp = lane / G;
  const int HW = h * w;
  const float inv_w = 1.f / (float)w, inv_h = 1.f / (float)h;
  const float* xb = x + (int64_t)n * HW * ldx;
  const float* mb = mask_lvl + (int64_t)n * HW * K;
  float* yb = y + (int64_t)n * HW * ldy;
  uint8_t* ab = argk + (int64_t)n * HW * C;
  const int warp_id = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  const int iters = (HW + PPW - 1) / PPW;
  const unsigned kmask = (1u << K) - 1u;
  for (int it = warp_id; it < iters; it += nwarps) {
    const int p = it * PPW + grp;
    const bool pvalid = p < HW;
    const int i = pvalid ? p / w : 0, j = pvalid ? p - i * w : 0;
    PackedGeom mine[SLOTS];
    unsigned my_active = 0;                          // bit s: slot s is an active part
#pragma unroll
    for (int sl = 0; sl < SLOTS; ++sl) {
      const int k = gl + sl * G;
      mine[sl].m = 0.f; mine[sl].fx = 0.f; mine[sl].fy = 0.f; mine[sl].xy = 0;
      if (pvalid && k < K) {
        const float m = __ldg(mb + p * K + k);
        if (m != 0.f) {
          const Theta t = s_theta[k];
          const float gx = (2.f * (float)j + 1.f) * inv_w - 1.f, gy = (2.f * (float)i + 1.f) * inv_h - 1.f;
          const float px = ((t.a * gx + t.b * gy + t.tx + 1.f) * (float)w - 1.f) * 0.5f;
          const float py = ((t.c * gx + t.d * gy + t.ty + 1.f) * (float)h - 1.f) * 0.5f;
          const float fx0 = floorf(px), fy0 = floorf(py);
          const int x0 = (int)fminf(fmaxf(fx0, -2.f), (float)w), y0 = (int)fminf(fmaxf(fy0, -2.f), (float)h);
          if (x0 >= -1 && x0 < w && y0 >= -1 && y0 < h) {     // at least one neighbour inside the image
            my_active |= 1u << sl;
            mine[sl].m = m; mine[sl].fx = px - fx0; mine[sl].fy = py - fy0;
            mine[sl].xy = ((y0 + 2) << 16) | (x0 + 2);
          }
        }
      }
    }
    // active-part bit mask of every group (bit k), and their union over the warp
    unsigned active = 0, uni = 0;
#pragma unroll
    for (int sl = 0; sl < SLOTS; ++sl) {
      const unsigned ball = __ballot_sync(0xffffffffu, (my_active >> sl) & 1u);
      active |= ((ball >> gbase) & ((G == 32) ? 0xffffffffu : ((1u << G) - 1u))) << (sl * G);
      unsigned u = ball;
      if (G <= 16) u |= u >> 16;
      if (G <= 8) u |= u >> 8;
      u &= (G == 32) ? 0xffffffffu : ((1u << G) - 1u);
      uni |= u << (sl * G);
    }
    active &= kmask; uni &= kmask;
    const unsigned inactive = ~active & kmask;
    const int kz = inactive ? __ffs(inactive) - 1 : 1 << 20;   // where the (single) "0, no gradient" candidate sits
    // Each lane owns two float4 per chunk of G*8 channels: channels [cb + 4 gl, +4) and [cb + 4 G + 4 gl, +4), so that
    // every 128-bit load / store instruction of a group covers one CONTIGUOUS run of 16*G bytes (whole L1 lines and
    // full sectors per request instead of the half-used sectors of an "8 consecutive channels per lane" mapping).
    for (int cb = 0; cb < C; cb += G * 8) {
      const int c0 = cb + gl * 4, c1 = c0 + G * 4;
      float best[8];
      unsigned char arg[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) { best[q] = -INFINITY; arg[q] = 255; }
      bool zero_done = false;
      unsigned rem = uni;
      while (rem) {
        const int k = __ffs(rem) - 1;
        rem &= rem - 1;
        const int sl = k / G;
        PackedGeom g = shfl_packed(SLOTS == 1 ? mine[0] : (sl == 0 ? mine[0] : mine[SLOTS - 1]), gbase + (k - sl * G));
        if (!((active >> k) & 1u)) continue;           // active only for another pixel of this warp
        if (!zero_done && k > kz) {
          zero_done = true;
#pragma unroll
          for (int q = 0; q < 8; ++q) if (0.f > best[q]) { best[q] = 0.f; arg[q] = 255; }
        }
        const int x0 = (g.xy & 0xffff) - 2, y0 = (g.xy >> 16) - 2;
        const bool xin0 = x0 >= 0, xin1 = x0 + 1 < w, yin0 = y0 >= 0, yin1 = y0 + 1 < h;
        const float w00 = (yin0 && xin0) ? (1.f - g.fy) * (1.f - g.fx) * g.m : 0.f;
        const float w01 = (yin0 && xin1) ? (1.f - g.fy) * g.fx * g.m : 0.f;
        const float w10 = (yin1 && xin0) ? g.fy * (1.f - g.fx) * g.m : 0.f;
        const float w11 = (yin1 && xin1) ? g.fy * g.fx * g.m : 0.f;
        const float* r0 = xb + (y0 * w + x0) * ldx;
        const float* r1 = r0 + w * ldx;
        float4 ca = make_float4(0.f, 0.f, 0.f, 0.f), cc = ca;
        if (yin0 && xin0) { fma4(ca, w00, __ldg(reinterpret_cast<const float4*>(r0 + c0))); fma4(cc, w00, __ldg(reinterpret_cast<const float4*>(r0 + c1))); }
        if (yin0 && xin1) { fma4(ca, w01, __ldg(reinterpret_cast<const float4*>(r0 + ldx + c0))); fma4(cc, w01, __ldg(reinterpret_cast<const float4*>(r0 + ldx + c1))); }
        if (yin1 && xin0) { fma4(ca, w10, __ldg(reinterpret_cast<const float4*>(r1 + c0))); fma4(cc, w10, __ldg(reinterpret_cast<const float4*>(r1 + c1))); }
        if (yin1 && xin1) { fma4(ca, w11, __ldg(reinterpret_cast<const float4*>(r1 + ldx + c0))); fma4(cc, w11, __ldg(reinterpret_cast<const float4*>(r1 + ldx + c1))); }
        const float cand[8] = {ca.x, ca.y, ca.z, ca.w, cc.x, cc.y, cc.z, cc.w};
#pragma unroll
        for (int q = 0; q < 8; ++q) if (cand[q] > best[q]) { best[q] = cand[q]; arg[q] = (unsigned char)k; }
      }
      if (!zero_done && inactive) {
#pragma unroll
        for (int q = 0; q < 8; ++q) if (0.f > best[q]) { best[q] = 0.f; arg[q] = 255; }
      }
      if (pvalid) {
        float* dst = yb + p * ldy;
        *reinterpret_cast<float4*>(dst + c0) = make_float4(apply_act(best[0], act), apply_act(best[1], act), apply_act(best[2], act), apply_act(best[3], act));
        *reinterpret_cast<float4*>(dst + c1) = make_float4(apply_act(best[4], act), apply_act(best[5], act), apply_act(best[6], act), apply_act(best[7], act));
        *reinterpret_cast<uint32_t*>(ab + p * C + c0) = arg[0] | (arg[1] << 8) | (arg[2] << 16) | ((unsigned)arg[3] << 24);
        *reinterpret_cast<uint32_t*>(ab + p * C + c1) = arg[4] | (arg[5] << 8) | (arg[6] << 16) | ((unsigned)arg[7] << 24);
      }
    }
  }
}

// ---------------------------------------------------------------- tiled forward (K == KP parts, C >= 64)
// One CTA owns a strip of XW = 8 * (32 / G) pixels x TH rows of one image and walks it row by row, so that the two
// source rows a bilinear footprint touches stay in this SM's L1 from one row to the next (the row-major grid-stride
// order of the kernel above re-fetched every source row from L2: 37 % L1 hit rate, 2.8x the input bytes over the
// crossbar).  The strip's mask values are staged once in shared memory together with a per-pixel bit set of the
// parts whose mask is non-zero; the row loop then iterates over set bits only and reads mask value and transform
// from shared memory, which leaves the gather as the only global load on the dependent chain.  Every lane of a
// pixel's group evaluates the (cheap) footprint itself: no shuffles, no ballots.
// max over parts with torch.max's first-maximum rule: parts whose mask is zero (or whose footprint lies outside the
// image) all contribute the same candidate "0, no gradient" (argk = 255); the first of them is merged into the bit
// loop as a pseudo part at its own index kz, so real candidates before / after it keep their tie-breaking order.
template <int ACT>
__device__ __forceinline__ float warp_act(float v) {
  if (ACT == PTK_ACT_RELU) return fmaxf(v, 0.f);
  if (ACT == PTK_ACT_LEAKY) return v > 0.f ? v : 0.2f * v;
  return v;
}

// G lanes per pixel, NV float4 (4 NV channels) per lane and channel chunk: lane gl owns channels
// [cb + 4 G q + 4 gl, +4) for q < NV, so every 128-bit request of a group is one contiguous run of 16 G bytes.
// ACT == ReLU (the only way the network uses this layer: the warped skip is stored post-ReLU and the backward masks the
// gradient with y > 0) folds the "0, no gradient" candidates into the initial value max(., 0) -- no pseudo part.
template <int G, int NV, int KP, int ACT>
__global__ void __launch_bounds__(256, (NV > 2 ? 2 : 4))
warp_forward_tile_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ warps,
                         const float* __restrict__ mask_lvl, float* __restrict__ y, int ldy, uint8_t* __restrict__ argk,
                         int C, int h, int w, int H0, int W0, int TH, int strips_x, int PD) {
  static_assert(KP % 2 == 0 && KP <= kMaxParts, "mask rows are staged as float2");
  constexpr int PPW = 32 / G, XW = 8 * PPW, NC = 4 * NV;   // PD: L1 prefetch distance in rows (0 = off)
  constexpr bool kRelu = ACT == PTK_ACT_RELU;
  extern __shared__ float s_dyn[];                       // [TH * XW][KP] mask values, then [TH * XW] part bit sets
  __shared__ Theta s_theta[KP];
  float* s_m = s_dyn;
  unsigned* s_bits = reinterpret_cast<unsigned*>(s_dyn + TH * XW * KP);
  const int n = blockIdx.y;
  const int sx = blockIdx.x % strips_x, sy = blockIdx.x / strips_x;
  const int x_begin = sx * XW, y_begin = sy * TH;
  const int rows = min(TH, h - y_begin), cols = min(XW, w - x_begin);
  const int tid = threadIdx.x;
  const int HW = h * w;
  if (tid < KP) s_theta[tid] = normalized_theta(warps + ((int64_t)n * KP + tid) * 8, h, w, H0, W0);
  // stage: one thread per pixel of the strip (KP floats = KP/2 float2, 8-byte aligned because KP is even)
  const float* mb = mask_lvl + (int64_t)n * HW * KP;
  for (int t = tid; t < rows * XW; t += 256) {
    const int r = t / XW, cx = t - r * XW;
    unsigned bits = 0u;
    if (cx < cols) {
      const float2* src = reinterpret_cast<const float2*>(mb + ((y_begin + r) * w + x_begin + cx) * KP);
      float2 v[KP / 2];
#pragma unroll
      for (int q = 0; q < KP / 2; ++q) v[q] = __ldg(src + q);
#pragma unroll
      for (int q = 0; q < KP / 2; ++q) {
        *reinterpret_cast<float2*>(s_m + t * KP + 2 * q) = v[q];
        bits |= ((v[q].x != 0.f ? 1u : 0u) | (v[q].y != 0.f ? 2u : 0u)) << (2 * q);
      }
    }
    s_bits[t] = bits;
  }
  __syncthreads();

  const int lane = tid & 31, wi = tid >> 5;
  const int gl = lane % G, grp = lane / G;
  const int lx = wi * PPW + grp;                           // pixel column inside the strip
  const int j = x_begin + lx;
  if (j >= w) return;
  const float fw = (float)w, fh = (float)h;
  const float gx = (2.f * (float)j + 1.f) / fw - 1.f;
  const float* xb = x + (int64_t)n * HW * ldx + gl * 4;
  float* yrow = y + ((int64_t)n * HW + (int64_t)y_begin * w + j) * ldy + gl * 4;
  uint8_t* arow = argk + ((int64_t)n * HW + (int64_t)y_begin * w + j) * C + gl * 4;
  const int ystep = w * ldy, astep = w * C;
  constexpr unsigned kmask = (1u << KP) - 1u;
  for (int r = 0; r < rows; ++r, yrow += ystep, arow += astep) {
    const int lp = r * XW + lx;
    const unsigned bits = s_bits[lp];
    const float gy = (2.f * (float)(y_begin + r) + 1.f) / fh - 1.f;
    if (PD > 0 && y_begin + r + PD < h) {
      // L1 prefetch of the first part's footprint PD rows ahead: holds no registers, so the gather of row r + PD finds
      // its lines on chip and the row loop stops being bound by one DRAM round trip per row.  Rows past the end of the
      // strip (another CTA's) are prefetched with this row's part set as a guess: that warms L2 for the neighbour.
      const unsigned pb = r + PD < rows ? s_bits[lp + PD * XW] : bits;
      if (pb) {
        const int k = __ffs(pb) - 1;
        const Theta t = s_theta[k];
        const float gy2 = (2.f * (float)(y_begin + r + PD) + 1.f) / fh - 1.f;
        const float px = ((t.a * gx + t.b * gy2 + t.tx + 1.f) * fw - 1.f) * 0.5f;
        const float py = ((t.c * gx + t.d * gy2 + t.ty + 1.f) * fh - 1.f) * 0.5f;
        const int x0 = (int)fminf(fmaxf(floorf(px), -2.f), fw), y0 = (int)fminf(fmaxf(floorf(py), -2.f), fh);
        if (x0 >= -1 && x0 < w && y0 >= -1 && y0 < h) {
          const int xa = max(x0, 0), xc = min(x0 + 1, w - 1), yc = min(y0 + 1, h - 1);
          // the upper footprint row was (for near-identity transforms) the lower row of an earlier output row
          const float* p10 = xb + (yc * w + xa) * ldx;
          const float* p11 = xb + (yc * w + xc) * ldx;
          for (int cb = 0; cb < C; cb += G * NC) {
#pragma unroll
            for (int q = 0; q < NV; ++q) {
              asm volatile("prefetch.global.L1 [%0];" ::"l"(p10 + cb + q * G * 4));
              asm volatile("prefetch.global.L1 [%0];" ::"l"(p11 + cb + q * G * 4));
            }
          }
        }
      }
    }
    const unsigned inactive = ~bits & kmask;
    const int kz = (!kRelu && inactive) ? __ffs(inactive) - 1 : -1;
    // ReLU: real parts only.  Otherwise: real parts + the first zero candidate as a pseudo part at its own index.
    const unsigned todo = kRelu ? bits : (bits | (inactive & (0u - inactive)));
    const float* mrow = s_m + lp * KP;
    for (int cb = 0; cb < C; cb += G * NC) {
      float best[NC];
      int arg[NC];
#pragma unroll
      for (int q = 0; q < NC; ++q) { best[q] = kRelu ? 0.f : -INFINITY; arg[q] = 255; }
      unsigned rem = todo;
      while (rem) {
        const int k = __ffs(rem) - 1;
        rem &= rem - 1;
        float4 acc[NV];
#pragma unroll
        for (int q = 0; q < NV; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        int tag = 255;
        if (kRelu || k != kz) {
          const Theta t = s_theta[k];
          const float m = mrow[k];
          const float px = ((t.a * gx + t.b * gy + t.tx + 1.f) * fw - 1.f) * 0.5f;
          const float py = ((t.c * gx + t.d * gy + t.ty + 1.f) * fh - 1.f) * 0.5f;
          const float fx0 = floorf(px), fy0 = floorf(py);
          const int x0 = (int)fminf(fmaxf(fx0, -2.f), fw), y0 = (int)fminf(fmaxf(fy0, -2.f), fh);
          if (x0 >= -1 && x0 < w && y0 >= -1 && y0 < h) {        // otherwise: footprint outside = "0, no gradient"
            tag = k;
            const float fx = px - fx0, fy = py - fy0;
            // Out-of-image taps: weight 0 at a clamped (valid) address, so that all 4 NV loads are unconditional and
            // issue back to back (one memory round trip per part instead of one per tap).
            const bool xin0 = x0 >= 0, xin1 = x0 + 1 < w, yin0 = y0 >= 0, yin1 = y0 + 1 < h;
            const float wy0 = yin0 ? (1.f - fy) * m : 0.f, wy1 = yin1 ? fy * m : 0.f;
            const float wx0 = xin0 ? 1.f - fx : 0.f, wx1 = xin1 ? fx : 0.f;
            const float w00 = wy0 * wx0, w01 = wy0 * wx1, w10 = wy1 * wx0, w11 = wy1 * wx1;
            const int xa = max(x0, 0), xc = min(x0 + 1, w - 1), ya = max(y0, 0), yc = min(y0 + 1, h - 1);
            const float* p00 = xb + (ya * w + xa) * ldx + cb;
            const float* p01 = xb + (ya * w + xc) * ldx + cb;
            const float* p10 = xb + (yc * w + xa) * ldx + cb;
            const float* p11 = xb + (yc * w + xc) * ldx + cb;
            float4 v00[NV], v01[NV], v10[NV], v11[NV];
#pragma unroll
            for (int q = 0; q < NV; ++q) {
              v00[q] = __ldg(reinterpret_cast<const float4*>(p00 + q * G * 4));
              v01[q] = __ldg(reinterpret_cast<const float4*>(p01 + q * G * 4));
              v10[q] = __ldg(reinterpret_cast<const float4*>(p10 + q * G * 4));
              v11[q] = __ldg(reinterpret_cast<const float4*>(p11 + q * G * 4));
            }
#pragma unroll
            for (int q = 0; q < NV; ++q) {
              fma4(acc[q], w00, v00[q]); fma4(acc[q], w01, v01[q]); fma4(acc[q], w10, v10[q]); fma4(acc[q], w11, v11[q]);
            }
          }
        }
#pragma unroll
        for (int q = 0; q < NV; ++q) {
          if (acc[q].x > best[4 * q + 0]) { best[4 * q + 0] = acc[q].x; arg[4 * q + 0] = tag; }
          if (acc[q].y > best[4 * q + 1]) { best[4 * q + 1] = acc[q].y; arg[4 * q + 1] = tag; }
          if (acc[q].z > best[4 * q + 2]) { best[4 * q + 2] = acc[q].z; arg[4 * q + 2] = tag; }
          if (acc[q].w > best[4 * q + 3]) { best[4 * q + 3] = acc[q].w; arg[4 * q + 3] = tag; }
        }
      }
#pragma unroll
      for (int q = 0; q < NV; ++q) {
        *reinterpret_cast<float4*>(yrow + cb + q * G * 4) = make_float4(warp_act<ACT>(best[4 * q]), warp_act<ACT>(best[4 * q + 1]),
                                                                        warp_act<ACT>(best[4 * q + 2]), warp_act<ACT>(best[4 * q + 3]));
        *reinterpret_cast<uint32_t*>(arow + cb + q * G * 4) = (unsigned)arg[4 * q] | ((unsigned)arg[4 * q + 1] << 8) |
                                                              ((unsigned)arg[4 * q + 2] << 16) | ((unsigned)arg[4 * q + 3] << 24);
      }
    }
  }
}

// ---------------------------------------------------------------- tiled forward with an asynchronous row pipeline
// EXPERIMENT (PTK_WARP_PF=1, off by default): measured 10 % SLOWER than warp_forward_tile_kernel on B200 -- ncu shows
// the extra LDGSTS + LDS traffic pushing the L1/TEX pipe to 80 % busy, so the kernel trades a latency bound for an L1
// bound.  Kept as the documented negative result (profiles/README.md).
// Same strip walk as warp_forward_tile_kernel (ReLU epilogue only), but the gather of the FIRST active part of row r + 1
// (the body part, present at every pixel) is issued with cp.async into a lane-private shared-memory slot while row r is
// still being reduced and stored: the copies hold no registers, so every warp keeps two rows of loads in flight and the
// row loop is no longer bound by one DRAM round trip per row.  Further parts of a pixel (sparse) use direct loads.
__device__ __forceinline__ void cp_async16(uint32_t dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

template <int G, int NV, int KP>
__global__ void __launch_bounds__(256, 2)
warp_forward_pf_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ warps,
                       const float* __restrict__ mask_lvl, float* __restrict__ y, int ldy, uint8_t* __restrict__ argk,
                       int C, int h, int w, int H0, int W0, int TH, int strips_x) {
  static_assert(KP % 2 == 0 && KP <= kMaxParts, "mask rows are staged as float2");
  constexpr int PPW = 32 / G, XW = 8 * PPW, NC = 4 * NV;
  extern __shared__ __align__(16) float s_dyn[];           // [4 NV][256] float4 gather slots, [TH*XW][KP] masks, [TH*XW] bit sets
  __shared__ Theta s_theta[KP];
  float4* s_slot = reinterpret_cast<float4*>(s_dyn);
  float* s_m = s_dyn + 4 * NV * 256 * 4;
  unsigned* s_bits = reinterpret_cast<unsigned*>(s_m + TH * XW * KP);
  const int n = blockIdx.y;
  const int sx = blockIdx.x % strips_x, sy = blockIdx.x / strips_x;
  const int x_begin = sx * XW, y_begin = sy * TH;
  const int rows = min(TH, h - y_begin), cols = min(XW, w - x_begin);
  const int tid = threadIdx.x;
  const int HW = h * w;
  if (tid < KP) s_theta[tid] = normalized_theta(warps + ((int64_t)n * KP + tid) * 8, h, w, H0, W0);
  const float* mb = mask_lvl + (int64_t)n * HW * KP;
  for (int t = tid; t < rows * XW; t += 256) {
    const int r = t / XW, cx = t - r * XW;
    unsigned bits = 0u;
    if (cx < cols) {
      const float2* src = reinterpret_cast<const float2*>(mb + ((y_begin + r) * w + x_begin + cx) * KP);
      float2 v[KP / 2];
#pragma unroll
      for (int q = 0; q < KP / 2; ++q) v[q] = __ldg(src + q);
#pragma unroll
      for (int q = 0; q < KP / 2; ++q) {
        *reinterpret_cast<float2*>(s_m + t * KP + 2 * q) = v[q];
        bits |= ((v[q].x != 0.f ? 1u : 0u) | (v[q].y != 0.f ? 2u : 0u)) << (2 * q);
      }
    }
    s_bits[t] = bits;
  }
  __syncthreads();

  const int lane = tid & 31, wi = tid >> 5;
  const int gl = lane % G, grp = lane / G;
  const int lx = wi * PPW + grp;
  const int j = x_begin + lx;
  if (j >= w) return;
  const float fw = (float)w, fh = (float)h;
  const float gx = (2.f * (float)j + 1.f) / fw - 1.f;
  const float* xb = x + (int64_t)n * HW * ldx + gl * 4;
  float* yrow = y + ((int64_t)n * HW + (int64_t)y_begin * w + j) * ldy + gl * 4;
  uint8_t* arow = argk + ((int64_t)n * HW + (int64_t)y_begin * w + j) * C + gl * 4;
  const int ystep = w * ldy, astep = w * C;
  const uint32_t slot0 = (uint32_t)__cvta_generic_to_shared(s_slot + tid);      // slot s of this lane: + s * 256 float4

  // footprint of part k at row (y_begin + r): returns false if it lies outside the image
  struct Geo { int x0, y0; float fx, fy; };
  auto footprint_of = [&](int k, int r, Geo& g) -> bool {
    const Theta t = s_theta[k];
    const float gy = (2.f * (float)(y_begin + r) + 1.f) / fh - 1.f;
    const float px = ((t.a * gx + t.b * gy + t.tx + 1.f) * fw - 1.f) * 0.5f;
    const float py = ((t.c * gx + t.d * gy + t.ty + 1.f) * fh - 1.f) * 0.5f;
    const float fx0 = floorf(px), fy0 = floorf(py);
    g.x0 = (int)fminf(fmaxf(fx0, -2.f), fw); g.y0 = (int)fminf(fmaxf(fy0, -2.f), fh);
    g.fx = px - fx0; g.fy = py - fy0;
    return g.x0 >= -1 && g.x0 < w && g.y0 >= -1 && g.y0 < h;
  };
  // issue the asynchronous gather of the first active part of row r (channel chunk 0); returns its index or -1
  auto prefetch_row = [&](int r, Geo& g) -> int {
    const unsigned b = s_bits[r * XW + lx];
    if (!b) return -1;
    const int k = __ffs(b) - 1;
    if (!footprint_of(k, r, g)) return -1;
    const int xa = max(g.x0, 0), xc = min(g.x0 + 1, w - 1), ya = max(g.y0, 0), yc = min(g.y0 + 1, h - 1);
    const float* p00 = xb + (ya * w + xa) * ldx;
    const float* p01 = xb + (ya * w + xc) * ldx;
    const float* p10 = xb + (yc * w + xa) * ldx;
    const float* p11 = xb + (yc * w + xc) * ldx;
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      cp_async16(slot0 + (uint32_t)((0 * NV + q) * 256 * 16), p00 + q * G * 4);
      cp_async16(slot0 + (uint32_t)((1 * NV + q) * 256 * 16), p01 + q * G * 4);
      cp_async16(slot0 + (uint32_t)((2 * NV + q) * 256 * 16), p10 + q * G * 4);
      cp_async16(slot0 + (uint32_t)((3 * NV + q) * 256 * 16), p11 + q * G * 4);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    return k;
  };

  Geo gcur;
  int pk = prefetch_row(0, gcur);
  for (int r = 0; r < rows; ++r, yrow += ystep, arow += astep) {
    const int lp = r * XW + lx;
    const unsigned bits = s_bits[lp];
    const float* mrow = s_m + lp * KP;
    float best[NC];
    int arg[NC];
#pragma unroll
    for (int q = 0; q < NC; ++q) { best[q] = 0.f; arg[q] = 255; }      // ReLU: max(., 0), "no gradient"
    unsigned rem = bits;
    Geo gnext;
    int pk_next = -1;
    if (pk >= 0) {
      // ---- first part: operands arrive through the lane-private shared-memory slots
      rem &= rem - 1;
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      float4 acc[NV];
      {
        const float m = mrow[pk];
        const bool xin0 = gcur.x0 >= 0, xin1 = gcur.x0 + 1 < w, yin0 = gcur.y0 >= 0, yin1 = gcur.y0 + 1 < h;
        const float wy0 = yin0 ? (1.f - gcur.fy) * m : 0.f, wy1 = yin1 ? gcur.fy * m : 0.f;
        const float wx0 = xin0 ? 1.f - gcur.fx : 0.f, wx1 = xin1 ? gcur.fx : 0.f;
        const float w00 = wy0 * wx0, w01 = wy0 * wx1, w10 = wy1 * wx0, w11 = wy1 * wx1;
#pragma unroll
        for (int q = 0; q < NV; ++q) {
          acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
          fma4(acc[q], w00, s_slot[(0 * NV + q) * 256 + tid]);
          fma4(acc[q], w01, s_slot[(1 * NV + q) * 256 + tid]);
          fma4(acc[q], w10, s_slot[(2 * NV + q) * 256 + tid]);
          fma4(acc[q], w11, s_slot[(3 * NV + q) * 256 + tid]);
        }
      }
      // the slots are consumed (the FMAs above depend on every shared-memory read): refill them for the next row
      if (r + 1 < rows) pk_next = prefetch_row(r + 1, gnext);
#pragma unroll
      for (int q = 0; q < NV; ++q) {
        if (acc[q].x > best[4 * q + 0]) { best[4 * q + 0] = acc[q].x; arg[4 * q + 0] = pk; }
        if (acc[q].y > best[4 * q + 1]) { best[4 * q + 1] = acc[q].y; arg[4 * q + 1] = pk; }
        if (acc[q].z > best[4 * q + 2]) { best[4 * q + 2] = acc[q].z; arg[4 * q + 2] = pk; }
        if (acc[q].w > best[4 * q + 3]) { best[4 * q + 3] = acc[q].w; arg[4 * q + 3] = pk; }
      }
    } else if (r + 1 < rows) {
      pk_next = prefetch_row(r + 1, gnext);
    }
    // ---- remaining parts (and every part of further channel chunks): direct loads
    for (int cb = 0; cb < C; cb += G * NC) {
      if (cb > 0) {
        // store the finished chunk, restart the reduction for the next one with ALL parts
        const int co = cb - G * NC;
#pragma unroll
        for (int q = 0; q < NV; ++q) {
          *reinterpret_cast<float4*>(yrow + co + q * G * 4) = make_float4(best[4 * q], best[4 * q + 1], best[4 * q + 2], best[4 * q + 3]);
          *reinterpret_cast<uint32_t*>(arow + co + q * G * 4) = (unsigned)arg[4 * q] | ((unsigned)arg[4 * q + 1] << 8) |
                                                                ((unsigned)arg[4 * q + 2] << 16) | ((unsigned)arg[4 * q + 3] << 24);
        }
#pragma unroll
        for (int q = 0; q < NC; ++q) { best[q] = 0.f; arg[q] = 255; }
        rem = bits;
      }
      while (rem) {
        const int k = __ffs(rem) - 1;
        rem &= rem - 1;
        Geo g;
        if (!footprint_of(k, r, g)) continue;
        const float m = mrow[k];
        const bool xin0 = g.x0 >= 0, xin1 = g.x0 + 1 < w, yin0 = g.y0 >= 0, yin1 = g.y0 + 1 < h;
        const float wy0 = yin0 ? (1.f - g.fy) * m : 0.f, wy1 = yin1 ? g.fy * m : 0.f;
        const float wx0 = xin0 ? 1.f - g.fx : 0.f, wx1 = xin1 ? g.fx : 0.f;
        const float w00 = wy0 * wx0, w01 = wy0 * wx1, w10 = wy1 * wx0, w11 = wy1 * wx1;
        const int xa = max(g.x0, 0), xc = min(g.x0 + 1, w - 1), ya = max(g.y0, 0), yc = min(g.y0 + 1, h - 1);
        const float* p00 = xb + (ya * w + xa) * ldx + cb;
        const float* p01 = xb + (ya * w + xc) * ldx + cb;
        const float* p10 = xb + (yc * w + xa) * ldx + cb;
        const float* p11 = xb + (yc * w + xc) * ldx + cb;
        float4 v00[NV], v01[NV], v10[NV], v11[NV];
#pragma unroll
        for (int q = 0; q < NV; ++q) {
          v00[q] = __ldg(reinterpret_cast<const float4*>(p00 + q * G * 4));
          v01[q] = __ldg(reinterpret_cast<const float4*>(p01 + q * G * 4));
          v10[q] = __ldg(reinterpret_cast<const float4*>(p10 + q * G * 4));
          v11[q] = __ldg(reinterpret_cast<const float4*>(p11 + q * G * 4));
        }
#pragma unroll
        for (int q = 0; q < NV; ++q) {
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
          fma4(a, w00, v00[q]); fma4(a, w01, v01[q]); fma4(a, w10, v10[q]); fma4(a, w11, v11[q]);
          if (a.x > best[4 * q + 0]) { best[4 * q + 0] = a.x; arg[4 * q + 0] = k; }
          if (a.y > best[4 * q + 1]) { best[4 * q + 1] = a.y; arg[4 * q + 1] = k; }
          if (a.z > best[4 * q + 2]) { best[4 * q + 2] = a.z; arg[4 * q + 2] = k; }
          if (a.w > best[4 * q + 3]) { best[4 * q + 3] = a.w; arg[4 * q + 3] = k; }
        }
      }
    }
    {
      const int co = ((C - 1) / (G * NC)) * (G * NC);
#pragma unroll
      for (int q = 0; q < NV; ++q) {
        *reinterpret_cast<float4*>(yrow + co + q * G * 4) = make_float4(best[4 * q], best[4 * q + 1], best[4 * q + 2], best[4 * q + 3]);
        *reinterpret_cast<uint32_t*>(arow + co + q * G * 4) = (unsigned)arg[4 * q] | ((unsigned)arg[4 * q + 1] << 8) |
                                                              ((unsigned)arg[4 * q + 2] << 16) | ((unsigned)arg[4 * q + 3] << 24);
      }
    }
    pk = pk_next;
    gcur = gnext;
  }
}

template <int G>
__global__ void __launch_bounds__(256)
warp_backward_coop_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ y, int ldy, int act,
                          const float* __restrict__ warps, const float* __restrict__ mask_lvl,
                          const uint8_t* __restrict__ argk, float* __restrict__ dx, int C, int h, int w, int K, int H0,
                          int W0, int align_corners) {
  __shared__ Theta s_theta[kMaxParts];
  const int n = blockIdx.y;
  if (threadIdx.x < K) s_theta[threadIdx.x] = normalized_theta(warps + ((int64_t)n * K + threadIdx.x) * 8, h, w, H0, W0);
  __syncthreads();
  constexpr int PPW = 32 / G;
  const int lane = threadIdx.x & 31;
  const int gl = lane % G, gbase = lane - gl;
  const int64_t HW = (int64_t)h * w;
  const float* mb = mask_lvl + (int64_t)n * HW * K;
  float* dxb = dx + (int64_t)n * HW * C;
  const int64_t warp_id = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t iters = (HW + PPW - 1) / PPW;
  for (int64_t it = warp_id; it < iters; it += nwarps) {
    const int64_t p = it * PPW + lane / G;
    const bool pvalid = p < HW;
    const int i = pvalid ? (int)(p / w) : 0, j = pvalid ? (int)(p - (int64_t)i * w) : 0;
    PartGeom mine;
    mine.m = 0.f; mine.x0 = 0; mine.y0 = 0; mine.w00 = mine.w01 = mine.w10 = mine.w11 = 0.f;
    if (pvalid && gl < K) {
      mine.m = __ldg(mb + p * K + gl);
      if (mine.m != 0.f) {
        const Footprint f = footprint(s_theta[gl], i, j, h, w, align_corners);
        mine.x0 = f.x0; mine.y0 = f.y0; mine.w00 = f.w00; mine.w01 = f.w01; mine.w10 = f.w10; mine.w11 = f.w11;
      }
    }
    for (int c0 = gl * 4; c0 < C; c0 += G * 4) {
      uchar4 a4 = make_uchar4(255, 255, 255, 255);
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (pvalid) {
        a4 = *reinterpret_cast<const uchar4*>(argk + ((int64_t)n * HW + p) * C + c0);
        g = __ldg(reinterpret_cast<const float4*>(dy + ((int64_t)n * HW + p) * lddy + c0));
        if (act != PTK_ACT_NONE) {
          const float4 yv = __ldg(reinterpret_cast<const float4*>(y + ((int64_t)n * HW + p) * ldy + c0));
          g.x *= act_grad_from_output(yv.x, act); g.y *= act_grad_from_output(yv.y, act);
          g.z *= act_grad_from_output(yv.z, act); g.w *= act_grad_from_output(yv.w, act);
        }
      }
      const unsigned char as[4] = {a4.x, a4.y, a4.z, a4.w};
      const float gs[4] = {g.x, g.y, g.z, g.w};
      const bool uniform = a4.x == a4.y && a4.y == a4.z && a4.z == a4.w;
      // every lane takes part in the shuffles; lanes without a winner read part 0 and add nothing
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int k = as[q] == 255 ? 0 : as[q];
        const PartGeom f = shfl_geom(mine, gbase + k);
        if (as[q] == 255) continue;
        if (uniform) {
          if (q > 0) continue;
          float* r0 = dxb + ((int64_t)f.y0 * w + f.x0) * C + c0;
          float* r1 = r0 + (int64_t)w * C;
          const float4 gm = make_float4(g.x * f.m, g.y * f.m, g.z * f.m, g.w * f.m);
          if (f.w00 != 0.f) atomicAdd(reinterpret_cast<float4*>(r0), make_float4(gm.x * f.w00, gm.y * f.w00, gm.z * f.w00, gm.w * f.w00));
          if (f.w01 != 0.f) atomicAdd(reinterpret_cast<float4*>(r0 + C), make_float4(gm.x * f.w01, gm.y * f.w01, gm.z * f.w01, gm.w * f.w01));
          if (f.w10 != 0.f) atomicAdd(reinterpret_cast<float4*>(r1), make_float4(gm.x * f.w10, gm.y * f.w10, gm.z * f.w10, gm.w * f.w10));
          if (f.w11 != 0.f) atomicAdd(reinterpret_cast<float4*>(r1 + C), make_float4(gm.x * f.w11, gm.y * f.w11, gm.z * f.w11, gm.w * f.w11));
        } else if (gs[q] != 0.f) {
          const float gm = gs[q] * f.m;
          float* r0 = dxb + ((int64_t)f.y0 * w + f.x0) * C + c0 + q;
          float* r1 = r0 + (int64_t)w * C;
          if (f.w00 != 0.f) atomicAdd(r0, gm * f.w00);
          if (f.w01 != 0.f) atomicAdd(r0 + C, gm * f.w01);
          if (f.w10 != 0.f) atomicAdd(r1, gm * f.w10);
          if (f.w11 != 0.f) atomicAdd(r1 + C, gm * f.w11);
        }
      }
    }
  }
}

// cv2.resize(INTER_LINEAR) == half-pixel bilinear; computed in double like the reference (masks are f64).
// A CTA produces 256 consecutive output pixels of one image for all K parts: the source planes are read part by part
// (coalesced along x), the [pixel][part] rows are transposed through shared memory and written as one contiguous run.
constexpr int kPyrPix = 256;
__global__ void __launch_bounds__(256)
mask_pyramid_kernel(const double* __restrict__ masks, int K, int H0, int W0, float* __restrict__ out, int h, int w) {
  extern __shared__ float s_row[];                 // [kPyrPix][K + 1]
  const int n = blockIdx.y;
  const int hw = h * w;
  const int p0 = blockIdx.x * kPyrPix;
  const int p = p0 + threadIdx.x;
  const double sy = (double)H0 / h, sx = (double)W0 / w;
  if (p < hw) {
    const int i = p / w, j = p - i * w;
    double fy = (i + 0.5) * sy - 0.5, fx = (j + 0.5) * sx - 0.5;
    if (fy < 0) fy = 0;
    if (fx < 0) fx = 0;
    int y0 = (int)fy, x0 = (int)fx;
    if (y0 > H0 - 1) y0 = H0 - 1;
    if (x0 > W0 - 1) x0 = W0 - 1;
    const int y1 = y0 + 1 < H0 ? y0 + 1 : H0 - 1, x1 = x0 + 1 < W0 ? x0 + 1 : W0 - 1;
    const double ly = fy - y0, lx = fx - x0;
    const bool exact = ly == 0.0 && lx == 0.0;      // same-size level: a plain conversion
    for (int k = 0; k < K; ++k) {
      const double* src = masks + ((int64_t)n * K + k) * (int64_t)H0 * W0;
      double v;
      if (exact) v = src[(int64_t)y0 * W0 + x0];
      else v = (1 - ly) * ((1 - lx) * src[(int64_t)y0 * W0 + x0] + lx * src[(int64_t)y0 * W0 + x1]) +
               ly * ((1 - lx) * src[(int64_t)y1 * W0 + x0] + lx * src[(int64_t)y1 * W0 + x1]);
      s_row[threadIdx.x * (K + 1) + k] = (float)v;
    }
  }
  __syncthreads();
  const int npix = min(kPyrPix, hw - p0);
  float* dst = out + ((int64_t)n * hw + p0) * K;
  for (int e = threadIdx.x; e < npix * K; e += 256) {
    const int pp = e / K, k = e - pp * K;
    dst[e] = s_row[pp * (K + 1) + k];
  }
}

}  // namespace ptk

using namespace ptk;

static inline dim3 warp_grid(int64_t work, int N) {
  int64_t b = (work + 255) / 256;
  int64_t cap = ((int64_t)num_sms() * 8 + N - 1) / N;
  if (cap < 1) cap = 1;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return dim3((unsigned)b, (unsigned)N);
}

extern "C" int ptk_mask_pyramid(const double* masks, int N, int K, int H0, int W0, float* out, int h, int w,
                                void* stream) {
  PTK_REQUIRE(N > 0 && N <= 65535 && K > 0 && K <= 64 && h > 0 && w > 0, "mask_pyramid: bad extents");
  dim3 grid((unsigned)(((int64_t)h * w + kPyrPix - 1) / kPyrPix), (unsigned)N);
  mask_pyramid_kernel<<<grid, 256, kPyrPix * (K + 1) * sizeof(float), (cudaStream_t)stream>>>(masks, K, H0, W0, out, h, w);
  PTK_LAUNCH_CHECK("mask_pyramid_kernel");
  return 0;
}

extern "C" int ptk_warp_forward(const float* x, int ldx, const float* warps, const float* mask_lvl, float* y,
                                int ldy, uint8_t* argk, int N, int C, int h, int w, int K, int H0, int W0,
                                int align_corners, int act, void* stream) {
  PTK_REQUIRE(N > 0 && N <= 65535 && C > 0 && C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0, "warp_forward: C/ld must be multiples of 4");
  PTK_REQUIRE(K > 0 && K <= kMaxParts, "warp_forward: K must be in [1,%d]", kMaxParts);
  PTK_REQUIRE(act == PTK_ACT_NONE || act == PTK_ACT_RELU || act == PTK_ACT_LEAKY, "warp_forward: bad act");
  // the cooperative kernel uses 32-bit element offsets inside one image and packs (x0, y0) in 16 bits each
  const bool coop_ok = !align_corners && h < 32000 && w < 32000 && (int64_t)h * w * (ldx > ldy ? ldx : ldy) < (1ll << 31);
  if (coop_ok && K == 10 && (C == 64 || C == 128 || C == 256 || C % 512 == 0)) {
    // NV float4 per lane and chunk (PTK_WARP_NV = 2 | 4, default 4): G = C / (4 NV) lanes per pixel, at most 32
    static int pd_env = -1;
    if (pd_env < 0) { const char* e = getenv("PTK_WARP_PD"); pd_env = e ? atoi(e) : 0; if (pd_env < 0 || pd_env > 8) pd_env = 0; }
    static int nv_env = -1;
    if (nv_env < 0) { const char* e = getenv("PTK_WARP_NV"); nv_env = (e && atoi(e) == 2) ? 2 : 4; }
    const int NV = nv_env;
    const int G = C / (4 * NV) >= 32 ? 32 : C / (4 * NV);
    const int XW = 8 * (32 / G);
    const int strips_x = (w + XW - 1) / XW;
    // rows per strip: 3..8, chosen for the fullest last wave
    const int occ = NV == 4 ? 2 : 4;
    int TH = 8;
    double best_eff = -1.0;
    for (int th = 8; th >= 3; --th) {
      if (th > h) continue;
      const int64_t blocks = (int64_t)strips_x * ((h + th - 1) / th) * N, slots = (int64_t)num_sms() * occ;
      const double eff = (double)blocks / (double)((blocks + slots - 1) / slots * slots) * (th / (th + 1.0));
      if (eff > best_eff) { best_eff = eff; TH = th; }
    }
    if (TH > h) TH = h;
    const int strips_y = (h + TH - 1) / TH;
    const size_t smem = (size_t)TH * XW * (10 + 1) * sizeof(float);
    dim3 grid((unsigned)(strips_x * strips_y), (unsigned)N);
#define PTK_WARP_TILE(G_, NV_, A_) warp_forward_tile_kernel<G_, NV_, 10, A_><<<grid, 256, smem, (cudaStream_t)stream>>>(x, ldx, warps, mask_lvl, y, ldy, argk, C, h, w, H0, W0, TH, strips_x, pd_env)
#define PTK_WARP_TILE_G(G_, NV_) do { if (act == PTK_ACT_RELU) PTK_WARP_TILE(G_, NV_, PTK_ACT_RELU); else if (act == PTK_ACT_LEAKY) PTK_WARP_TILE(G_, NV_, PTK_ACT_LEAKY); else PTK_WARP_TILE(G_, NV_, PTK_ACT_NONE); } while (0)
    static int pf_env = -1;
    if (pf_env < 0) { const char* e = getenv("PTK_WARP_PF"); pf_env = (e && atoi(e) == 1) ? 1 : 0; }   // measured slower: opt-in
    if (NV == 4 && act == PTK_ACT_RELU && pf_env) {
      const size_t smem_pf = smem + (size_t)4 * 4 * 256 * 16;
#define PTK_WARP_PF(G_)                                                                                                   \
  do {                                                                                                                    \
    static bool attr = false;                                                                                             \
    if (!attr) { cudaFuncSetAttribute(warp_forward_pf_kernel<G_, 4, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); attr = true; } \
    warp_forward_pf_kernel<G_, 4, 10><<<grid, 256, smem_pf, (cudaStream_t)stream>>>(x, ldx, warps, mask_lvl, y, ldy, argk, C, h, w, H0, W0, TH, strips_x); \
  } while (0)
      if (G == 4) PTK_WARP_PF(4);
      else if (G == 8) PTK_WARP_PF(8);
      else if (G == 16) PTK_WARP_PF(16);
      else PTK_WARP_PF(32);
#undef PTK_WARP_PF
    } else if (NV == 4) {
      if (G == 4) PTK_WARP_TILE_G(4, 4);
      else if (G == 8) PTK_WARP_TILE_G(8, 4);
      else if (G == 16) PTK_WARP_TILE_G(16, 4);
      else PTK_WARP_TILE_G(32, 4);
    } else {
      if (G == 8) PTK_WARP_TILE_G(8, 2);
      else if (G == 16) PTK_WARP_TILE_G(16, 2);
      else PTK_WARP_TILE_G(32, 2);
    }
#undef PTK_WARP_TILE_G
#undef PTK_WARP_TILE
  } else if (coop_ok && C == 64) {
    warp_forward_coop_kernel<8><<<warp_grid((int64_t)h * w * 8, N), 256, 0, (cudaStream_t)stream>>>(
        x, ldx, warps, mask_lvl, y, ldy, argk, C, h, w, K, H0, W0, act);
  } else if (coop_ok && C == 128) {
    warp_forward_coop_kernel<16><<<warp_grid((int64_t)h * w * 16, N), 256, 0, (cudaStream_t)stream>>>(
        x, ldx, warps, mask_lvl, y, ldy, argk, C, h, w, K, H0, W0, act);
  } else if (coop_ok && C % 256 == 0) {
    warp_forward_coop_kernel<32><<<warp_grid((int64_t)h * w * 32, N), 256, 0, (cudaStream_t)stream>>>(
        x, ldx, warps, mask_lvl, y, ldy, argk, C, h, w, K, H0, W0, act);
  } else {
    warp_forward_kernel<<<warp_grid((int64_t)h * w * (C / 4), N), 256, 0, (cudaStream_t)stream>>>(
        x, ldx, warps, mask_lvl, y, ldy, argk, C, h, w, K, H0, W0, align_corners, act);
  }
  PTK_LAUNCH_CHECK("warp_forward_kernel");
  return 0;
}

extern "C" int ptk_warp_backward(const float* dy, int lddy, const float* y, int ldy, int act, const float* warps,
                                 const float* mask_lvl, const uint8_t* argk, float* dx, int N, int C, int h, int w,
                                 int K, int H0, int W0, int align_corners, void* stream) {
  PTK_REQUIRE(N > 0 && N <= 65535 && C > 0 && C % 4 == 0 && lddy % 4 == 0, "warp_backward: C/ld must be multiples of 4");
  PTK_REQUIRE(K > 0 && K <= kMaxParts, "warp_backward: K must be in [1,%d]", kMaxParts);
  PTK_REQUIRE(act == PTK_ACT_NONE || (y != nullptr && ldy % 4 == 0), "warp_backward: y required for act backward");
  if (C == 64) {
    warp_backward_coop_kernel<16><<<warp_grid((int64_t)h * w * 16, N), 256, 0, (cudaStream_t)stream>>>(
        dy, lddy, y, ldy, act, warps, mask_lvl, argk, dx, C, h, w, K, H0, W0, align_corners);
  } else if (C % 128 == 0) {
    warp_backward_coop_kernel<32><<<warp_grid((int64_t)h * w * 32, N), 256, 0, (cudaStream_t)stream>>>(
        dy, lddy, y, ldy, act, warps, mask_lvl, argk, dx, C, h, w, K, H0, W0, align_corners);
  } else {
    warp_backward_kernel<<<warp_grid((int64_t)h * w * (C / 4), N), 256, 0, (cudaStream_t)stream>>>(
        dy, lddy, y, ldy, act, warps, mask_lvl, argk, dx, C, h, w, K, H0, W0, align_corners);
  }
  PTK_LAUNCH_CHECK("warp_backward_kernel");
  return 0;
}
