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


// ---------------------------------------------------------------- fast path (C = 64 / 128 / multiples of 256, K <= 15)
// Record-staged tile kernels.  A CTA owns a tile of 32 x TH output pixels of one level and every warp a run of them:
//   A  one LANE per pixel evaluates the geometry of that pixel's active parts once (mask value, affine grid point,
//      bilinear weights incl. the mask factor, clamped tap offsets) and parks it in shared memory as 32-byte records plus
//      one header word per pixel (count, the part index of each record, the position of the first "0, no gradient"
//      candidate).  Parts whose mask is zero at the pixel (or whose footprint lies outside the image) get no record: they
//      all contribute the same candidate "0, no gradient".
//   B  the warp gathers (forward) or scatters (backward): G = 8 / 16 / 32 lanes own ONE pixel, lane gl holding channels
//      [cb + 4 G q + 4 gl, +4) for q < NV, so that every 128-bit request of a group is one contiguous run of 16 G bytes
//      (an "8-16 channels per lane" mapping touches half-used lines and doubles the L1 wavefronts per byte); per (pixel,
//      part) the lanes read the record with two broadcast loads and do nothing but 4 NV loads, 16 NV FMAs and the
//      running max (forward) or 4 NV vector reductions (backward).
// History (profiles/README.md): a per-row-geometry "strip" predecessor re-did the coordinate arithmetic in the 16-32 lanes
// of every pixel -- 160 M (forward) / 167 M (backward) warp instructions per launch, issue-bound at 0.36 / 0.14 of the
// HBM roofline; the records cut that to 67 M and 0.55 / 0.33.
// The winner record is packed to 4 bits per element (15 = "0, no gradient").
constexpr int kNoPart = 15;

struct WarpLevelDev {
  const float* x; const float* mask; float* y; uint8_t* argk; const float* dy; float* dx;
  int ldx, ldy, lddy, C, h, w, TH, strips_x, strips_y, cta_begin, cfg;     // cfg: 0 = C 64, 1 = C 128, 2 = C % 256 == 0, 3 = C % 512 == 0 (lane mappings in the kernels' dispatch)
};
struct WarpLaunchDev {
  WarpLevelDev lv[4];
  int nlevels, K, H0, W0, act, ctas_per_image, prefetch;
};

// geometry of part `gl` at output pixel (i, j); returns false for "0, no gradient" (mask 0 / footprint outside)
__device__ __forceinline__ bool part_geometry(const Theta& t, float m, int i, int j, int h, int w, int ld, float4& wgt, int4& off) {
  if (m == 0.f) return false;
  // explicit rounding intrinsics: forward and backward must take bit-identical decisions whatever the inlining context
  const float fw = (float)w, fh = (float)h;
  const float gx = __fsub_rn(__fdiv_rn(__fmaf_rn(2.f, (float)j, 1.f), fw), 1.f);
  const float gy = __fsub_rn(__fdiv_rn(__fmaf_rn(2.f, (float)i, 1.f), fh), 1.f);
  const float sx = __fmaf_rn(t.a, gx, __fmaf_rn(t.b, gy, t.tx)), sy = __fmaf_rn(t.c, gx, __fmaf_rn(t.d, gy, t.ty));
  const float px = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(sx, 1.f), fw), 1.f), 0.5f);
  const float py = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(sy, 1.f), fh), 1.f), 0.5f);
  const float fx0 = floorf(px), fy0 = floorf(py);
  const int x0 = (int)fminf(fmaxf(fx0, -2.f), fw), y0 = (int)fminf(fmaxf(fy0, -2.f), fh);
  if (!(x0 >= -1 && x0 < w && y0 >= -1 && y0 < h)) return false;
  const float fx = __fsub_rn(px, fx0), fy = __fsub_rn(py, fy0);
  // out-of-image taps: weight 0 at a clamped (valid) address, so that all loads are unconditional
  const bool xin0 = x0 >= 0, xin1 = x0 + 1 < w, yin0 = y0 >= 0, yin1 = y0 + 1 < h;
  const float wy0 = yin0 ? __fmul_rn(__fsub_rn(1.f, fy), m) : 0.f, wy1 = yin1 ? __fmul_rn(fy, m) : 0.f;
  const float wx0 = xin0 ? __fsub_rn(1.f, fx) : 0.f, wx1 = xin1 ? fx : 0.f;
  wgt = make_float4(__fmul_rn(wy0, wx0), __fmul_rn(wy0, wx1), __fmul_rn(wy1, wx0), __fmul_rn(wy1, wx1));
  const int xa = max(x0, 0), xc = min(x0 + 1, w - 1), ya = max(y0, 0), yc = min(y0 + 1, h - 1);
  off = make_int4((ya * w + xa) * ld, (ya * w + xc) * ld, (yc * w + xa) * ld, (yc * w + xc) * ld);
  return true;
}

__device__ __forceinline__ void fma4(float4& a, float w, const float4& v) {
  a.x = fmaf(w, v.x, a.x); a.y = fmaf(w, v.y, a.y); a.z = fmaf(w, v.z, a.z); a.w = fmaf(w, v.w, a.w);
}

template <int ACT>
__device__ __forceinline__ float warp_act(float v) {
  if (ACT == PTK_ACT_RELU) return fmaxf(v, 0.f);
  if (ACT == PTK_ACT_LEAKY) return v > 0.f ? v : 0.2f * v;
  return v;
}

constexpr int kTileW = 32, kTileH = 8, kMaxRec = 6;     // records per pixel held in shared memory (more: inline slow path)

struct FwdTileSmem {
  Theta theta[kMaxParts];
  uint32_t hdr[kTileW * kTileH];                 // bits [0,24): part index of record r at 4r; [24,27): count; 27: overflow;
                                                 // [28,32): kz (first part that contributes the zero candidate), 15 = none
  float4 rec[kTileW * kTileH * kMaxRec * 2];
};

// Phase A of the record-staged kernels: every warp evaluates the geometry of ITS OWN run of pixels (32 * TH / 8 of one
// tile row, one lane per pixel) and parks it in shared memory; only a __syncwarp() separates it from the warp's gather /
// scatter phase, so the warps of a CTA drift apart and the ALU-bound geometry of one overlaps the memory phase of the
// others (ncu on the CTA-wide variant with a __syncthreads(): 1.5 barrier-stall cycles per issued instruction).
// ld = pixel stride of the tensor the tap offsets address (x in the forward pass, dx in the backward pass).
template <int TH>
__device__ __forceinline__ void tile_records(const WarpLevelDev& L, FwdTileSmem& S, int K, int64_t img, int x0, int y0, int ld) {
  constexpr int SPR = 8 / TH, SEG = kTileW / SPR;
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const int row = wi / SPR, col = (wi % SPR) * SEG + lane;
  const int i = y0 + row, j = x0 + col;
  const int h = L.h, w = L.w;
  const unsigned kmask = (1u << K) - 1u;
  if (lane < SEG) {
    const int t = row * kTileW + col;
    uint32_t hdr = 15u << 28;
    if (i < h && j < w) {
      const float* mp = L.mask + (img + (int64_t)i * w + j) * K;
      int cnt = 0;
      unsigned seen = 0u;
      for (int k = 0; k < K; ++k) {
        const float m = __ldg(mp + k);
        float4 wgt;
        int4 off;
        if (part_geometry(S.theta[k], m, i, j, h, w, ld, wgt, off)) {
          seen |= 1u << k;
          if (cnt < kMaxRec) {
            S.rec[(t * kMaxRec + cnt) * 2] = wgt;
            S.rec[(t * kMaxRec + cnt) * 2 + 1] = make_float4(__int_as_float(off.x), __int_as_float(off.y), __int_as_float(off.z), __int_as_float(off.w));
            hdr |= (uint32_t)k << (4 * cnt);
            ++cnt;
          } else {
            hdr |= 1u << 27;
          }
        }
      }
      const unsigned inactive = ~seen & kmask;
      const uint32_t kz = inactive ? (uint32_t)(__ffs(inactive) - 1) : 15u;
      hdr = (hdr & 0x0fffffffu) | ((uint32_t)cnt << 24) | (kz << 28);
    }
    S.hdr[t] = hdr;
  }
  __syncwarp();
}

// TH = tile rows: every CTA carries the same number of bytes (32 x TH pixels x C channels = 64 KB of output), so the four
// levels of a pass balance over the machine: TH = 8 / 4 / 2 / 1 for C = 64 / 128 / 256 / >= 512.  A warp owns a run of
// 32 * TH / 8 pixels of one tile row; G lanes hold one pixel (16 NV-byte... 4 NV channels per lane at stride 4 G: every
// 128-bit request of the group is whole lines), 32 / G pixels per warp are in flight.
template <int G, int NV, int TH, int PX, int ACT>
__device__ __forceinline__ void warp_fwd_tile(const WarpLevelDev& L, FwdTileSmem& S, int K, int n, int tile, int prefetch) {
  constexpr int PPW = 32 / G, CH = 4 * G * NV, SPR = 8 / TH, SEG = kTileW / SPR, ITERS = SEG / (PPW * PX);
  static_assert(SEG % (PPW * PX) == 0 && ITERS >= 1, "segment must be a whole number of warp iterations");
  constexpr bool kRelu = ACT == PTK_ACT_RELU;
  const int h = L.h, w = L.w, C = L.C, ldx = L.ldx, ldy = L.ldy;
  const int tx = tile % L.strips_x, ty = tile / L.strips_x;
  const int x0 = tx * kTileW, y0 = ty * TH;
  const int64_t img = (int64_t)n * h * w;
  // ------------------------------------------------------------------ phase A: geometry records of this warp's pixels
  tile_records<TH>(L, S, K, img, x0, y0, ldx);
  // ------------------------------------------------------------------ phase B: gather
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const int gl = lane % G, grp = lane / G;
  const int row = wi / SPR, seg = wi % SPR;
  const int i = y0 + row;
  if (i >= h) return;
  const float* xb = L.x + img * ldx + gl * 4;
  float* yb = L.y + (img + (int64_t)i * w) * ldy + gl * 4;
  uint8_t* ab = L.argk + (((img + (int64_t)i * w) * C) >> 1) + gl * 2;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    // PX pixels per lane group in flight: tile columns col0 + px * PPW (the groups of a warp interleave, so that one
    // 128-bit request of the warp still covers PPW adjacent pixels)
    const int col0 = seg * SEG + it * (PPW * PX) + grp;
    if (x0 + col0 >= w) continue;
    uint32_t hd[PX];
    int cnt[PX], nrec = 0;
#pragma unroll
    for (int px = 0; px < PX; ++px) {
      const int col = col0 + px * PPW;
      hd[px] = S.hdr[row * kTileW + col];
      cnt[px] = (x0 + col < w) ? (int)((hd[px] >> 24) & 7u) : 0;
      nrec = max(nrec, cnt[px]);
    }
    const float4* rbase = S.rec + (row * kTileW + col0) * (kMaxRec * 2);
    if (prefetch && it + 1 < ITERS && x0 + col0 + PPW * PX < w) {
      // warm L2 with the lower tap row of the first (body) part of the pixels this group handles next
#pragma unroll
      for (int px = 0; px < PX; ++px) {
        const int tn = row * kTileW + col0 + (PX + px) * PPW;
        if ((S.hdr[tn] >> 24) & 7u) {
          const float4 of = S.rec[tn * (kMaxRec * 2) + 1];
#pragma unroll
          for (int q = 0; q < NV; ++q) asm volatile("prefetch.global.L2 [%0];" ::"l"(xb + q * G * 4 + __float_as_int(of.w)));
        }
      }
    }
#pragma unroll 1
    for (int cb = 0; cb < C; cb += CH) {
      float best[PX][NV * 4];
      int arg[PX][NV * 4];
      bool zdone[PX];
#pragma unroll
      for (int px = 0; px < PX; ++px) {
        zdone[px] = kRelu;
#pragma unroll
        for (int q = 0; q < NV * 4; ++q) { best[px][q] = kRelu ? 0.f : -INFINITY; arg[px][q] = kNoPart; }
      }
#pragma unroll 1
      for (int r = 0; r < nrec; ++r) {
        float4 wv[PX];
        float4 v[PX][NV][4];
#pragma unroll
        for (int px = 0; px < PX; ++px) {
          const bool on = r < cnt[px];
          wv[px] = on ? rbase[(px * PPW * kMaxRec + r) * 2] : zero4;
          const float4 of = on ? rbase[(px * PPW * kMaxRec + r) * 2 + 1] : zero4;
          const float* p = xb + cb;
#pragma unroll
          for (int q = 0; q < NV; ++q) {
            v[px][q][0] = on ? __ldg(reinterpret_cast<const float4*>(p + q * G * 4 + __float_as_int(of.x))) : zero4;
            v[px][q][1] = on ? __ldg(reinterpret_cast<const float4*>(p + q * G * 4 + __float_as_int(of.y))) : zero4;
            v[px][q][2] = on ? __ldg(reinterpret_cast<const float4*>(p + q * G * 4 + __float_as_int(of.z))) : zero4;
            v[px][q][3] = on ? __ldg(reinterpret_cast<const float4*>(p + q * G * 4 + __float_as_int(of.w))) : zero4;
          }
        }
#pragma unroll
        for (int px = 0; px < PX; ++px) {
          if (r >= cnt[px]) continue;
          const int kk = (int)((hd[px] >> (4 * r)) & 15u);
          if (!kRelu && !zdone[px] && kk > (int)(hd[px] >> 28)) {      // the zero candidate sits before this part
            zdone[px] = true;
#pragma unroll
            for (int q = 0; q < NV * 4; ++q) if (0.f > best[px][q]) { best[px][q] = 0.f; arg[px][q] = kNoPart; }
          }
#pragma unroll
          for (int q = 0; q < NV; ++q) {
            float4 acc = zero4;
            fma4(acc, wv[px].x, v[px][q][0]); fma4(acc, wv[px].y, v[px][q][1]);
            fma4(acc, wv[px].z, v[px][q][2]); fma4(acc, wv[px].w, v[px][q][3]);
            if (acc.x > best[px][4 * q + 0]) { best[px][4 * q + 0] = acc.x; arg[px][4 * q + 0] = kk; }
            if (acc.y > best[px][4 * q + 1]) { best[px][4 * q + 1] = acc.y; arg[px][4 * q + 1] = kk; }
            if (acc.z > best[px][4 * q + 2]) { best[px][4 * q + 2] = acc.z; arg[px][4 * q + 2] = kk; }
            if (acc.w > best[px][4 * q + 3]) { best[px][4 * q + 3] = acc.w; arg[px][4 * q + 3] = kk; }
          }
        }
      }
#pragma unroll
      for (int px = 0; px < PX; ++px) {
        const int j = x0 + col0 + px * PPW;
        if (j >= w) continue;
        const int kz = (int)(hd[px] >> 28);
        if (hd[px] & (1u << 27)) {
          // more active parts than records (never seen in practice): the remaining parts, geometry evaluated inline
          const int last = (int)((hd[px] >> (4 * (kMaxRec - 1))) & 15u);
          for (int k = last + 1; k < K; ++k) {
            float4 wgt;
            int4 off;
            const float m = __ldg(L.mask + (img + (int64_t)i * w + j) * K + k);
            if (!part_geometry(S.theta[k], m, i, j, h, w, ldx, wgt, off)) continue;
            if (!kRelu && !zdone[px] && k > kz) {
              zdone[px] = true;
#pragma unroll
              for (int q = 0; q < NV * 4; ++q) if (0.f > best[px][q]) { best[px][q] = 0.f; arg[px][q] = kNoPart; }
            }
#pragma unroll
            for (int q = 0; q < NV; ++q) {
              const float* p = xb + cb + q * G * 4;
              float4 acc = zero4;
              fma4(acc, wgt.x, __ldg(reinterpret_cast<const float4*>(p + off.x))); fma4(acc, wgt.y, __ldg(reinterpret_cast<const float4*>(p + off.y)));
              fma4(acc, wgt.z, __ldg(reinterpret_cast<const float4*>(p + off.z))); fma4(acc, wgt.w, __ldg(reinterpret_cast<const float4*>(p + off.w)));
              if (acc.x > best[px][4 * q + 0]) { best[px][4 * q + 0] = acc.x; arg[px][4 * q + 0] = k; }
              if (acc.y > best[px][4 * q + 1]) { best[px][4 * q + 1] = acc.y; arg[px][4 * q + 1] = k; }
              if (acc.z > best[px][4 * q + 2]) { best[px][4 * q + 2] = acc.z; arg[px][4 * q + 2] = k; }
              if (acc.w > best[px][4 * q + 3]) { best[px][4 * q + 3] = acc.w; arg[px][4 * q + 3] = k; }
            }
          }
        }
        if (!kRelu && !zdone[px] && kz != 15) {      // zero candidate after the last real part
#pragma unroll
          for (int q = 0; q < NV * 4; ++q) if (0.f > best[px][q]) { best[px][q] = 0.f; arg[px][q] = kNoPart; }
        }
#pragma unroll
        for (int q = 0; q < NV; ++q) {
          *reinterpret_cast<float4*>(yb + j * ldy + cb + q * G * 4) =
              make_float4(warp_act<ACT>(best[px][4 * q]), warp_act<ACT>(best[px][4 * q + 1]), warp_act<ACT>(best[px][4 * q + 2]),
                          warp_act<ACT>(best[px][4 * q + 3]));
          *reinterpret_cast<uint16_t*>(ab + ((j * C + cb + q * G * 4) >> 1)) =
              (uint16_t)(arg[px][4 * q] | (arg[px][4 * q + 1] << 4) | (arg[px][4 * q + 2] << 8) | (arg[px][4 * q + 3] << 12));
        }
      }
    }
  }
}

// Backward on the same tiles and records (tap offsets address dx): dx[taps of the winner] += dy * (mask * bilinear weight).
// A lane's 4 channels usually share their winner: one 128-bit vector reduction per tap and record, with the channels won by
// other parts zeroed; zero-weight taps are skipped.  ReLU / none need no look at y: "no winner" (15) == zero candidate
// == no gradient.  (ncu on the per-row-geometry predecessor: 167 M warp instructions per launch, 2.5x the forward pass.)
__device__ __forceinline__ void scatter4(float* p, const float4& wv, const int4& of, const float4& gm) {
  if (wv.x != 0.f) atomicAdd(reinterpret_cast<float4*>(p + of.x), make_float4(gm.x * wv.x, gm.y * wv.x, gm.z * wv.x, gm.w * wv.x));
  if (wv.y != 0.f) atomicAdd(reinterpret_cast<float4*>(p + of.y), make_float4(gm.x * wv.y, gm.y * wv.y, gm.z * wv.y, gm.w * wv.y));
  if (wv.z != 0.f) atomicAdd(reinterpret_cast<float4*>(p + of.z), make_float4(gm.x * wv.z, gm.y * wv.z, gm.z * wv.z, gm.w * wv.z));
  if (wv.w != 0.f) atomicAdd(reinterpret_cast<float4*>(p + of.w), make_float4(gm.x * wv.w, gm.y * wv.w, gm.z * wv.w, gm.w * wv.w));
}

template <int G, int NV, int TH, int PX>
__device__ __forceinline__ void warp_bwd_tile(const WarpLevelDev& L, FwdTileSmem& S, int K, int n, int tile, int act) {
  constexpr int PPW = 32 / G, CH = 4 * G * NV, SPR = 8 / TH, SEG = kTileW / SPR, ITERS = SEG / (PPW * PX);
  static_assert(SEG % (PPW * PX) == 0 && ITERS >= 1, "segment must be a whole number of warp iterations");
  const int h = L.h, w = L.w, C = L.C;
  const int tx = tile % L.strips_x, ty = tile / L.strips_x;
  const int x0 = tx * kTileW, y0 = ty * TH;
  const int64_t img = (int64_t)n * h * w;
  tile_records<TH>(L, S, K, img, x0, y0, C);
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const int gl = lane % G, grp = lane / G;
  const int row = wi / SPR, seg = wi % SPR;
  const int i = y0 + row;
  if (i >= h) return;
  float* dxb = L.dx + img * C + gl * 4;
  const float* dyb = L.dy + (img + (int64_t)i * w) * L.lddy + gl * 4;
  const float* yb = (act == PTK_ACT_LEAKY && L.y) ? L.y + (img + (int64_t)i * w) * L.ldy + gl * 4 : nullptr;
  const uint8_t* ab = L.argk + (((img + (int64_t)i * w) * C) >> 1) + gl * 2;
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    const int col0 = seg * SEG + it * (PPW * PX) + grp;
    if (x0 + col0 >= w) continue;
    uint32_t hd[PX];
    int cnt[PX];
#pragma unroll
    for (int px = 0; px < PX; ++px) {
      const int col = col0 + px * PPW;
      hd[px] = S.hdr[row * kTileW + col];
      cnt[px] = (x0 + col < w) ? (int)((hd[px] >> 24) & 7u) : 0;
    }
    const float4* rbase = S.rec + (row * kTileW + col0) * (kMaxRec * 2);
#pragma unroll 1
    for (int cb = 0; cb < C; cb += CH) {
      uint32_t a16[PX][NV];
      float4 g[PX][NV];
#pragma unroll
      for (int px = 0; px < PX; ++px) {
        const int j = x0 + col0 + px * PPW;
#pragma unroll
        for (int q = 0; q < NV; ++q) {
          a16[px][q] = j < w ? *reinterpret_cast<const uint16_t*>(ab + ((j * C + cb + q * G * 4) >> 1)) : 0xffffu;
          g[px][q] = a16[px][q] != 0xffffu ? __ldg(reinterpret_cast<const float4*>(dyb + j * L.lddy + cb + q * G * 4))
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      if (yb != nullptr) {
#pragma unroll
        for (int px = 0; px < PX; ++px) {
          const int j = x0 + col0 + px * PPW;
#pragma unroll
          for (int q = 0; q < NV; ++q) {
            if (a16[px][q] == 0xffffu) continue;
            const float4 yv = __ldg(reinterpret_cast<const float4*>(yb + j * L.ldy + cb + q * G * 4));
            g[px][q].x *= act_grad_from_output(yv.x, act); g[px][q].y *= act_grad_from_output(yv.y, act);
            g[px][q].z *= act_grad_from_output(yv.z, act); g[px][q].w *= act_grad_from_output(yv.w, act);
          }
        }
      }
#pragma unroll
      for (int px = 0; px < PX; ++px) {
        const int j = x0 + col0 + px * PPW;
#pragma unroll 1
        for (int r = 0; r < cnt[px]; ++r) {
          const uint32_t kk = (hd[px] >> (4 * r)) & 15u;
          const float4 wv = rbase[(px * PPW * kMaxRec + r) * 2];
          const float4 ofv = rbase[(px * PPW * kMaxRec + r) * 2 + 1];
          const int4 of = make_int4(__float_as_int(ofv.x), __float_as_int(ofv.y), __float_as_int(ofv.z), __float_as_int(ofv.w));
#pragma unroll
          for (int q = 0; q < NV; ++q) {
            const uint32_t a = a16[px][q];
            const bool m0 = (a & 15u) == kk, m1 = ((a >> 4) & 15u) == kk, m2 = ((a >> 8) & 15u) == kk, m3 = ((a >> 12) & 15u) == kk;
            if (!(m0 || m1 || m2 || m3)) continue;
            const float4 gv = g[px][q];
            scatter4(dxb + cb + q * G * 4, wv, of, make_float4(m0 ? gv.x : 0.f, m1 ? gv.y : 0.f, m2 ? gv.z : 0.f, m3 ? gv.w : 0.f));
          }
        }
        if (hd[px] & (1u << 27)) {
          // more active parts than records (never seen in practice): winners beyond the last record, geometry inline
          const int last = (int)((hd[px] >> (4 * (kMaxRec - 1))) & 15u);
          for (int k = last + 1; k < K; ++k) {
            float4 wgt;
            int4 off;
            const float m = __ldg(L.mask + (img + (int64_t)i * w + j) * K + k);
            if (j >= w || !part_geometry(S.theta[k], m, i, j, h, w, C, wgt, off)) continue;
#pragma unroll
            for (int q = 0; q < NV; ++q) {
              const uint32_t a = a16[px][q], kk = (uint32_t)k;
              const bool m0 = (a & 15u) == kk, m1 = ((a >> 4) & 15u) == kk, m2 = ((a >> 8) & 15u) == kk, m3 = ((a >> 12) & 15u) == kk;
              if (!(m0 || m1 || m2 || m3)) continue;
              const float4 gv = g[px][q];
              scatter4(dxb + cb + q * G * 4, wgt, off, make_float4(m0 ? gv.x : 0.f, m1 ? gv.y : 0.f, m2 ? gv.z : 0.f, m3 ? gv.w : 0.f));
            }
          }
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256, 3)
warp_backward_tiles_kernel(const __grid_constant__ WarpLaunchDev P, const float* __restrict__ warps) {
  extern __shared__ __align__(16) uint8_t s_raw[];
  FwdTileSmem& S = *reinterpret_cast<FwdTileSmem*>(s_raw);
  const int n = blockIdx.y;
  int li = 0;
#pragma unroll
  for (int q = 1; q < 4; ++q) if (q < P.nlevels && (int)blockIdx.x >= P.lv[q].cta_begin) li = q;
  const WarpLevelDev& L = P.lv[li];
  if (threadIdx.x < P.K) S.theta[threadIdx.x] = normalized_theta(warps + ((int64_t)n * P.K + threadIdx.x) * 8, L.h, L.w, P.H0, P.W0);
  __syncthreads();
  const int tile = blockIdx.x - L.cta_begin;
  switch (L.cfg) {
    case 0: warp_bwd_tile<8, 2, 8, 2>(L, S, P.K, n, tile, P.act); break;
    case 1: warp_bwd_tile<8, 4, 4, 1>(L, S, P.K, n, tile, P.act); break;
    case 2: warp_bwd_tile<16, 4, 2, 1>(L, S, P.K, n, tile, P.act); break;
    default: warp_bwd_tile<32, 4, 1, 1>(L, S, P.K, n, tile, P.act); break;
  }
}

template <int ACT>
__global__ void __launch_bounds__(256, 2)
warp_forward_tiles_kernel(const __grid_constant__ WarpLaunchDev P, const float* __restrict__ warps) {
  extern __shared__ __align__(16) uint8_t s_raw[];
  FwdTileSmem& S = *reinterpret_cast<FwdTileSmem*>(s_raw);
  const int n = blockIdx.y;
  int li = 0;
#pragma unroll
  for (int q = 1; q < 4; ++q) if (q < P.nlevels && (int)blockIdx.x >= P.lv[q].cta_begin) li = q;
  const WarpLevelDev& L = P.lv[li];
  if (threadIdx.x < P.K) S.theta[threadIdx.x] = normalized_theta(warps + ((int64_t)n * P.K + threadIdx.x) * 8, L.h, L.w, P.H0, P.W0);
  __syncthreads();
  const int tile = blockIdx.x - L.cta_begin;
  switch (L.cfg) {
    case 0: warp_fwd_tile<8, 2, 8, 2, ACT>(L, S, P.K, n, tile, P.prefetch); break;      // C = 64: two pixels per lane group
    case 1: warp_fwd_tile<8, 4, 4, 1, ACT>(L, S, P.K, n, tile, P.prefetch); break;      // C = 128
    case 2: warp_fwd_tile<16, 4, 2, 1, ACT>(L, S, P.K, n, tile, P.prefetch); break;     // C = 256 (and 768, ...)
    default: warp_fwd_tile<32, 4, 1, 1, ACT>(L, S, P.K, n, tile, P.prefetch); break;    // C % 512 == 0
  }
}

// zero-fill of up to four buffers in one launch (the dx targets of the backward scatter)
struct Fill4 { float4* p[4]; long long n4[4]; };
__global__ void __launch_bounds__(256) fill4_kernel(const __grid_constant__ Fill4 f) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int q = 0; q < 4; ++q)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < f.n4[q]; i += (long long)gridDim.x * blockDim.x) f.p[q][i] = z;
}

// cv2.resize(INTER_LINEAR) == half-pixel bilinear; computed in double like the reference (masks are f64).
// ONE launch builds the pyramid of a generator pass (up to four levels): blockIdx.x enumerates 256-pixel chunks of every
// level, blockIdx.y the image.  A CTA produces 256 consecutive output pixels for all K parts: the source planes are read
// four parts at a time (coalesced along x, up to 16 independent loads in flight per thread), the [pixel][part] rows are
// transposed through shared memory and written as one contiguous run.  The first level streams the f64 planes from
// HBM; the later levels re-read a subset of them out of L2.
constexpr int kPyrPix = 256;
struct PyrLaunch { float* out[4]; int h[4], w[4], cta_begin[4]; int nlevels, K, H0, W0; };

__global__ void __launch_bounds__(256)
mask_pyramid_kernel(const double* __restrict__ masks, const __grid_constant__ PyrLaunch P) {
  extern __shared__ float s_row[];                 // [kPyrPix][K + 1]
  pdl_trigger();
  const int n = blockIdx.y, K = P.K, H0 = P.H0, W0 = P.W0;
  int li = 0;
#pragma unroll
  for (int q = 1; q < 4; ++q) if (q < P.nlevels && (int)blockIdx.x >= P.cta_begin[q]) li = q;
  const int h = P.h[li], w = P.w[li];
  const int hw = h * w;
  const int p0 = ((int)blockIdx.x - P.cta_begin[li]) * kPyrPix;
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
    const int64_t plane = (int64_t)H0 * W0;
    const double* src = masks + (int64_t)n * K * plane;
    const int64_t o00 = (int64_t)y0 * W0 + x0, o01 = (int64_t)y0 * W0 + x1, o10 = (int64_t)y1 * W0 + x0, o11 = (int64_t)y1 * W0 + x1;
    for (int k0 = 0; k0 < K; k0 += 4) {
      double v00[4], v01[4], v10[4], v11[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const double* sp = src + (int64_t)min(k0 + q, K - 1) * plane;
        v00[q] = __ldg(sp + o00);
        if (!exact) { v01[q] = __ldg(sp + o01); v10[q] = __ldg(sp + o10); v11[q] = __ldg(sp + o11); }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (k0 + q >= K) break;
        const double v = exact ? v00[q] : (1 - ly) * ((1 - lx) * v00[q] + lx * v01[q]) + ly * ((1 - lx) * v10[q] + lx * v11[q]);
        s_row[threadIdx.x * (K + 1) + k0 + q] = (float)v;
      }
    }
  }
  __syncthreads();
  const int npix = min(kPyrPix, hw - p0);
  float* dst = P.out[li] + ((int64_t)n * hw + p0) * K;
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

extern "C" int ptk_mask_pyramid_levels(const double* masks, int N, int K, int H0, int W0, float* const* outs, const int* hs,
                                       const int* ws, int nlevels, void* stream) {
  PTK_REQUIRE(masks && outs && hs && ws && nlevels >= 1 && nlevels <= 4, "mask_pyramid: 1..4 levels");
  PTK_REQUIRE(N > 0 && N <= 65535 && K > 0 && K <= 64 && H0 > 0 && W0 > 0, "mask_pyramid: bad extents");
  PyrLaunch P;
  memset(&P, 0, sizeof(P));
  P.nlevels = nlevels; P.K = K; P.H0 = H0; P.W0 = W0;
  int ctas = 0;
  for (int q = 0; q < nlevels; ++q) {
    PTK_REQUIRE(outs[q] && hs[q] > 0 && ws[q] > 0, "mask_pyramid: bad level");
    P.out[q] = outs[q]; P.h[q] = hs[q]; P.w[q] = ws[q]; P.cta_begin[q] = ctas;
    ctas += (int)(((int64_t)hs[q] * ws[q] + kPyrPix - 1) / kPyrPix);
  }
  mask_pyramid_kernel<<<dim3((unsigned)ctas, (unsigned)N), 256, kPyrPix * (K + 1) * sizeof(float), (cudaStream_t)stream>>>(masks, P);
  PTK_LAUNCH_CHECK("mask_pyramid_kernel");
  return 0;
}

extern "C" int ptk_mask_pyramid(const double* masks, int N, int K, int H0, int W0, float* out, int h, int w,
                                void* stream) {
  return ptk_mask_pyramid_levels(masks, N, K, H0, W0, &out, &h, &w, 1, stream);
}

// configuration of the fast path for a level, or -1 (=> generic kernels, byte-sized winner record)
static int warp_fast_cfg(int C, int h, int w, int K, int ld_max, int align_corners) {
  if (align_corners || K > 15 || h >= 32000 || w >= 32000 || (int64_t)h * w * ld_max >= (1ll << 31)) return -1;
  if (C == 64) return 0;
  if (C == 128) return 1;
  if (C > 0 && C % 512 == 0) return 3;
  if (C > 0 && C % 256 == 0) return 2;
  return -1;
}

// fills the device-side launch description; returns false if some level cannot take the fast path
static bool warp_plan(const ptk_warp_level* lv, int nlevels, int K, int H0, int W0, int act, bool backward, WarpLaunchDev& P) {
  memset(&P, 0, sizeof(P));
  P.nlevels = nlevels; P.K = K; P.H0 = H0; P.W0 = W0; P.act = act;
  int ctas = 0;
  for (int q = 0; q < nlevels; ++q) {
    const ptk_warp_level& s = lv[q];
    const int ld_max = backward ? (s.lddy > s.C ? s.lddy : s.C) : (s.ldx > s.ldy ? s.ldx : s.ldy);
    const int cfg = warp_fast_cfg(s.C, s.h, s.w, K, ld_max, 0);
    if (cfg < 0) return false;
    WarpLevelDev& d = P.lv[q];
    d.x = s.x; d.mask = s.mask; d.y = s.y; d.argk = s.argk; d.dy = s.dy; d.dx = s.dx;
    d.ldx = s.ldx; d.ldy = s.ldy; d.lddy = s.lddy; d.C = s.C; d.h = s.h; d.w = s.w; d.cfg = cfg;
    // 32 x TH pixel tiles carrying 64 KB of output each, so that the levels of a pass balance over the machine
    const int xw = kTileW;
    d.TH = cfg == 0 ? 8 : (cfg == 1 ? 4 : (cfg == 2 ? 2 : 1));
    if (!backward) {
      const char* e = getenv("PTK_WARP_PF");
      P.prefetch = (e && atoi(e) == 0) ? 0 : 1;
    }
    d.strips_x = (s.w + xw - 1) / xw;
    d.strips_y = (s.h + d.TH - 1) / d.TH;
    d.cta_begin = ctas;
    ctas += d.strips_x * d.strips_y;
  }
  P.ctas_per_image = ctas;
  return true;
}

static int warp_check_level(const ptk_warp_level& s, bool backward) {
  PTK_REQUIRE(s.C > 0 && s.C % 4 == 0 && s.h > 0 && s.w > 0, "warp: C must be a positive multiple of 4");
  if (!backward) PTK_REQUIRE(s.x && s.y && s.mask && s.argk && s.ldx % 4 == 0 && s.ldy % 4 == 0, "warp_forward: x / y / mask / argk and 4-float strides required");
  else PTK_REQUIRE(s.dy && s.dx && s.mask && s.argk && s.lddy % 4 == 0, "warp_backward: dy / dx / mask / argk and 4-float strides required");
  return 0;
}

static int warp_forward_generic(const ptk_warp_level& s, const float* warps, int N, int K, int H0, int W0, int align_corners, int act,
                                cudaStream_t st) {
  warp_forward_kernel<<<warp_grid((int64_t)s.h * s.w * (s.C / 4), N), 256, 0, st>>>(s.x, s.ldx, warps, s.mask, s.y, s.ldy, s.argk, s.C,
                                                                                  s.h, s.w, K, H0, W0, align_corners, act);
  PTK_LAUNCH_CHECK("warp_forward_kernel");
  return 0;
}

static int warp_backward_generic(const ptk_warp_level& s, const float* warps, int N, int K, int H0, int W0, int align_corners, int act,
                                 cudaStream_t st) {
  warp_backward_kernel<<<warp_grid((int64_t)s.h * s.w * (s.C / 4), N), 256, 0, st>>>(s.dy, s.lddy, s.y, s.ldy, act, warps, s.mask, s.argk,
                                                                                   s.dx, s.C, s.h, s.w, K, H0, W0, align_corners);
  PTK_LAUNCH_CHECK("warp_backward_kernel");
  return 0;
}

extern "C" int ptk_warp_forward_levels(const ptk_warp_level* lv, int nlevels, const float* warps, int N, int K, int H0, int W0,
                                       int act, void* stream) {
  PTK_REQUIRE(nlevels >= 1 && nlevels <= 4 && N > 0 && N <= 65535, "warp_forward_levels: 1..4 levels, N in [1,65535]");
  PTK_REQUIRE(K > 0 && K <= kMaxParts, "warp_forward: K must be in [1,%d]", kMaxParts);
  PTK_REQUIRE(act == PTK_ACT_NONE || act == PTK_ACT_RELU || act == PTK_ACT_LEAKY, "warp_forward: bad act");
  for (int q = 0; q < nlevels; ++q) { int rc = warp_check_level(lv[q], false); if (rc) return rc; }
  cudaStream_t st = (cudaStream_t)stream;
  WarpLaunchDev P;
  if (!warp_plan(lv, nlevels, K, H0, W0, act, false, P)) {
    for (int q = 0; q < nlevels; ++q) { int rc = warp_forward_generic(lv[q], warps, N, K, H0, W0, 0, act, st); if (rc) return rc; }
    return 0;
  }
  dim3 grid((unsigned)P.ctas_per_image, (unsigned)N);
  {
    const size_t smem = sizeof(FwdTileSmem);
#define PTK_WARP_TILES(A_)                                                                                              \
  do {                                                                                                                  \
    static bool attr = false;                                                                                           \
    if (!attr) { cudaFuncSetAttribute(warp_forward_tiles_kernel<A_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; } \
    warp_forward_tiles_kernel<A_><<<grid, 256, smem, st>>>(P, warps);                                                   \
  } while (0)
    if (act == PTK_ACT_RELU) PTK_WARP_TILES(PTK_ACT_RELU);
    else if (act == PTK_ACT_LEAKY) PTK_WARP_TILES(PTK_ACT_LEAKY);
    else PTK_WARP_TILES(PTK_ACT_NONE);
#undef PTK_WARP_TILES
  }
  PTK_LAUNCH_CHECK("warp_forward_tiles_kernel");
  return 0;
}

extern "C" int ptk_warp_backward_levels(const ptk_warp_level* lv, int nlevels, const float* warps, int N, int K, int H0, int W0,
                                        int act, int zero_dx, void* stream) {
  PTK_REQUIRE(nlevels >= 1 && nlevels <= 4 && N > 0 && N <= 65535, "warp_backward_levels: 1..4 levels, N in [1,65535]");
  PTK_REQUIRE(K > 0 && K <= kMaxParts, "warp_backward: K must be in [1,%d]", kMaxParts);
  for (int q = 0; q < nlevels; ++q) {
    int rc = warp_check_level(lv[q], true);
    if (rc) return rc;
    PTK_REQUIRE(act == PTK_ACT_NONE || (lv[q].y != nullptr && lv[q].ldy % 4 == 0), "warp_backward: y required for act backward");
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (zero_dx) {
    Fill4 f;
    memset(&f, 0, sizeof(f));
    long long most = 0;
    for (int q = 0; q < nlevels; ++q) {
      PTK_REQUIRE((reinterpret_cast<uintptr_t>(lv[q].dx) & 15) == 0, "warp_backward: dx must be 16-byte aligned");
      f.p[q] = reinterpret_cast<float4*>(lv[q].dx);
      f.n4[q] = (long long)N * lv[q].h * lv[q].w * lv[q].C / 4;
      if (f.n4[q] > most) most = f.n4[q];
    }
    long long blocks = (most + 255) / 256;
    if (blocks > (long long)num_sms() * 8) blocks = (long long)num_sms() * 8;
    fill4_kernel<<<(unsigned)blocks, 256, 0, st>>>(f);
    PTK_LAUNCH_CHECK("fill4_kernel");
  }
  WarpLaunchDev P;
  if (!warp_plan(lv, nlevels, K, H0, W0, act, true, P)) {
    for (int q = 0; q < nlevels; ++q) { int rc = warp_backward_generic(lv[q], warps, N, K, H0, W0, 0, act, st); if (rc) return rc; }
    return 0;
  }
  if (act != PTK_ACT_LEAKY) for (int q = 0; q < nlevels; ++q) P.lv[q].y = nullptr;   // ReLU / none: the winner record says it all
  dim3 grid((unsigned)P.ctas_per_image, (unsigned)N);
  {
    const size_t smem = sizeof(FwdTileSmem);
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(warp_backward_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
    warp_backward_tiles_kernel<<<grid, 256, smem, st>>>(P, warps);
  }
  PTK_LAUNCH_CHECK("warp_backward_tiles_kernel");
  return 0;
}

extern "C" int ptk_warp_forward(const float* x, int ldx, const float* warps, const float* mask_lvl, float* y,
                                int ldy, uint8_t* argk, int N, int C, int h, int w, int K, int H0, int W0,
                                int align_corners, int act, void* stream) {
  ptk_warp_level s;
  memset(&s, 0, sizeof(s));
  s.x = x; s.ldx = ldx; s.mask = mask_lvl; s.y = y; s.ldy = ldy; s.argk = argk; s.C = C; s.h = h; s.w = w;
  if (!align_corners) return ptk_warp_forward_levels(&s, 1, warps, N, K, H0, W0, act, stream);
  PTK_REQUIRE(N > 0 && N <= 65535 && K > 0 && K <= kMaxParts, "warp_forward: bad N / K");
  PTK_REQUIRE(act == PTK_ACT_NONE || act == PTK_ACT_RELU || act == PTK_ACT_LEAKY, "warp_forward: bad act");
  int rc = warp_check_level(s, false);
  if (rc) return rc;
  return warp_forward_generic(s, warps, N, K, H0, W0, 1, act, (cudaStream_t)stream);
}

extern "C" int ptk_warp_backward(const float* dy, int lddy, const float* y, int ldy, int act, const float* warps,
                                 const float* mask_lvl, const uint8_t* argk, float* dx, int N, int C, int h, int w,
                                 int K, int H0, int W0, int align_corners, void* stream) {
  ptk_warp_level s;
  memset(&s, 0, sizeof(s));
  s.dy = dy; s.lddy = lddy; s.y = const_cast<float*>(y); s.ldy = ldy; s.mask = mask_lvl; s.argk = const_cast<uint8_t*>(argk);
  s.dx = dx; s.C = C; s.h = h; s.w = w;
  if (!align_corners) return ptk_warp_backward_levels(&s, 1, warps, N, K, H0, W0, act, 0, stream);
  PTK_REQUIRE(N > 0 && N <= 65535 && K > 0 && K <= kMaxParts, "warp_backward: bad N / K");
  PTK_REQUIRE(act == PTK_ACT_NONE || (y != nullptr && ldy % 4 == 0), "warp_backward: y required for act backward");
  int rc = warp_check_level(s, true);
  if (rc) return rc;
  return warp_backward_generic(s, warps, N, K, H0, W0, 1, act, (cudaStream_t)stream);
}
