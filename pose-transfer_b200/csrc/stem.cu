// Encoder stem: Conv2d(Cin <= 24 -> 64, k3, s1, p1, bias) (models/networks.py:186) on tcgen05 with the im2col done in
// SHARED memory.  The generic implicit-GEMM kernel (conv_tc.cu) fetches one activation box per tap from L2: for a stem
// that is 9 x the input over the crossbar (885 MB for a 67 MB tensor, 12 % tensor-pipe active).  Here a CTA stages the
// (8+2) x (16+2) pixel halo of its 8 x 16 pixel tile once, and four builder warps assemble the K-major, 128-byte-swizzled
// A operand (K = 9 taps x 24 channels = 216 -> 7 chunks of 32) chunk by chunk from that halo while a single thread
// issues the MMAs of the previous chunk (two-slot ring, generic -> async proxy fences, mbarriers).
//   D[pixel, co] = sum_{tap, c} halo[pixel + off(tap)][c] * W[co][tap*24 + c]      (TF32 operands, fp32 accumulate)
#include "tc_ptx.cuh"

namespace ptk {

constexpr int kStemCout = 64;
constexpr int kStemCpad = 24;                 // channels per tap in the K dimension
constexpr int kStemK = 9 * kStemCpad;         // 216
constexpr int kStemChunks = 7;                // ceil(216 / 32)
constexpr int kStemKpad = kStemChunks * 32;   // 224
constexpr int kTileH = 8, kTileW = 16;
constexpr int kHaloH = kTileH + 2, kHaloW = kTileW + 2;
constexpr int kHaloStride = 28;               // floats per halo pixel (24 + 4: conflict-free 128-bit reads at pixel stride)

// w_bwd layout [tap][64][cin_pad] (what the arena / pack kernels hold) -> wk[co][tap*24 + c], zero padded to 224
__global__ void stem_pack_kernel(const float* __restrict__ w, int cin, int cin_pad, float* __restrict__ wk) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kStemCout * kStemKpad; i += gridDim.x * blockDim.x) {
    const int co = i / kStemKpad, kk = i - co * kStemKpad;
    float v = 0.f;
    if (kk < kStemK) {
      const int tap = kk / kStemCpad, c = kk - tap * kStemCpad;
      if (c < cin) v = __ldg(w + ((int64_t)tap * kStemCout + co) * cin_pad + c);
    }
    wk[i] = v;
  }
}

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Persistent CTAs (2 per SM): the weight chunks are staged once per CTA, then every tile = stage halo -> build / MMA ring
// -> epilogue.  192 threads: warp 1 = TMEM + MMA issuer, warps 2..5 = builders + epilogue, all six warps stage.
__global__ void __launch_bounds__(192)
stem_conv_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ wk, const float* __restrict__ bias,
                 float* __restrict__ y, int ldy, int H, int W, int tiles_x, int tiles_per_img, int total_tiles) {
  constexpr uint32_t A_SLOT = 128 * 128, B_CHUNK = kStemCout * 128;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sA = base;                                   // 2 slots x 16 KB
  const uint32_t sB = sA + 2 * A_SLOT;                         // 7 chunks x 8 KB
  float* halo = reinterpret_cast<float*>(gbase + 2 * A_SLOT + kStemChunks * B_CHUNK);   // 180 px x 28 floats
  const uint32_t sBar = sB + kStemChunks * B_CHUNK + kHaloH * kHaloW * kHaloStride * 4;
  const uint32_t bar_full = sBar, bar_empty = sBar + 16, bar_tmem = sBar + 32, tmem_slot = sBar + 40;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(bar_full, 4); mbar_init(bar_full + 8, 4);       // one arrival per builder warp
    mbar_init(bar_empty, 1); mbar_init(bar_empty + 8, 1);
    mbar_init(bar_tmem, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // ---- weights (swizzled K-major chunks), once per CTA
  for (int i = tid; i < kStemCout * (kStemKpad / 4); i += 192) {
    const int co = i / (kStemKpad / 4), g = i - co * (kStemKpad / 4);      // g: 16-byte column 0..55
    const int c = g >> 3, col = g & 7;
    const float4 v = __ldg(reinterpret_cast<const float4*>(wk + (int64_t)co * kStemKpad + g * 4));
    *reinterpret_cast<float4*>(gbase + 2 * A_SLOT + c * B_CHUNK + (co >> 3) * 1024 + (co & 7) * 128 + ((col ^ (co & 7)) << 4)) = v;
  }
  fence_proxy_async();                 // the weight chunks are read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // The halo of the NEXT tile is fetched into registers right after the current one is published, so its global-memory
  // latency hides under the build / MMA / epilogue of the current tile.
  constexpr int kHaloVec = kHaloH * kHaloW * (kStemCpad / 4);      // 1080 float4
  constexpr int kHaloPer = (kHaloVec + 191) / 192;                  // 6 per thread
  float4 hreg[kHaloPer];
  auto fetch_halo = [&](int tile) {
    const int n = tile / tiles_per_img, tt = tile - n * tiles_per_img;
    const int ty0 = (tt / tiles_x) * kTileH, tx0 = (tt % tiles_x) * kTileW;
    const float* xb = x + (int64_t)n * H * W * ldx;
#pragma unroll
    for (int q = 0; q < kHaloPer; ++q) {
      const int i = tid + q * 192;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < kHaloVec) {
        const int p = i / (kStemCpad / 4), f = i - p * (kStemCpad / 4);
        const int yy = ty0 - 1 + p / kHaloW, xx = tx0 - 1 + p % kHaloW;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = __ldg(reinterpret_cast<const float4*>(xb + ((int64_t)yy * W + xx) * ldx + f * 4));
      }
      hreg[q] = v;
    }
  };
  if ((int)blockIdx.x < total_tiles) fetch_halo(blockIdx.x);
  int it = 0;                                                   // tiles done by this CTA (ring / barrier phases continue)
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
    const int n = tile / tiles_per_img, tt = tile - n * tiles_per_img;
    const int ty0 = (tt / tiles_x) * kTileH, tx0 = (tt % tiles_x) * kTileW;
    // ---- publish the halo (the previous tile's builders are past their last read: end-of-tile barrier)
#pragma unroll
    for (int q = 0; q < kHaloPer; ++q) {
      const int i = tid + q * 192;
      if (i < kHaloVec) {
        const int p = i / (kStemCpad / 4), f = i - p * (kStemCpad / 4);
        *reinterpret_cast<float4*>(halo + p * kHaloStride + f * 4) = hreg[q];
      }
    }
    __syncthreads();
    if (tile + (int)gridDim.x < total_tiles) fetch_halo(tile + gridDim.x);
    const int c0 = it * kStemChunks;                            // global chunk counter of this tile's first chunk

    if (warp == 1) {
      if (lane == 0) {
        constexpr uint32_t idesc = idesc_tf32(128, kStemCout);
#pragma unroll 1
        for (int c = 0; c < kStemChunks; ++c) {
          const int cg = c0 + c, s = cg & 1;
          mbar_wait(bar_full + 8 * s, (uint32_t)(cg >> 1) & 1u);
          tc_fence_after();
          const uint64_t da = smem_desc_k_sw128(sA + s * A_SLOT), db = smem_desc_k_sw128(sB + c * B_CHUNK);
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_tf32(tmem_base, da + 2 * k, db + 2 * k, idesc, (c > 0 || k > 0) ? 1u : 0u);
          tc_commit(bar_empty + 8 * s);
        }
        tc_commit(bar_tmem);
      }
      __syncwarp();
    } else if (warp >= 2) {
      // ---- builders: thread b owns pixel row b of the A tile
      const int b = tid - 64;
      const int py = b / kTileW, px = b % kTileW;
      const float* hrow = halo + (py * kHaloW + px) * kHaloStride;
      const uint32_t row_off = (uint32_t)((b >> 3) * 1024 + (b & 7) * 128);
#pragma unroll
      for (int c = 0; c < kStemChunks; ++c) {
        const int cg = c0 + c, s = cg & 1;
        if (cg >= 2) mbar_wait(bar_empty + 8 * s, (uint32_t)((cg >> 1) - 1) & 1u);
        uint8_t* dst = gbase + s * A_SLOT + row_off;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int g = c * 8 + i;               // 16-byte column of the K dimension: tap = g / 6, float4 f = g % 6
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (g < 9 * (kStemCpad / 4)) {
            const int tap = g / (kStemCpad / 4), f = g % (kStemCpad / 4);
            v = *reinterpret_cast<const float4*>(hrow + ((tap / 3) * kHaloW + (tap % 3)) * kHaloStride + f * 4);
          }
          *reinterpret_cast<float4*>(dst + ((i ^ (b & 7)) << 4)) = v;
        }
        fence_proxy_async();               // every lane orders its own generic-proxy writes before the async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_full + 8 * s);
      }
      // ---- epilogue (the next tile's first MMA is ordered behind these TMEM reads through the full barrier)
      const int lg = warp & 3;                    // TMEM lane group of this warp
      const int row = lg * 32 + lane;
      const int oy = ty0 + row / kTileW, ox = tx0 + row % kTileW;
      mbar_wait(bar_tmem, (uint32_t)it & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < kStemCout / 32; ++c) {
        float v[32];
        tc_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(c * 32), v);
        if (oy < H && ox < W) {
          float* dst = y + (((int64_t)n * H + oy) * W + ox) * ldy + c * 32;
          if (bias != nullptr) {
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] += __ldg(bias + c * 32 + q);
          }
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(dst + q * 4) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
      }
      tc_fence_before();
    }
    __syncthreads();            // every halo read / TMEM read of this tile is done before the next tile overwrites them
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
  }
}

}  // namespace ptk

using namespace ptk;

extern "C" int ptk_stem_pack(const float* w_tap_major, int cin, int cin_pad, float* wk, void* stream) {
  PTK_REQUIRE(w_tap_major && wk && cin >= 1 && cin <= kStemCpad && cin_pad >= cin, "stem_pack: Cin must be <= %d", kStemCpad);
  stem_pack_kernel<<<(kStemCout * kStemKpad + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w_tap_major, cin, cin_pad, wk);
  PTK_LAUNCH_CHECK("stem_pack_kernel");
  return 0;
}

extern "C" int ptk_stem_conv(const float* x, int ldx, const float* wk, const float* bias, float* y, int ldy, int N, int H,
                             int W, void* stream) {
  PTK_REQUIRE(x && wk && y && N > 0 && N <= 65535 && H > 0 && W > 0, "stem_conv: bad arguments");
  PTK_REQUIRE(ldx >= kStemCpad && ldx % 4 == 0 && ldy >= kStemCout && ldy % 4 == 0, "stem_conv: ldx >= 24, ldy >= 64, both multiples of 4");
  PTK_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(wk) & 15) == 0, "stem_conv: pointers must be 16-byte aligned");
  const int tiles_x = (W + kTileW - 1) / kTileW, tiles_y = (H + kTileH - 1) / kTileH;
  const size_t smem = 2 * 128 * 128 + kStemChunks * kStemCout * 128 + kHaloH * kHaloW * kHaloStride * 4 + 64 + 1024;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(stem_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
  const int total = tiles_x * tiles_y * N;
  const int ctas = total < 2 * num_sms() ? total : 2 * num_sms();
  stem_conv_kernel<<<ctas, 192, smem, (cudaStream_t)stream>>>(x, ldx, wk, bias, y, ldy, H, W, tiles_x, tiles_x * tiles_y, total);
  PTK_LAUNCH_CHECK("stem_conv_kernel");
  return 0;
}
