// tcgen05 (5th-gen tensor core) TF32 implicit-GEMM convolution for sm_100a.
//
//   D[m, n] = sum_taps sum_c A_tap[m, c] * B_tap[n, c]      (fp32 storage, TF32 operands, fp32 accumulate in TMEM)
//
// * A_tap tile = 128 output-grid pixels x 32 channels, fetched by ONE 4-D TMA box straight out of the NHWC
//   activation tensor (the im2col matrix never exists): box (32 ch, BW, BH, BI images) at the tap's pixel offset,
//   out-of-image taps zero-filled by TMA.  Stride-2 convs use four parity-split tensor maps (one per input
//   row/column parity) so that every box is a dense unit-stride tile.
// * B_tap tile = BLOCK_N output channels x 32 input channels of the K-major packed weight [tap][Cout][Cin].
// * Both land in shared memory in the 128-byte-swizzled K-major layout that tcgen05.mma consumes directly.
// * Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer, warps 2..5 = epilogue
//   (tcgen05.ld -> registers -> global, plus the per-sample sum / sum-of-squares that the following
//   InstanceNorm3d(1) needs, models/networks.py:159).
// * Transposed convs (decoder, and the dgrad of every strided conv) run as 4 output-parity phases
//   (blockIdx.z), each a 2x2-tap stride-1 gather.
#include <cuda.h>
#include "common.cuh"
#include <map>
#include <mutex>
#include "tc_ptx.cuh"

namespace ptk {

struct TcPhase {
  int GH, GW, py, px, ntaps;
  signed char cy[16], cx[16];
  unsigned char wt[16], map[16];
};

struct TcGeom {
  int N, OH, OW, ldy, so;
  int BW, BH, BI, tiles_x, tiles_y, tiles_i;
  int kchunks;
  int nph;         // output-parity phases of the launch (1, or 4 for transposed / dgrad-of-strided gathers)
  int ntile_n;     // output-channel tiles (Cout / BLOCK_N)
                   // Launch order = phase fastest, then output-channel tile, then pixel tile: every CTA that reads one input region
                   // runs at about the same time, so the region comes from HBM once and from L2 for the others (the weights, a
                   // few MB, always sit in L2).  Pixel-tile-fastest order: ncu showed decoder.5 reading its 268 MB input 5x from
                   // DRAM (0.503 -> 0.412 ms) and its dgrad 2x.
  int splits;      // split-K factor: the k-blocks of a tile are divided over `splits` CTAs
  long long part_stride;   // > 0: split z stores its partial tile into y + z * part_stride (summed in a fixed order by
                           // splitk_reduce_kernel => deterministic); 0: the splits accumulate atomically into y
  TcPhase ph[4];
};

struct TmapSet {
  CUtensorMap a[4];
  CUtensorMap b;
};

// MH = number of 128-pixel halves of the CTA's M tile (1 or 2): with MH = 2 the same weight (B) stage feeds two
// M = 128 MMAs, and BLOCK_N = 256 lets one activation (A) stage feed a twice-as-wide MMA -- both raise the MACs per
// byte fetched from L2, which (not the tensor pipe) is what bounds fp32-operand tiles.  TMEM: MH * BLOCK_N columns.
// (A CTA-pair variant that multicast the weight tile into both shared memories halved the L2 -> SM weight traffic but measured
// 3 % slower over the conv stack -- lock-step stage recycling and two cluster barriers per CTA -- and was removed.)
template <int BLOCK_N, int STAGES, int MH>
__global__ void __launch_bounds__(192)
conv_tc_kernel(const __grid_constant__ TmapSet maps, const __grid_constant__ TcGeom g, float* __restrict__ y,
               double* __restrict__ stats, const float* __restrict__ bias, int act) {
  constexpr uint32_t A_BYTES = MH * 128 * 128, B_BYTES = BLOCK_N * 128, TMEM_COLS = MH * BLOCK_N;
  static_assert(TMEM_COLS == 32 || TMEM_COLS == 64 || TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM columns must be a power of two");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + STAGES * A_BYTES, sBar = sB + STAGES * B_BYTES;
  // barriers: full[STAGES], empty[STAGES], tmem_full ; then the TMEM base-address slot
  const uint32_t bar_full = sBar, bar_empty = sBar + 8 * STAGES, bar_tmem = sBar + 16 * STAGES;
  const uint32_t tmem_slot = bar_tmem + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  pdl_trigger();
  // launch order: output-parity phase fastest, then output-channel tile, then pixel tile (TcGeom::nph)
  const TcPhase ph = g.ph[blockIdx.x % g.nph];
  const int split = blockIdx.z;
  const int nt = (blockIdx.x / g.nph) % g.ntile_n;      // output-channel tile
  int t = blockIdx.x / (g.nph * g.ntile_n);
  const int tx = t % g.tiles_x; t /= g.tiles_x;
  const int ty = t % g.tiles_y;
  const int ti = t / g.tiles_y;
  const int gx0 = tx * g.BW, gy0 = ty * g.BH, n0 = ti * g.BI;
  const int KB_all = ph.ntaps * g.kchunks;
  const int kb_per = (KB_all + g.splits - 1) / g.splits;
  const int kb0 = split * kb_per;
  const int KB = (kb0 + kb_per < KB_all ? kb0 + kb_per : KB_all) - kb0;   // k-blocks of this CTA (may be <= 0)

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_tmem, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();      // nothing above touched global memory: the setup overlapped the previous kernel's tail

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES;
        const uint32_t par = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(bar_empty + 8 * s, par ^ 1u);
        const int tap = (kb0 + kb) / g.kchunks;
        const int c0 = ((kb0 + kb) - tap * g.kchunks) * 32;
        mbar_expect_tx(bar_full + 8 * s, A_BYTES + B_BYTES);
        tma_load_4d(sA + s * A_BYTES, &maps.a[ph.map[tap]], bar_full + 8 * s, c0, gx0 + ph.cx[tap], gy0 + ph.cy[tap], n0);
        tma_load_3d(sB + s * B_BYTES, &maps.b, bar_full + 8 * s, c0, nt * BLOCK_N, ph.wt[tap]);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_tf32(128, BLOCK_N);
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES;
        const uint32_t par = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(bar_full + 8 * s, par);
        tc_fence_after();
        const uint64_t da = smem_desc_k_sw128(sA + s * A_BYTES), db = smem_desc_k_sw128(sB + s * B_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k)   // 4 x (K = 8 tf32 = 32 bytes) per 128-byte swizzle row
#pragma unroll
          for (int hm = 0; hm < MH; ++hm)   // pixel rows [128 hm, 128 hm + 128) of the stage -> accumulator hm
            tc_mma_tf32(tmem_base + (uint32_t)(hm * BLOCK_N), da + (uint64_t)(hm * (16384 >> 4)) + 2 * k, db + 2 * k, idesc,
                        (kb > 0 || k > 0) ? 1u : 0u);
        tc_commit(bar_empty + 8 * s);       // frees the smem stage once these MMAs have read it
      }
      if (KB > 0) tc_commit(bar_tmem);  // accumulator complete
    }
    __syncwarp();
  } else if (KB > 0) {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int lg = warp & 3;            // TMEM lane group this warp may access
    mbar_wait(bar_tmem, 0);
    tc_fence_after();
#pragma unroll 1
    for (int hm = 0; hm < MH; ++hm) {
      const int row = hm * 128 + lg * 32 + lane;
      const int x = row % g.BW;
      const int yy = (row / g.BW) % g.BH;
      const int ii = row / (g.BW * g.BH);
      const int gx = gx0 + x, gy = gy0 + yy, n = n0 + ii;
      const bool valid = gx < ph.GW && gy < ph.GH && n < g.N;
      float* dst = nullptr;
      if (valid) {
        const int oy = gy * g.so + ph.py, ox = gx * g.so + ph.px;
        dst = y + (int64_t)split * g.part_stride + (((int64_t)n * g.OH + oy) * g.OW + ox) * g.ldy + nt * BLOCK_N;
      }
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        float v[32];
        tc_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(hm * BLOCK_N + c * 32), v);
        if (valid && g.splits > 1 && g.part_stride == 0) {        // split-K partial sums: accumulate into the zero-filled output
#pragma unroll
          for (int q = 0; q < 32; ++q) atomicAdd(dst + c * 32 + q, v[q]);
        } else if (valid) {
          if (bias != nullptr) {
            const float* bp = bias + nt * BLOCK_N + c * 32;
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] += __ldg(bp + q);
          }
#pragma unroll
          for (int q = 0; q < 32; ++q) { s1 += v[q]; s2 = fmaf(v[q], v[q], s2); }
          if (act != PTK_ACT_NONE) {
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] = apply_act(v[q], act);
          }
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(dst + c * 32 + q * 4) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
      }
      if (stats != nullptr && g.splits == 1) {   // (split-K: the statistics come out of the reduction kernel)
        if (g.BW * g.BH >= 32) {          // the warp's 32 rows belong to one image
          const float a = warp_sum(s1), b = warp_sum(s2);
          const int nw = n0 + (hm * 128 + lg * 32) / (g.BW * g.BH);
          if (lane == 0 && nw < g.N) { atomicAdd(stats + 2 * nw, (double)a); atomicAdd(stats + 2 * nw + 1, (double)b); }
        } else if (valid) {
          atomicAdd(stats + 2 * n, (double)s1);
          atomicAdd(stats + 2 * n + 1, (double)s2);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}


// ----------------------------------------------------------------------------- persistent variant
// Same tile arithmetic as conv_tc_kernel, but a CTA stays resident and walks tiles T = blockIdx.x, + gridDim.x, ... of the
// whole launch (output-parity phase fastest, then output-channel tile, then pixel tile).  The TMA producer and the MMA
// issuer run ahead across tile boundaries (one shared-memory ring for the whole CTA lifetime) and the accumulator is
// DOUBLE-BUFFERED in tensor memory: while the four epilogue warps drain tile i (tcgen05.ld -> bias / statistics /
// activation -> global), the MMAs of tile i + 1 already fill the other buffer.  For the layers with short K loops (stems,
// encoder level 1, the 1x1 head GEMMs, the PatchGAN) the setup (barrier init, TMEM allocation) and the epilogue were
// 30-60 % of a one-tile CTA's life; here they are paid once per SM / hidden behind the next tile.  No split-K.
// CO = coalescing epilogue: the accumulator block is transposed through shared memory so that every store instruction
// writes whole 128-byte lines (a thread owns a ROW of the accumulator; storing it directly touches 32 partial lines per
// instruction).  Costs 19 KB of shared memory: enabled for the tile shapes that are one CTA per SM anyway; the small
// shapes keep two resident CTAs and store directly.
template <int BLOCK_N, int STAGES, int MH, bool CO>
__global__ void __launch_bounds__(192)
conv_tc_persist_kernel(const __grid_constant__ TmapSet maps, const __grid_constant__ TcGeom g, float* __restrict__ y,
                       double* __restrict__ stats, const float* __restrict__ bias, int act, int ntile_n, int nphases) {
  constexpr uint32_t A_BYTES = MH * 128 * 128, B_BYTES = BLOCK_N * 128, ACC_COLS = MH * BLOCK_N, TMEM_COLS = 2 * ACC_COLS;
  static_assert(TMEM_COLS == 64 || TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "two accumulators must fit tensor memory");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + STAGES * A_BYTES, sBar = sB + STAGES * B_BYTES;
  // barriers: full[STAGES], empty[STAGES], acc_full[2], acc_empty[2]; then the TMEM base-address slot
  const uint32_t bar_full = sBar, bar_empty = sBar + 8 * STAGES, bar_accf = sBar + 16 * STAGES, bar_acce = bar_accf + 16;
  const uint32_t tmem_slot = bar_acce + 16;
  // epilogue staging (one 32 x 32 fp32 block per epilogue warp, rows padded to 36 floats: conflict-free 128-bit accesses)
  // and the global element offset of each of the warp's 32 rows (-1: row outside the tensor)
  float* s_stage = reinterpret_cast<float*>(smem_raw + (sBar + 16 * STAGES + 64 - smem_u32(smem_raw)));
  long long* s_rowoff = reinterpret_cast<long long*>(s_stage + 4 * 32 * 36);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(bar_accf + 8 * a, 1); mbar_init(bar_acce + 8 * a, 4); }    // 4 epilogue warps release a buffer
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();

  const int mtiles = g.tiles_x * g.tiles_y * g.tiles_i;
  const int ntiles = mtiles * ntile_n * nphases;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;                                   // k-blocks issued by this CTA so far (ring position)
      for (int T = blockIdx.x; T < ntiles; T += gridDim.x) {
        int t = T / (nphases * ntile_n);                 // phase fastest, then output-channel tile, then pixel tile
        const int nt = (T / nphases) % ntile_n;
        const TcPhase& ph = g.ph[T % nphases];
        const int tx = t % g.tiles_x; t /= g.tiles_x;
        const int ty = t % g.tiles_y;
        const int ti = t / g.tiles_y;
        const int gx0 = tx * g.BW, gy0 = ty * g.BH, n0 = ti * g.BI;
        const int KB = ph.ntaps * g.kchunks;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const uint32_t s = it % STAGES, par = (it / STAGES) & 1u;
          mbar_wait(bar_empty + 8 * s, par ^ 1u);
          const int tap = kb / g.kchunks;
          const int c0 = (kb - tap * g.kchunks) * 32;
          mbar_expect_tx(bar_full + 8 * s, A_BYTES + B_BYTES);
          tma_load_4d(sA + s * A_BYTES, &maps.a[ph.map[tap]], bar_full + 8 * s, c0, gx0 + ph.cx[tap], gy0 + ph.cy[tap], n0);
          tma_load_3d(sB + s * B_BYTES, &maps.b, bar_full + 8 * s, c0, nt * BLOCK_N, ph.wt[tap]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_tf32(128, BLOCK_N);
      uint32_t it = 0, tile_i = 0;
      for (int T = blockIdx.x; T < ntiles; T += gridDim.x, ++tile_i) {
        const int KB = g.ph[T % nphases].ntaps * g.kchunks;
        const uint32_t acc = tile_i & 1u, use = tile_i >> 1;          // use-th time this buffer is filled
        mbar_wait(bar_acce + 8 * acc, (use & 1u) ^ 1u);                // the epilogue has drained its previous contents
        tc_fence_after();
        const uint32_t tacc = tmem_base + acc * ACC_COLS;
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const uint32_t s = it % STAGES, par = (it / STAGES) & 1u;
          mbar_wait(bar_full + 8 * s, par);
          tc_fence_after();
          const uint64_t da = smem_desc_k_sw128(sA + s * A_BYTES), db = smem_desc_k_sw128(sB + s * B_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int hm = 0; hm < MH; ++hm)
              tc_mma_tf32(tacc + (uint32_t)(hm * BLOCK_N), da + (uint64_t)(hm * (16384 >> 4)) + 2 * k, db + 2 * k, idesc,
                          (kb > 0 || k > 0) ? 1u : 0u);
          tc_commit(bar_empty + 8 * s);
        }
        tc_commit(bar_accf + 8 * acc);                                  // accumulator of this tile complete
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int lg = warp & 3;
    uint32_t tile_i = 0;
    for (int T = blockIdx.x; T < ntiles; T += gridDim.x, ++tile_i) {
      int t = T / (nphases * ntile_n);
      const int nt = (T / nphases) % ntile_n;
      const TcPhase& ph = g.ph[T % nphases];
      const int tx = t % g.tiles_x; t /= g.tiles_x;
      const int ty = t % g.tiles_y;
      const int ti = t / g.tiles_y;
      const int gx0 = tx * g.BW, gy0 = ty * g.BH, n0 = ti * g.BI;
      const uint32_t acc = tile_i & 1u, use = tile_i >> 1;
      mbar_wait(bar_accf + 8 * acc, use & 1u);
      tc_fence_after();
      const uint32_t tacc = tmem_base + acc * ACC_COLS;
      float* stg = CO ? s_stage + lg * (32 * 36) : nullptr;
      long long* roff = CO ? s_rowoff + lg * 32 : nullptr;
#pragma unroll 1
      for (int hm = 0; hm < MH; ++hm) {
        const int row = hm * 128 + lg * 32 + lane;
        const int x = row % g.BW;
        const int yy = (row / g.BW) % g.BH;
        const int ii = row / (g.BW * g.BH);
        const int gx = gx0 + x, gy = gy0 + yy, n = n0 + ii;
        const bool valid = gx < ph.GW && gy < ph.GH && n < g.N;
        long long off = -1;
        if (valid) {
          const int oy = gy * g.so + ph.py, ox = gx * g.so + ph.px;
          off = (((long long)n * g.OH + oy) * g.OW + ox) * g.ldy + nt * BLOCK_N;
        }
        if (CO) { __syncwarp(); roff[lane] = off; }
        float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          float v[32];
          tc_ld32(tacc + ((uint32_t)(lg * 32) << 16) + (uint32_t)(hm * BLOCK_N + c * 32), v);
          if (bias != nullptr) {
            const float* bp = bias + nt * BLOCK_N + c * 32;
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] += __ldg(bp + q);
          }
          if (valid) {
#pragma unroll
            for (int q = 0; q < 32; ++q) { s1 += v[q]; s2 = fmaf(v[q], v[q], s2); }
          }
          if (act != PTK_ACT_NONE) {
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] = apply_act(v[q], act);
          }
          if (CO) {
            // 8 lanes per row of 128 contiguous bytes: 4 whole lines per store instruction
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<float4*>(stg + lane * 36 + q * 4) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            __syncwarp();
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {
              const int r = i4 * 4 + (lane >> 3);
              const long long o = roff[r];
              const float4 w4 = *reinterpret_cast<const float4*>(stg + r * 36 + (lane & 7) * 4);
              if (o >= 0) *reinterpret_cast<float4*>(y + o + c * 32 + (lane & 7) * 4) = w4;
            }
          } else if (valid) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<float4*>(y + off + c * 32 + q * 4) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          }
        }
        if (stats != nullptr) {
          if (g.BW * g.BH >= 32) {          // the warp's 32 rows belong to one image
            const float a = warp_sum(s1), b = warp_sum(s2);
            const int nw = n0 + (hm * 128 + lg * 32) / (g.BW * g.BH);
            if (lane == 0 && nw < g.N) { atomicAdd(stats + 2 * nw, (double)a); atomicAdd(stats + 2 * nw + 1, (double)b); }
          } else if (valid) {
            atomicAdd(stats + 2 * n, (double)s1);
            atomicAdd(stats + 2 * n + 1, (double)s2);
          }
        }
      }
      // this warp has read everything it needs from the buffer: one arrival per warp hands it back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_acce + 8 * acc) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// Deterministic split-K combine: y[i] = sum_{z < nparts} parts[z * part_stride + i] (fixed order), fused with the per-sample
// {sum, sum of squares} that the following norm needs.  grid (chunks, N); per_sample = OH * OW * Cout (multiple of 4).
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ parts, int nparts, long long part_stride, float* __restrict__ y, long long per_sample,
                     double* __restrict__ stats) {
  pdl_trigger();
  const int n = blockIdx.y;
  const float4* src = reinterpret_cast<const float4*>(parts + (long long)n * per_sample);
  float4* dst = reinterpret_cast<float4*>(y + (long long)n * per_sample);
  const long long n4 = per_sample >> 2, ps4 = part_stride >> 2;
  float s1 = 0.f, s2 = 0.f;
  double d1 = 0.0, d2 = 0.0;
  int cnt = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 a = __ldg(src + i);
    for (int z = 1; z < nparts; ++z) {
      const float4 b = __ldg(src + z * ps4 + i);
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    dst[i] = a;
    s1 += (a.x + a.y) + (a.z + a.w);
    s2 += (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w);
    if (++cnt == 32) { d1 += s1; d2 += s2; s1 = 0.f; s2 = 0.f; cnt = 0; }
  }
  if (stats != nullptr) {
    d1 += s1; d2 += s2;
    __shared__ double ra[8], rb[8];
    d1 = warp_sum(d1); d2 = warp_sum(d2);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { ra[wid] = d1; rb[wid] = d2; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0.0, b = 0.0;
      for (int q = 0; q < 8; ++q) { a += ra[q]; b += rb[q]; }
      atomicAdd(stats + 2 * n, a);
      atomicAdd(stats + 2 * n + 1, b);
    }
  }
}

// ----------------------------------------------------------------------------- host side
static bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("PTK_PDL"); v = (e && atoi(e) == 0) ? 0 : 1; }
  return v != 0;
}

// launch with programmatic stream serialization (see pdl_wait / pdl_trigger in tc_ptx.cuh)
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// Tensor maps are pure functions of (base pointer, extents, strides, box, swizzle) and the step re-uses the same buffers
// every iteration: descriptors are encoded once and looked up afterwards (a step needs ~500 of them).
struct TmapKey {
  uint64_t v[14];
  bool operator<(const TmapKey& o) const { return memcmp(v, o.v, sizeof(v)) < 0; }
};
static std::map<TmapKey, CUtensorMap>& tmap_cache() { static std::map<TmapKey, CUtensorMap> c; return c; }
static std::mutex& tmap_mutex() { static std::mutex m; return m; }

static int encode_uncached(CUtensorMap* m, const float* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                           const uint32_t* box, CUtensorMapSwizzle swizzle);

static int encode(CUtensorMap* m, const float* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  TmapKey k;
  memset(&k, 0, sizeof(k));
  k.v[0] = (uint64_t)reinterpret_cast<uintptr_t>(base);
  k.v[1] = (uint64_t)rank | ((uint64_t)swizzle << 8);
  for (int i = 0; i < rank; ++i) { k.v[2 + i] = dims[i]; k.v[10 + i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) k.v[6 + i] = strides_bytes[i];
  {
    std::lock_guard<std::mutex> lock(tmap_mutex());
    auto it = tmap_cache().find(k);
    if (it != tmap_cache().end()) { *m = it->second; return 0; }
  }
  int rc = encode_uncached(m, base, rank, dims, strides_bytes, box, swizzle);
  if (rc) return rc;
  std::lock_guard<std::mutex> lock(tmap_mutex());
  if (tmap_cache().size() > 16384) tmap_cache().clear();      // (buffers of a long-gone shape: start over)
  tmap_cache()[k] = *m;
  return 0;
}

static int encode_uncached(CUtensorMap* m, const float* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                           const uint32_t* box, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(5, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, (void*)base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(5, "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu %llu box %u %u %u %u", (int)r, rank,
                (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], box[1], box[2], rank > 3 ? box[3] : 0);
  return 0;
}

static int pow2_ge(int v) { int p = 1; while (p < v) p <<= 1; return p; }
static int floordiv2(int q) { return q >= 0 ? q / 2 : -((-q + 1) / 2); }

// ---------------------------------------------------------------- first-use autotuning of the tile shape
// Keyed by the layer geometry (and the call flavour); the winner of a few timed launches on the real operands is kept for
// the life of the process.  PTK_TC_AUTOTUNE=0 keeps the static cost model (bit-reproducible tile choice across runs).
struct TuneKey {
  int v[20];
  bool operator<(const TuneKey& o) const { return memcmp(v, o.v, sizeof(v)) < 0; }
};
struct TuneChoice { int mh, bn, splits, tpc, persist; };

static TuneKey tune_key(const ptk_conv_geom& c, int kind, int flags, int64_t capacity) {
  TuneKey k;
  memset(&k, 0, sizeof(k));
  const int f[] = {c.N, c.H, c.W, c.Cin, c.ldx, c.OH, c.OW, c.Cout, c.ldy, c.k, c.stride, c.pad, c.transposed, kind, flags,
                   (int)(capacity >> 20)};
  for (size_t i = 0; i < sizeof(f) / sizeof(f[0]); ++i) k.v[i] = f[i];
  return k;
}

static std::map<TuneKey, TuneChoice>& tune_table() { static std::map<TuneKey, TuneChoice> t; return t; }
static std::mutex& tune_mutex() { static std::mutex m; return m; }

// PTK_TC_TUNE_FILE=<path>: the table is read from that file at first use and every new winner is appended to it, so that a
// later process (a profiler run, a production job) makes exactly the tile choices of the process that tuned.
static void tune_load_locked() {
  static bool loaded = false;
  if (loaded) return;
  loaded = true;
  const char* path = getenv("PTK_TC_TUNE_FILE");
  if (!path) return;
  FILE* f = fopen(path, "r");
  if (!f) return;
  for (;;) {
    TuneKey k;
    TuneChoice c;
    bool ok = true;
    for (int i = 0; i < 20 && ok; ++i) ok = fscanf(f, "%d", &k.v[i]) == 1;
    ok = ok && fscanf(f, "%d %d %d %d %d", &c.mh, &c.bn, &c.splits, &c.tpc, &c.persist) == 5;
    if (!ok) break;
    tune_table()[k] = c;
  }
  fclose(f);
}

static bool tune_lookup(const TuneKey& k, TuneChoice* out) {
  std::lock_guard<std::mutex> lock(tune_mutex());
  tune_load_locked();
  auto it = tune_table().find(k);
  if (it == tune_table().end()) return false;
  *out = it->second;
  return true;
}

static void tune_store(const TuneKey& k, const TuneChoice& c) {
  std::lock_guard<std::mutex> lock(tune_mutex());
  tune_load_locked();
  tune_table()[k] = c;
  if (const char* path = getenv("PTK_TC_TUNE_FILE")) {
    if (FILE* f = fopen(path, "a")) {
      for (int i = 0; i < 20; ++i) fprintf(f, "%d ", k.v[i]);
      fprintf(f, "%d %d %d %d %d\n", c.mh, c.bn, c.splits, c.tpc, c.persist);
      fclose(f);
    }
  }
}

static bool tc_autotune_enabled(cudaStream_t st) {
  const char* e = getenv("PTK_TC_AUTOTUNE");         // (read per call: tests switch it inside one process)
  if (e && atoi(e) == 0) return false;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return false;
  return true;
}

// one warm-up + three individually timed launches of `fn` on stream st, minimum taken (drains the device first, blocks the
// host until they finish)
template <typename F>
static int tune_time(cudaStream_t st, F fn, float* ms) {
  cudaEvent_t ev[4];
  for (int i = 0; i < 4; ++i)
    if (cudaEventCreate(&ev[i]) != cudaSuccess) return fail(3, "autotune: cudaEventCreate failed");
  cudaDeviceSynchronize();            // the trial runs alone: work queued on other streams would distort the comparison
  int rc = fn();
  for (int i = 0; i < 3 && !rc; ++i) {
    if (i == 0) cudaEventRecord(ev[0], st);
    rc = fn();
    cudaEventRecord(ev[i + 1], st);
  }
  if (!rc) {
    if (cudaEventSynchronize(ev[3]) != cudaSuccess) rc = fail(3, "autotune: %s", cudaGetErrorString(cudaGetLastError()));
    else {
      float best = 0.f;
      for (int i = 0; i < 3; ++i) {
        float t = 0.f;
        cudaEventElapsedTime(&t, ev[i], ev[i + 1]);
        if (i == 0 || t < best) best = t;
      }
      *ms = best;
    }
  }
  for (int i = 0; i < 4; ++i) cudaEventDestroy(ev[i]);
  return rc;
}

bool conv_tc_supported(const ptk_conv_geom& c) {
  if (c.Cin % 32 != 0 || c.Cout % 32 != 0) return false;
  if (c.ldx % 4 != 0 || c.ldy % 4 != 0) return false;
  if (!((c.k == 4 && c.stride == 2) || (c.k == 3 && c.stride == 1) || (c.k == 1 && c.stride == 1 && c.pad == 0 && !c.transposed))) return false;
  // (1- and 2-pixel extents work on this path -- verified against the CUDA-core kernels -- but stay on the fp32 kernels: at
  //  64x64 / 128x64 inputs the bottleneck norms average over a few thousand elements only, and TF32 there multiplies the
  //  rounding noise of the cancellation-dominated scalar gradients)
  if (c.H < 4 || c.W < 4 || c.OH < 4 || c.OW < 4) return false;
  if (c.N < 1 || c.N > 4096) return false;
  return true;
}

int conv_forward_tc(const ptk_conv_geom& c, const float* x, const float* w_k, const float* bias, int act, float* y,
                    double* stats, float* scratch, int64_t scratch_floats, cudaStream_t st) {
  PTK_REQUIRE(act == PTK_ACT_NONE || act == PTK_ACT_LEAKY || act == PTK_ACT_RELU, "conv_forward(tc): unsupported activation epilogue");
  PTK_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(w_k) & 15) == 0, "conv_forward(tc): pointers must be 16-byte aligned");
  TcGeom g;
  memset(&g, 0, sizeof(g));
  g.N = c.N; g.OH = c.OH; g.OW = c.OW; g.ldy = c.ldy; g.kchunks = c.Cin / 32;
  int nphases, maxGH = 0, maxGW = 0;
  const int k = c.k, s = c.stride;
  if (!c.transposed) {
    nphases = 1; g.so = 1;
    TcPhase& p = g.ph[0];
    p.GH = c.OH; p.GW = c.OW; p.py = p.px = 0; p.ntaps = 0;
    for (int kh = 0; kh < k; ++kh)
      for (int kw = 0; kw < k; ++kw) {
        const int t = p.ntaps++;
        const int qy = kh - c.pad, qx = kw - c.pad;
        if (s == 2) {
          p.map[t] = (unsigned char)((((qy % 2) + 2) % 2) * 2 + (((qx % 2) + 2) % 2));
          p.cy[t] = (signed char)floordiv2(qy); p.cx[t] = (signed char)floordiv2(qx);
        } else {
          p.map[t] = 0; p.cy[t] = (signed char)qy; p.cx[t] = (signed char)qx;
        }
        p.wt[t] = (unsigned char)(kh * k + kw);
      }
  } else {
    nphases = s * s; g.so = s;
    for (int py = 0; py < s; ++py)
      for (int px = 0; px < s; ++px) {
        TcPhase& p = g.ph[py * s + px];
        p.GH = (c.OH - py + s - 1) / s; p.GW = (c.OW - px + s - 1) / s; p.py = py; p.px = px; p.ntaps = 0;
        for (int kh = 0; kh < k; ++kh) {
          if ((py + c.pad - kh) % s != 0) continue;
          for (int kw = 0; kw < k; ++kw) {
            if ((px + c.pad - kw) % s != 0) continue;
            const int t = p.ntaps++;
            p.map[t] = 0;
            p.cy[t] = (signed char)((py + c.pad - kh) / s); p.cx[t] = (signed char)((px + c.pad - kw) / s);
            p.wt[t] = (unsigned char)(kh * k + kw);
          }
        }
      }
  }
  for (int i = 0; i < nphases; ++i) { maxGH = g.ph[i].GH > maxGH ? g.ph[i].GH : maxGH; maxGW = g.ph[i].GW > maxGW ? g.ph[i].GW : maxGW; }
  int min_kb = 1 << 30;
  for (int i = 0; i < nphases; ++i) { const int kb = g.ph[i].ntaps * g.kchunks; if (kb < min_kb) min_kb = kb; }

  // ---- tile selection.  fp32 operands make every tile L2-bandwidth bound (bytes per MAC ~ 1/M + 1/N), so the largest
  // tile that still fills the machine wins: (MH = 2, BN = 256) moves half the bytes per MAC of (1, 128).  Cost model:
  // waves x per-tile time, with a relative efficiency per tile shape and a fixed prologue/epilogue charge for the
  // one-CTA-per-SM shapes (nothing overlaps them) that grows with the tile's output.  PTK_TC_TILE="mh,bn" overrides (tests /
  // experiments).  Calibrated against tools/bench_conv.py on B200 (gpurun_out r2n): with the coalescing epilogue (2, 64)
  // beats (1, 64) on every Cout = 64 layer (stems 0.090 -> 0.067 ms); a split-K pass costs a streamed read of every
  // partial buffer plus a launch, which makes (1, 256) un-split the better choice on the 64^2 / 32^2 encoder layers.
  struct TileCfg { int mh, bn, stages, occ; float eff; };
  static const TileCfg kCfgs[] = {{1, 32, 4, 2, 0.20f}, {2, 32, 3, 2, 0.25f}, {1, 64, 4, 2, 0.36f}, {2, 64, 4, 1, 0.42f}, {1, 128, 3, 2, 0.50f},
                                  {1, 256, 4, 1, 0.64f}, {2, 128, 4, 1, 0.64f}, {2, 256, 3, 1, 0.80f}};
  int forced_mh = 0, forced_bn = 0;
  if (const char* e = getenv("PTK_TC_TILE")) sscanf(e, "%d,%d", &forced_mh, &forced_bn);
  struct Cand { const TileCfg* cfg; int splits; double cost; int persist; };     // persist: 1 = if the shape allows, 0 = never
  Cand cands[40];
  int ncand = 0;
  const Cand* forced_cand = nullptr;
  const bool can_split = c.ldy == c.Cout && bias == nullptr && act == PTK_ACT_NONE;
  const int64_t out_floats = (int64_t)c.N * c.OH * c.OW * c.Cout;
  const bool use_parts = scratch != nullptr && scratch_floats >= 2 * out_floats && (reinterpret_cast<uintptr_t>(scratch) & 15) == 0 &&
                         out_floats % 4 == 0;
  for (const TileCfg& t : kCfgs) {
    if (c.Cout % t.bn != 0) continue;
    if (t.bn == 64 && c.Cout % 128 == 0) continue;
    if (t.bn == 32 && c.Cout % 64 == 0) continue;
    const int M = 128 * t.mh;
    const int bw = pow2_ge(maxGW < M ? maxGW : M);
    const int bh = pow2_ge(maxGH < M / bw ? maxGH : M / bw);
    const int bi = M / (bw * bh);
    if (bi > 256) continue;
    const bool forced = forced_mh == t.mh && forced_bn == t.bn;
    if (t.mh > 1 && !forced && bi > pow2_ge(c.N)) continue;        // do not pad the batch beyond the next power of two
    const int64_t ctas = (int64_t)((maxGW + bw - 1) / bw) * ((maxGH + bh - 1) / bh) * ((c.N + bi - 1) / bi) * (c.Cout / t.bn) * nphases;
    const int64_t slots = (int64_t)num_sms() * t.occ;
    // Layers with fewer tiles than CTA slots split the K loop so that every slot of ONE wave is busy (traffic per MAC
    // is set by the tile shape, not by the split; the partial sums are combined with atomics).
    int sp = 1;
    if (can_split && ctas * 2 <= slots) {
      sp = (int)(slots / ctas);
      if (sp > min_kb / 4) sp = min_kb / 4;
      if (use_parts && (int64_t)sp * out_floats > scratch_floats) sp = (int)(scratch_floats / out_floats);
      if (sp < 1) sp = 1;
      while (sp > 1 && (min_kb + sp - 1) / sp * (sp - 1) >= min_kb) --sp;
    }
    if (t.occ == 1 && !forced && ctas * sp < slots / 2) continue;   // big one-CTA-per-SM shapes must fill the machine
    // (2, 64) pays off through its coalescing epilogue: only on short K loops (stems, Cout = 64 dgrads / ConvT); with 32+
    // k-blocks per tile two co-resident (1, 64) CTAs hide the N = 64 MMA latency better (PatchGAN stem: 0.104 vs 0.111 ms)
    const double eff = (t.mh == 2 && t.bn == 64 && min_kb > 16) ? 0.30 : t.eff;
    const double tile_clk = (double)((min_kb + sp - 1) / sp) * t.mh * (t.bn / 128.0) * 256.0 / eff +
                            (t.occ == 1 ? 3000.0 * t.mh * (t.bn / 128.0) : 1500.0);
    // combining the splits: fp32 L2 atomics (~100 floats / clk chip-wide), or one streamed pass that reads every partial
    // buffer and writes the result (~460 floats / clk at 3.5 TB/s) behind one more launch
    const double combine = sp > 1 ? (use_parts ? (double)out_floats * (sp + 1) / 460.0 + 5000.0 : (double)out_floats * sp / 100.0) : 0.0;
    const double cost = (double)((ctas * sp + slots - 1) / slots) * t.occ * tile_clk + combine;
    cands[ncand] = Cand{&t, sp, cost, 1};
    if (forced && !forced_cand) forced_cand = &cands[ncand];
    ++ncand;
    // variants for the first-use autotuner only (rated slightly behind the model's own choice): half / double the split
    // count, and the one-tile-per-CTA kernel where the persistent one would be picked
    auto valid_split = [&](int q) {
      if (q < 1 || q == sp || !can_split || q > min_kb / 4) return false;
      if (use_parts && (int64_t)q * out_floats > scratch_floats) return false;
      if (q > 1 && (min_kb + q - 1) / q * (q - 1) >= min_kb) return false;
      return true;
    };
    if (valid_split(sp / 2)) cands[ncand++] = Cand{&t, sp / 2, cost * 1.05, 1};
    if (sp > 1 && valid_split(1)) cands[ncand++] = Cand{&t, 1, cost * 1.10, 1};
    if (ctas * 2 * sp <= 4 * slots && valid_split(sp * 2)) cands[ncand++] = Cand{&t, sp * 2, cost * 1.05, 1};
    if (sp == 1 && 2 * t.mh * t.bn <= 512 && ctas > slots) cands[ncand++] = Cand{&t, 1, cost * 1.08, 0};
  }
  PTK_REQUIRE(ncand > 0, "conv_forward(tc): no tile configuration for Cout=%d", c.Cout);
  for (int i = 1; i < ncand; ++i)            // by model cost (stable: ties keep the table order)
    for (int j = i; j > 0 && cands[j].cost < cands[j - 1].cost; --j) { const Cand tmp = cands[j]; cands[j] = cands[j - 1]; cands[j - 1] = tmp; }
  if (forced_cand) {                         // (the pointer was taken before the sort: find the model's variant of that shape again)
    for (int i = 0; i < ncand; ++i)
      if (cands[i].cfg->mh == forced_mh && cands[i].cfg->bn == forced_bn && cands[i].persist == 1) { forced_cand = &cands[i]; break; }
  }

  // ---- everything that depends on the chosen tile: tensor maps, grid, launch (+ the split-K reduce)
  const TcGeom g0 = g;
  auto run = [&](const TileCfg* best, int best_splits, int allow_persist, double* stats) -> int {
  TcGeom g = g0;
  TmapSet maps;
  memset(&maps, 0, sizeof(maps));
  const int MH = best->mh, BN = best->bn, MT = 128 * MH;
  g.BW = pow2_ge(maxGW < MT ? maxGW : MT);
  g.BH = pow2_ge(maxGH < MT / g.BW ? maxGH : MT / g.BW);
  g.BI = MT / (g.BW * g.BH);
  g.tiles_x = (maxGW + g.BW - 1) / g.BW; g.tiles_y = (maxGH + g.BH - 1) / g.BH; g.tiles_i = (c.N + g.BI - 1) / g.BI;

  // activation tensor maps
  const uint32_t boxA[4] = {32u, (uint32_t)g.BW, (uint32_t)g.BH, (uint32_t)g.BI};
  if (!c.transposed && s == 2) {
    for (int pyy = 0; pyy < 2; ++pyy)
      for (int pxx = 0; pxx < 2; ++pxx) {
        const uint64_t Hp = (uint64_t)(c.H - pyy + 1) / 2, Wp = (uint64_t)(c.W - pxx + 1) / 2;
        const uint64_t dims[4] = {(uint64_t)c.Cin, Wp, Hp, (uint64_t)c.N};
        const uint64_t str[3] = {(uint64_t)2 * c.ldx * 4, (uint64_t)2 * c.W * c.ldx * 4, (uint64_t)c.H * c.W * c.ldx * 4};
        int rc = encode(&maps.a[pyy * 2 + pxx], x + ((int64_t)pyy * c.W + pxx) * c.ldx, 4, dims, str, boxA);
        if (rc) return rc;
      }
  } else {
    const uint64_t dims[4] = {(uint64_t)c.Cin, (uint64_t)c.W, (uint64_t)c.H, (uint64_t)c.N};
    const uint64_t str[3] = {(uint64_t)c.ldx * 4, (uint64_t)c.W * c.ldx * 4, (uint64_t)c.H * c.W * c.ldx * 4};
    int rc = encode(&maps.a[0], x, 4, dims, str, boxA);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)c.Cin, (uint64_t)c.Cout, (uint64_t)(k * k)};
    const uint64_t str[2] = {(uint64_t)c.Cin * 4, (uint64_t)c.Cin * c.Cout * 4};
    const uint32_t boxB[3] = {32u, (uint32_t)BN, 1u};
    int rc = encode(&maps.b, w_k, 3, dims, str, boxB);
    if (rc) return rc;
  }
  const int splits = best_splits;
  g.splits = splits;
  const bool parts = splits > 1 && use_parts;
  g.part_stride = parts ? out_floats : 0;
  float* y_kernel = parts ? scratch : y;
  if (splits > 1 && !parts) {
    int rc = ptk_fill(y, out_floats, 0.f, st);
    if (rc) return rc;
  }
  const int mtiles = g.tiles_x * g.tiles_y * g.tiles_i;
  g.nph = nphases;
  g.ntile_n = c.Cout / BN;
  dim3 grid((unsigned)(mtiles * nphases * g.ntile_n), 1u, (unsigned)splits);
  // Persistent CTAs with a double-buffered accumulator (conv_tc_persist_kernel) when every resident CTA gets several tiles:
  // PTK_TC_PERSIST=0 disables, =2 forces it whenever the tile shape allows (tests), =k >= 3 requires more than (k - 2) tiles
  // per resident CTA (default: more tiles than resident CTAs).
  int ps_env = 1;
  if (const char* e = getenv("PTK_TC_PERSIST")) ps_env = atoi(e);      // (read per call: the tests switch it inside one process)
  const int64_t all_tiles = (int64_t)mtiles * (c.Cout / BN) * nphases;
  const int occ_ps = (size_t)best->stages * (MH * 128 * 128 + BN * 128) > 100 * 1024 ? 1 : 2;
  const int64_t slots_ps = (int64_t)num_sms() * occ_ps;
  const bool persist = allow_persist && splits == 1 && 2 * MH * BN <= 512 && all_tiles < (1 << 30) &&
                       (ps_env == 2 ? all_tiles >= 2 : (ps_env >= 1 && all_tiles > (int64_t)(ps_env >= 3 ? ps_env - 2 : 1) * slots_ps));
  const unsigned grid_ps = (unsigned)(ps_env == 2 ? (all_tiles + 1) / 2 : (all_tiles < slots_ps ? all_tiles : slots_ps));
#define PTK_TC_LAUNCH(BN_, ST_, MH_)                                                                                       \
  do {                                                                                                                     \
    const size_t smem = (size_t)ST_ * (MH_ * 128 * 128 + BN_ * 128) + 16 * ST_ + 16 + 1024;                                 \
    static bool attr = false;                                                                                              \
    if (!attr) { cudaFuncSetAttribute(conv_tc_kernel<BN_, ST_, MH_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; } \
    launch_pdl(conv_tc_kernel<BN_, ST_, MH_>, grid, dim3(192), smem, st, maps, g, y_kernel, stats, bias, act);            \
  } while (0)
#define PTK_TC_LAUNCH_PS(BN_, ST_, MH_, CO_)                                                                               \
  do {                                                                                                                     \
    const size_t smem = (size_t)ST_ * (MH_ * 128 * 128 + BN_ * 128) + 16 * ST_ + 64 + (CO_ ? 18432 + 1024 : 0) + 1024 + 64;  \
    static bool attr = false;                                                                                              \
    if (!attr) { cudaFuncSetAttribute(conv_tc_persist_kernel<BN_, ST_, MH_, CO_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; } \
    launch_pdl(conv_tc_persist_kernel<BN_, ST_, MH_, CO_>, dim3(grid_ps), dim3(192), smem, st, maps, g, y_kernel, stats, bias, act, \
               c.Cout / BN_, nphases);                                                                                     \
  } while (0)
  if (persist) {
    if (MH == 1 && BN == 32) PTK_TC_LAUNCH_PS(32, 4, 1, false);
    else if (MH == 2 && BN == 32) PTK_TC_LAUNCH_PS(32, 3, 2, false);
    else if (MH == 1 && BN == 64) PTK_TC_LAUNCH_PS(64, 4, 1, false);
    else if (MH == 2 && BN == 64) PTK_TC_LAUNCH_PS(64, 4, 2, true);
    else if (MH == 1 && BN == 128) PTK_TC_LAUNCH_PS(128, 3, 1, false);
    else if (MH == 1 && BN == 256) PTK_TC_LAUNCH_PS(256, 4, 1, true);
    else PTK_TC_LAUNCH_PS(128, 4, 2, true);
    PTK_LAUNCH_CHECK("conv_tc_persist_kernel");
    return 0;
  }
#undef PTK_TC_LAUNCH_PS
  if (MH == 1 && BN == 32) PTK_TC_LAUNCH(32, 4, 1);
  else if (MH == 2 && BN == 32) PTK_TC_LAUNCH(32, 3, 2);
  else if (MH == 1 && BN == 64) PTK_TC_LAUNCH(64, 4, 1);
  else if (MH == 2 && BN == 64) PTK_TC_LAUNCH(64, 4, 2);
  else if (MH == 1 && BN == 128) PTK_TC_LAUNCH(128, 3, 1);
  else if (MH == 1 && BN == 256) PTK_TC_LAUNCH(256, 4, 1);
  else if (MH == 2 && BN == 128) PTK_TC_LAUNCH(128, 4, 2);
  else PTK_TC_LAUNCH(256, 3, 2);
#undef PTK_TC_LAUNCH
  PTK_LAUNCH_CHECK("conv_tc_kernel");
  if (parts) {
    const long long per_sample = (long long)c.OH * c.OW * c.Cout;
    int blocks = (int)((per_sample / 4 + 255) / 256);
    const int cap = (num_sms() * 8 + c.N - 1) / c.N;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    splitk_reduce_kernel<<<dim3((unsigned)blocks, (unsigned)c.N), 256, 0, st>>>(scratch, splits, (long long)out_floats, y, per_sample, stats);
    PTK_LAUNCH_CHECK("splitk_reduce_kernel");
    return 0;
  }
  if (splits > 1 && stats != nullptr)   // atomic split-K: the statistics need one extra pass over a tiny tensor
    return ptk_gn_stats(y, c.ldy, c.N, (int64_t)c.OH * c.OW, c.Cout, stats, st);
  return 0;
  };   // run

  if (forced_cand) return run(forced_cand->cfg, forced_cand->splits, 1, stats);
  const Cand* pick = &cands[0];
  for (int i = 0; i < ncand; ++i) if (cands[i].persist == 1 && cands[i].cost <= pick->cost) { pick = &cands[i]; break; }   // model's choice
  if (ncand > 1 && tc_autotune_enabled(st)) {
    // First use of this layer geometry: time the model's best candidates on the real operands and remember the winner
    // (the cost model ranks tile shapes within ~15 %; wave quantisation, persistence and the ragged edges decide the rest).
    const TuneKey key = tune_key(c, 0, (bias != nullptr) * 2 + (use_parts ? 1 : 0) + act * 4, scratch_floats);
    TuneChoice ch;
    if (!tune_lookup(key, &ch)) {
      const int ntry = ncand < 12 ? ncand : 12;
      float best_ms = 0.f;
      int best_i = 0;
      for (int i = 0; i < ntry; ++i) {
        float ms = 0.f;
        int rc = tune_time(st, [&]() { return run(cands[i].cfg, cands[i].splits, cands[i].persist, nullptr); }, &ms);
        if (rc) return rc;
        if (i == 0 || ms < best_ms) { best_ms = ms; best_i = i; }
      }
      ch.mh = cands[best_i].cfg->mh; ch.bn = cands[best_i].cfg->bn; ch.splits = cands[best_i].splits; ch.tpc = 1;
      ch.persist = cands[best_i].persist;
      tune_store(key, ch);
    }
    for (int i = 0; i < ncand; ++i)
      if (cands[i].cfg->mh == ch.mh && cands[i].cfg->bn == ch.bn && cands[i].splits == ch.splits && cands[i].persist == ch.persist) { pick = &cands[i]; break; }
  }
  return run(pick->cfg, pick->splits, pick->persist, stats);
}


// ============================================================================= weight gradient on tcgen05
//   dW[t][a][b] = sum_pixels S[m][a] * Bg[2*m + q(t)][b]     (S: small grid, Bg: stride-2 addressed tensor)
// Both operands are "MN-major" for the tensor core (channels contiguous, the GEMM-K index = pixel is the row).  For
// 32-bit MN-major operands tcgen05 requires the "128B swizzle with 32B atoms" layout (UMMA layout type 1; TMA
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): a TMA box (32 channels x 32 pixels) lands as a column of 32x4 atoms
// (4-pixel groups 512 B apart = SBO, next 32 channels = next box = LBO).  Split-K over pixel tiles across CTAs.
struct WgTcGeom {
  int N, GH, GW;                 // small grid
  int ntaps;
  int BW, BH, BI, tiles_x, tiles_y, tiles_i, ntiles;
  int Ca, Cb_pad;                // dw[t][Ca][Cb_pad]
  int splits;
  long long part_stride;         // > 0: split z writes its own partial buffer dw + z * part_stride (plain stores, summed
                                 // by ptk_unpack_weight_grad_parts => deterministic); 0: splits accumulate atomically
  signed char cy[16], cx[16];
  unsigned char map[16];
};

struct WgTmapSet {
  CUtensorMap s;       // small tensor
  CUtensorMap b[4];    // big tensor, one per (row parity, col parity)
};

__device__ __forceinline__ uint64_t smem_desc_mn_sw128_32b(uint32_t addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;   // stride between 32-channel (128 B) groups along M/N
  d |= (uint64_t)(512 >> 4) << 32;                     // stride between 4-pixel groups along K
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                              // SWIZZLE_128B_BASE32B
  return d;
}

// MH = number of 128-channel halves of the A (Ca) tile: MH = 2 / BLOCK_N = 256 halve the L2 bytes per MAC (see
// conv_tc_kernel).  TPC = taps per CTA (MH == 1 only): narrow layers (Ca <= 128, Cb <= 64: stems, encoder level 1, the
// PatchGAN stem) let one staged gradient tile S feed TPC taps' worth of MMAs, each tap with its own shifted activation
// tile and its own TMEM accumulator -- 2-3.5x fewer L2 bytes per MAC than one tap per CTA.  TMEM: TPC * MH * BLOCK_N cols.
template <int BLOCK_N, int STAGES, int MH, int TPC>
__global__ void __launch_bounds__(192)
wgrad_tc_kernel(const __grid_constant__ WgTmapSet maps, const __grid_constant__ WgTcGeom g, float* __restrict__ dw) {
  static_assert(TPC == 1 || MH == 1, "taps-per-CTA > 1 needs a single 128-channel A half");
  constexpr uint32_t KP = 32;                                    // pixels per stage
  constexpr uint32_t A_BYTES = MH * 128 * KP * 4, B_BYTES = BLOCK_N * KP * 4, BOX_BYTES = 32 * KP * 4;
  constexpr uint32_t NEED_COLS = TPC * MH * BLOCK_N;
  constexpr uint32_t TMEM_COLS = NEED_COLS <= 32 ? 32 : NEED_COLS <= 64 ? 64 : NEED_COLS <= 128 ? 128 : NEED_COLS <= 256 ? 256 : 512;
  static_assert(NEED_COLS <= 512, "accumulators exceed TMEM");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + STAGES * A_BYTES, sBar = sB + STAGES * TPC * B_BYTES;
  const uint32_t bar_full = sBar, bar_empty = sBar + 8 * STAGES, bar_tmem = sBar + 16 * STAGES;
  const uint32_t tmem_slot = bar_tmem + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  const int btiles = g.Cb_pad / BLOCK_N;
  const int at = blockIdx.x / btiles, bt = blockIdx.x - at * btiles;
  const int tap0 = blockIdx.y * TPC, split = blockIdx.z;
  const int ntap = (g.ntaps - tap0 < TPC) ? g.ntaps - tap0 : TPC;        // taps of this CTA
  const int per = (g.ntiles + g.splits - 1) / g.splits;
  const int t_begin = split * per;
  const int t_end = t_begin + per < g.ntiles ? t_begin + per : g.ntiles;
  const int KB = t_end > t_begin ? t_end - t_begin : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_tmem, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES;
        const uint32_t par = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(bar_empty + 8 * s, par ^ 1u);
        int t = t_begin + kb;
        const int tx = t % g.tiles_x; t /= g.tiles_x;
        const int ty = t % g.tiles_y;
        const int ti = t / g.tiles_y;
        const int gx0 = tx * g.BW, gy0 = ty * g.BH, n0 = ti * g.BI;
        mbar_expect_tx(bar_full + 8 * s, A_BYTES + (uint32_t)ntap * B_BYTES);
#pragma unroll
        for (int j = 0; j < 4 * MH; ++j)
          tma_load_4d(sA + s * A_BYTES + j * BOX_BYTES, &maps.s, bar_full + 8 * s, at * (128 * MH) + j * 32, gx0, gy0, n0);
#pragma unroll 1
        for (int q = 0; q < ntap; ++q) {
          const int tap = tap0 + q;
          const CUtensorMap* mb = &maps.b[g.map[tap]];
#pragma unroll
          for (int j = 0; j < BLOCK_N / 32; ++j)
            tma_load_4d(sB + (s * TPC + q) * B_BYTES + j * BOX_BYTES, mb, bar_full + 8 * s, bt * BLOCK_N + j * 32, gx0 + g.cx[tap],
                        gy0 + g.cy[tap], n0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // both operands MN-major: bits 15 (A) and 16 (B) of the instruction descriptor
      constexpr uint32_t idesc = idesc_tf32(128, BLOCK_N) | (1u << 15) | (1u << 16);
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % STAGES;
        const uint32_t par = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(bar_full + 8 * s, par);
        tc_fence_after();
        const uint64_t da = smem_desc_mn_sw128_32b(sA + s * A_BYTES, BOX_BYTES);
#pragma unroll 1
        for (int q = 0; q < ntap; ++q) {
          const uint64_t db = smem_desc_mn_sw128_32b(sB + (s * TPC + q) * B_BYTES, BOX_BYTES);
#pragma unroll
          for (int k = 0; k < (int)KP / 8; ++k)   // 8 pixels (one 1024-byte swizzle atom row-group) per MMA
#pragma unroll
            for (int hm = 0; hm < MH; ++hm)       // channels [128 hm, 128 hm + 128) of the A stage = boxes 4 hm .. 4 hm + 3
              tc_mma_tf32(tmem_base + (uint32_t)((q * MH + hm) * BLOCK_N), da + (uint64_t)((hm * 4 * BOX_BYTES) >> 4) + (uint64_t)(k * (1024 >> 4)),
                          db + (uint64_t)(k * (1024 >> 4)), idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        tc_commit(bar_empty + 8 * s);
      }
      tc_commit(bar_tmem);
    }
    __syncwarp();
  } else if (KB > 0) {
    const int lg = warp & 3;
    mbar_wait(bar_tmem, 0);
    tc_fence_after();
#pragma unroll 1
    for (int q = 0; q < ntap; ++q) {
#pragma unroll 1
      for (int hm = 0; hm < MH; ++hm) {
        const int a = at * (128 * MH) + hm * 128 + lg * 32 + lane;
        float* dst = dw + (int64_t)split * g.part_stride + ((int64_t)(tap0 + q) * g.Ca + a) * g.Cb_pad + bt * BLOCK_N;
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          float v[32];
          tc_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)((q * MH + hm) * BLOCK_N + c * 32), v);
          if (a >= g.Ca) continue;       // rows beyond Ca were TMA zero-fill (Ca = 64 layers use half of the M = 128 tile)
          if (g.splits == 1 || g.part_stride > 0) {
#pragma unroll
            for (int q4 = 0; q4 < 8; ++q4)
              *reinterpret_cast<float4*>(dst + c * 32 + q4 * 4) = make_float4(v[4 * q4], v[4 * q4 + 1], v[4 * q4 + 2], v[4 * q4 + 3]);
          } else {
#pragma unroll
            for (int q1 = 0; q1 < 32; ++q1) atomicAdd(dst + c * 32 + q1, v[q1]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

bool conv_wgrad_tc_supported(const ptk_conv_geom& c) {
  if (!((c.k == 4 && c.stride == 2) || (c.k == 3 && c.stride == 1 && !c.transposed) || (c.k == 1 && c.stride == 1 && c.pad == 0))) return false;
  const int Ca = c.transposed ? c.Cin : c.Cout, Cb = c.transposed ? c.Cout : c.Cin;
  if (Ca % 64 != 0 || Cb % 32 != 0) return false;
  if (c.ldx % 4 != 0 || c.ldy % 4 != 0) return false;
  const int GH = c.transposed ? c.H : c.OH, GW = c.transposed ? c.W : c.OW;
  if (GH < 2 || GW < 2) return false;
  return true;
}

// Legacy mode (nparts == nullptr): dw holds taps*Ca*Cb floats and receives the complete gradient (split-K partials are
// accumulated atomically into a zero fill issued here).  Partial mode (nparts != nullptr): dw has room for
// dw_capacity floats; every split writes its own buffer of taps*Ca*Cb floats with plain stores and *nparts reports
// how many there are -- the caller sums them (ptk_unpack_weight_grad_parts), which makes the result deterministic.
int conv_wgrad_tc(const ptk_conv_geom& c, const float* x, const float* dy, float* dw, int64_t dw_capacity, int* nparts,
                  cudaStream_t st, bool plan_only) {
  WgTcGeom g;
  memset(&g, 0, sizeof(g));
  WgTmapSet maps;
  memset(&maps, 0, sizeof(maps));
  const float *S, *Bg;
  int lda, ldb, BHt, BWt, Cb;
  if (!c.transposed) { S = dy; lda = c.ldy; g.Ca = c.Cout; g.GH = c.OH; g.GW = c.OW; Bg = x; ldb = c.ldx; Cb = c.Cin; BHt = c.H; BWt = c.W; }
  else { S = x; lda = c.ldx; g.Ca = c.Cin; g.GH = c.H; g.GW = c.W; Bg = dy; ldb = c.ldy; Cb = c.Cout; BHt = c.OH; BWt = c.OW; }
  PTK_REQUIRE(plan_only || ((reinterpret_cast<uintptr_t>(S) & 15) == 0 && (reinterpret_cast<uintptr_t>(Bg) & 15) == 0 &&
                            (reinterpret_cast<uintptr_t>(dw) & 15) == 0), "conv_wgrad(tc): pointers must be 16-byte aligned");
  g.N = c.N; g.Cb_pad = Cb;
  g.BW = pow2_ge(g.GW < 32 ? g.GW : 32);
  g.BH = pow2_ge(g.GH < 32 / g.BW ? g.GH : 32 / g.BW);
  g.BI = 32 / (g.BW * g.BH);
  g.tiles_x = (g.GW + g.BW - 1) / g.BW; g.tiles_y = (g.GH + g.BH - 1) / g.BH; g.tiles_i = (c.N + g.BI - 1) / g.BI;
  g.ntiles = g.tiles_x * g.tiles_y * g.tiles_i;
  g.ntaps = c.k * c.k;
  for (int kh = 0; kh < c.k; ++kh)
    for (int kw = 0; kw < c.k; ++kw) {
      const int t = kh * c.k + kw, qy = kh - c.pad, qx = kw - c.pad;
      if (c.stride == 2) {
        g.map[t] = (unsigned char)((((qy % 2) + 2) % 2) * 2 + (((qx % 2) + 2) % 2));
        g.cy[t] = (signed char)floordiv2(qy); g.cx[t] = (signed char)floordiv2(qx);
      } else {
        g.map[t] = 0; g.cy[t] = (signed char)qy; g.cx[t] = (signed char)qx;
      }
    }
  const uint32_t box[4] = {32u, (uint32_t)g.BW, (uint32_t)g.BH, (uint32_t)g.BI};
  if (!plan_only) {
    const uint64_t dims[4] = {(uint64_t)g.Ca, (uint64_t)g.GW, (uint64_t)g.GH, (uint64_t)c.N};
    const uint64_t str[3] = {(uint64_t)lda * 4, (uint64_t)g.GW * lda * 4, (uint64_t)g.GH * g.GW * lda * 4};
    int rc = encode(&maps.s, S, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
  }
  if (plan_only) {
  } else if (c.stride == 2) {
    for (int pyy = 0; pyy < 2; ++pyy)
      for (int pxx = 0; pxx < 2; ++pxx) {
        const uint64_t Hp = (uint64_t)(BHt - pyy + 1) / 2, Wp = (uint64_t)(BWt - pxx + 1) / 2;
        const uint64_t dims[4] = {(uint64_t)Cb, Wp, Hp, (uint64_t)c.N};
        const uint64_t str[3] = {(uint64_t)2 * ldb * 4, (uint64_t)2 * BWt * ldb * 4, (uint64_t)BHt * BWt * ldb * 4};
        int rc = encode(&maps.b[pyy * 2 + pxx], Bg + ((int64_t)pyy * BWt + pxx) * ldb, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        if (rc) return rc;
      }
  } else {
    const uint64_t dims[4] = {(uint64_t)Cb, (uint64_t)BWt, (uint64_t)BHt, (uint64_t)c.N};
    const uint64_t str[3] = {(uint64_t)ldb * 4, (uint64_t)BWt * ldb * 4, (uint64_t)BHt * BWt * ldb * 4};
    int rc = encode(&maps.b[0], Bg, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
  }
  // ---- tile + split-K selection (same reasoning as conv_forward_tc): candidates (MH, BN); for each, the split count
  // that minimises waves x (k-iterations x stage time + fixed prologue/epilogue).
  struct WgCfg { int mh, bn, occ; float eff; int tpc; };
  static const WgCfg kCfgs[] = {{1, 32, 2, 0.30f, 1}, {2, 32, 2, 0.33f, 1}, {1, 64, 2, 0.36f, 1}, {1, 128, 2, 0.50f, 1},
                                {2, 128, 1, 0.64f, 1}, {1, 256, 1, 0.64f, 1}, {2, 256, 1, 0.80f, 1},
                                {1, 32, 1, 0.60f, 9}, {1, 64, 1, 0.60f, 4}};     // several taps per CTA (narrow layers)
  static int tpc_env = -1;
  if (tpc_env < 0) { const char* e = getenv("PTK_WG_TPC"); tpc_env = (e && atoi(e) == 0) ? 0 : 1; }
  int forced_mh = 0, forced_bn = 0;
  if (const char* e = getenv("PTK_WG_TILE")) sscanf(e, "%d,%d", &forced_mh, &forced_bn);
  const int bn_small = (Cb % 128 == 0) ? 128 : (Cb % 64 == 0 ? 64 : 32);
  struct Cand { const WgCfg* cfg; int splits; double cost; };
  Cand cands[32];
  int ncand = 0;
  const Cand* forced_cand = nullptr;
  const int64_t part_floats = (int64_t)g.ntaps * g.Ca * Cb;
  for (const WgCfg& t : kCfgs) {
    if (Cb % t.bn != 0) continue;
    if (t.bn < bn_small) continue;
    if (t.mh == 2 && g.Ca % 256 != 0) continue;
    if (t.tpc > 1 && (!tpc_env || g.Ca > 128 || Cb != t.bn || g.ntaps < t.tpc)) continue;
    const bool forced = forced_mh == t.mh && forced_bn == t.bn && t.tpc == 1;
    const int64_t base_ctas = (int64_t)((g.Ca + 128 * t.mh - 1) / (128 * t.mh)) * (Cb / t.bn) * ((g.ntaps + t.tpc - 1) / t.tpc);
    const int64_t slots = (int64_t)num_sms() * t.occ;
    const double stage_clk = t.tpc * t.mh * (t.bn / 128.0) * 256.0 / t.eff;
    int cfg_splits = 1;
    double cfg_cost = 0.0;
    for (int sp = 1; sp <= g.ntiles && sp <= 512; ++sp) {
      if (sp > 1 && (g.ntiles + sp - 1) / sp * (sp - 1) >= g.ntiles) continue;   // every split must own a pixel tile
      if (nparts != nullptr && sp > 1 && part_floats * sp > dw_capacity) break;
      const int64_t ctas = base_ctas * sp;
      const double tile_clk = (double)((g.ntiles + sp - 1) / sp) * stage_clk + (t.occ == 1 ? 6000.0 : 1500.0);
      // cost of combining the splits: L2 atomics (~100 floats / clk chip-wide) or one more streamed read per part
      const double combine = sp > 1 ? (double)part_floats * sp / (nparts != nullptr ? 800.0 : 100.0) : 0.0;
      const double cost = (double)((ctas + slots - 1) / slots) * t.occ * tile_clk + combine;
      if (sp == 1 || cost < cfg_cost) { cfg_cost = cost; cfg_splits = sp; }
    }
    cands[ncand] = Cand{&t, cfg_splits, cfg_cost};
    ++ncand;
    if (forced) { forced_cand = &cands[ncand - 1]; break; }
    // variants for the first-use autotuner: half / double the split count
    auto valid_split = [&](int q) {
      if (q < 1 || q == cfg_splits || q > g.ntiles || q > 512) return false;
      if (q > 1 && (g.ntiles + q - 1) / q * (q - 1) >= g.ntiles) return false;
      if (nparts != nullptr && q > 1 && part_floats * q > dw_capacity) return false;
      return true;
    };
    if (ncand < 30 && valid_split(cfg_splits / 2)) cands[ncand++] = Cand{&t, cfg_splits / 2, cfg_cost * 1.05};
    if (ncand < 30 && valid_split(cfg_splits * 2)) cands[ncand++] = Cand{&t, cfg_splits * 2, cfg_cost * 1.05};
  }
  PTK_REQUIRE(ncand > 0, "conv_wgrad(tc): no tile configuration for Ca=%d Cb=%d", g.Ca, Cb);
  const Cand* pick = forced_cand;
  if (!pick) {
    for (int i = 1; i < ncand; ++i)          // by model cost (stable)
      for (int j = i; j > 0 && cands[j].cost < cands[j - 1].cost; --j) { const Cand tmp = cands[j]; cands[j] = cands[j - 1]; cands[j - 1] = tmp; }
    pick = &cands[0];
  }

  const WgTcGeom g0 = g;
  auto run = [&](const WgCfg* best, int best_splits, bool launch) -> int {
  WgTcGeom g = g0;
  const int BN = best->bn, MH = best->mh, TPC = best->tpc;
  const int atiles = (g.Ca + 128 * MH - 1) / (128 * MH);
  const int splits = best_splits;
  g.splits = splits;
  g.part_stride = nparts != nullptr ? (long long)g.ntaps * g.Ca * Cb : 0;
  if (nparts != nullptr) *nparts = splits;
  if (!launch) return 0;
  if (splits > 1 && nparts == nullptr) {
    int rc = ptk_fill(dw, (int64_t)g.ntaps * g.Ca * Cb, 0.f, st);
    if (rc) return rc;
  }
  dim3 grid((unsigned)(atiles * (Cb / BN)), (unsigned)((g.ntaps + TPC - 1) / TPC), (unsigned)splits);
#define PTK_WG_LAUNCH(BN_, ST_, MH_, TPC_)                                                                                 \
  do {                                                                                                                     \
    const size_t smem = (size_t)ST_ * (MH_ * 128 * 128 + TPC_ * BN_ * 128) + 16 * ST_ + 16 + 1024;                          \
    static bool attr = false;                                                                                              \
    if (!attr) { cudaFuncSetAttribute(wgrad_tc_kernel<BN_, ST_, MH_, TPC_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; } \
    launch_pdl(wgrad_tc_kernel<BN_, ST_, MH_, TPC_>, grid, dim3(192), smem, st, maps, g, dw);                                \
  } while (0)
  if (TPC == 9) PTK_WG_LAUNCH(32, 3, 1, 9);
  else if (TPC == 4) PTK_WG_LAUNCH(64, 4, 1, 4);
  else if (MH == 1 && BN == 32) PTK_WG_LAUNCH(32, 3, 1, 1);
  else if (MH == 2 && BN == 32) PTK_WG_LAUNCH(32, 3, 2, 1);
  else if (MH == 1 && BN == 64) PTK_WG_LAUNCH(64, 3, 1, 1);
  else if (MH == 1 && BN == 128) PTK_WG_LAUNCH(128, 3, 1, 1);
  else if (MH == 1 && BN == 256) PTK_WG_LAUNCH(256, 4, 1, 1);
  else if (MH == 2 && BN == 128) PTK_WG_LAUNCH(128, 4, 2, 1);
  else PTK_WG_LAUNCH(256, 3, 2, 1);
#undef PTK_WG_LAUNCH
  PTK_LAUNCH_CHECK("wgrad_tc_kernel");
  return 0;
  };   // run

  // First real launch of this geometry in partial-buffer mode: time the model's best candidates (those whose split count
  // fits the caller's buffer) and keep the winner; plan-only calls and later launches read the remembered choice.
  if (!forced_cand && ncand > 1 && nparts != nullptr && tc_autotune_enabled(st)) {
    const int64_t allowed = part_floats > 0 ? dw_capacity / part_floats : 1;
    const TuneKey key = tune_key(c, 1, 0, (allowed > 1024 ? 1024 : allowed) << 20);
    TuneChoice ch;
    bool have = tune_lookup(key, &ch);
    if (!have && !plan_only) {
      const int ntry = ncand < 12 ? ncand : 12;
      float best_ms = 0.f;
      int best_i = 0;
      for (int i = 0; i < ntry; ++i) {
        float ms = 0.f;
        int rc = tune_time(st, [&]() { return run(cands[i].cfg, cands[i].splits, true); }, &ms);
        if (rc) return rc;
        if (i == 0 || ms < best_ms) { best_ms = ms; best_i = i; }
      }
      ch.mh = cands[best_i].cfg->mh; ch.bn = cands[best_i].cfg->bn; ch.splits = cands[best_i].splits; ch.tpc = cands[best_i].cfg->tpc;
      ch.persist = 1;
      tune_store(key, ch);
      have = true;
    }
    if (have)
      for (int i = 0; i < ncand; ++i)
        if (cands[i].cfg->mh == ch.mh && cands[i].cfg->bn == ch.bn && cands[i].cfg->tpc == ch.tpc && cands[i].splits == ch.splits) { pick = &cands[i]; break; }
  }
  return run(pick->cfg, pick->splits, !plan_only);
}

}  // namespace ptk
