// "InstanceNorm3d(1, eps=1e-3, affine)" of the reference Block (models/networks.py:159,164-172): per-sample
// mean / biased variance over C*H*W with a scalar gain and bias, fused with the activation that the NEXT
// Block applies to its input (LeakyReLU(0.2) / ReLU, models/networks.py:150,152), with Dropout2d
// (models/networks.py:161) and with the concat placement (writes into channel slices).  HBM-bound.
#include "common.cuh"

namespace ptk {

constexpr float kEps = 1e-3f;

__device__ __forceinline__ void block_reduce2_atomic(double a, double b, double* out) {
  __shared__ double ra[32], rb[32];
  a = warp_sum(a); b = warp_sum(b);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { ra[wid] = a; rb[wid] = b; }
  __syncthreads();
  if (wid == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    a = lane < nw ? ra[lane] : 0.0; b = lane < nw ? rb[lane] : 0.0;
    a = warp_sum(a); b = warp_sum(b);
    if (lane == 0) { atomicAdd(out, a); atomicAdd(out + 1, b); }
  }
}

// grid (chunks, N); every thread strides over float4 groups of sample n, two independent loads in flight.
__global__ void __launch_bounds__(256)
gn_stats_kernel(const float* __restrict__ z, int ld, int64_t HW, int C, double* __restrict__ stats) {
  pdl_trigger();
  const int n = blockIdx.y;
  const int C4 = C >> 2;
  const int total4 = (int)(HW * C4);          // < 2^31 float4 groups per sample (checked by the host)
  const float* base = z + (int64_t)n * HW * ld;
  float s = 0.f, q = 0.f;
  double ds = 0.0, dq = 0.0;
  int cnt = 0;
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += 2 * stride) {
    const int i2 = i + stride;
    const int p = i / C4;
    const int c = (i - p * C4) << 2;
    const float4 v = __ldg(reinterpret_cast<const float4*>(base + (int64_t)p * ld + c));
    float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i2 < total4) {
      const int p2 = i2 / C4;
      const int c2 = (i2 - p2 * C4) << 2;
      u = __ldg(reinterpret_cast<const float4*>(base + (int64_t)p2 * ld + c2));
    }
    s += ((v.x + v.y) + (v.z + v.w)) + ((u.x + u.y) + (u.z + u.w));
    q += ((v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w)) + ((u.x * u.x + u.y * u.y) + (u.z * u.z + u.w * u.w));
    if (++cnt == 32) { ds += s; dq += q; s = 0.f; q = 0.f; cnt = 0; }
  }
  ds += s; dq += q;
  block_reduce2_atomic(ds, dq, stats + 2 * n);
}

// scalar fallback when C % 4 != 0 (PatchGAN head: C = 1 has no norm, so this is only for generality)
__global__ void __launch_bounds__(256)
gn_stats_scalar_kernel(const float* __restrict__ z, int ld, int64_t HW, int C, double* __restrict__ stats) {
  const int n = blockIdx.y;
  const int64_t total = HW * C;
  const float* base = z + (int64_t)n * HW * ld;
  double ds = 0.0, dq = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / C;
    const float v = base[p * ld + (i - p * C)];
    ds += v; dq += (double)v * v;
  }
  block_reduce2_atomic(ds, dq, stats + 2 * n);
}

__device__ __forceinline__ void mean_rstd(const double* stats, int n, double count, float* mean, float* rstd) {
  const double m = stats[2 * n] / count;
  double var = stats[2 * n + 1] / count - m * m;
  if (var < 0.0) var = 0.0;
  *mean = (float)m;
  *rstd = (float)(1.0 / sqrt(var + (double)kEps));
}

// Element indexing shared by the streaming kernels below: when the total thread count is a multiple of C/4 (host
// guarantees it on this path), a thread keeps ONE channel group c and walks pixels with a constant stride, so the loop
// body has no integer division and U independent 128-bit loads per operand are in flight per thread.
struct PixWalk { int c; int64_t p, step, npix; };
__device__ __forceinline__ PixWalk pix_walk(int64_t HW, int C4) {
  PixWalk w;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  w.c = (int)(t % C4) << 2;
  w.p = t / C4;
  w.step = ((int64_t)gridDim.x * blockDim.x) / C4;
  w.npix = HW;
  return w;
}

template <int U>
__global__ void __launch_bounds__(256)
gn_apply_kernel(const float* __restrict__ z, int ldz, const double* __restrict__ stats,
                const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ drop,
                int64_t HW, int C, float* __restrict__ out1, int ld1, int act1, float* __restrict__ out2,
                int ld2, int act2) {
  pdl_trigger();
  const int n = blockIdx.y;
  float scale = 1.f, shift = 0.f;
  if (gamma) {
    float mean, rstd;
    mean_rstd(stats, n, (double)HW * C, &mean, &rstd);
    const float g = __ldg(gamma), b = __ldg(beta);
    scale = rstd * g;
    shift = b - mean * scale;
  }
  const PixWalk w = pix_walk(HW, C >> 2);
  const float* zb = z + (int64_t)n * HW * ldz + w.c;
  float* o1 = out1 + (int64_t)n * HW * ld1 + w.c;
  float* o2 = out2 ? out2 + (int64_t)n * HW * ld2 + w.c : nullptr;
  float4 d = make_float4(1.f, 1.f, 1.f, 1.f);
  if (drop) d = __ldg(reinterpret_cast<const float4*>(drop + (int64_t)n * C + w.c));
  for (int64_t p = w.p; p < HW; p += U * w.step) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t q = p + u * w.step;
      if (q < HW) v[u] = __ldg(reinterpret_cast<const float4*>(zb + q * ldz));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t q = p + u * w.step;
      if (q >= HW) break;
      float4 t = v[u];
      t.x = fmaf(t.x, scale, shift) * d.x; t.y = fmaf(t.y, scale, shift) * d.y;
      t.z = fmaf(t.z, scale, shift) * d.z; t.w = fmaf(t.w, scale, shift) * d.w;
      *reinterpret_cast<float4*>(o1 + q * ld1) = make_float4(apply_act(t.x, act1), apply_act(t.y, act1), apply_act(t.z, act1), apply_act(t.w, act1));
      if (o2) *reinterpret_cast<float4*>(o2 + q * ld2) = make_float4(apply_act(t.x, act2), apply_act(t.y, act2), apply_act(t.z, act2), apply_act(t.w, act2));
    }
  }
}

template <int U, int OCC>
__global__ void __launch_bounds__(256, OCC)
gn_bwd_reduce_kernel(const float* __restrict__ g1, int ldg1, const float* __restrict__ a1, int lda1, int act1,
                     const float* __restrict__ g2, int ldg2, const float* __restrict__ a2, int lda2, int act2,
                     const float* __restrict__ drop, const float* __restrict__ z, int ldz,
                     const double* __restrict__ stats, int64_t HW, int C, float* __restrict__ dy,
                     double* __restrict__ sums) {
  pdl_trigger();
  const int n = blockIdx.y;
  float mean = 0.f, rstd = 1.f;
  const bool normed = sums != nullptr;
  if (normed) mean_rstd(stats, n, (double)HW * C, &mean, &rstd);
  const PixWalk w = pix_walk(HW, C >> 2);
  const int64_t pix0 = (int64_t)n * HW;
  g1 += pix0 * ldg1 + w.c;
  if (a1) a1 += pix0 * lda1 + w.c;
  if (g2) g2 += pix0 * ldg2 + w.c;
  if (a2) a2 += pix0 * lda2 + w.c;
  if (normed) z += pix0 * ldz + w.c;
  dy += pix0 * C + w.c;
  float4 d = make_float4(1.f, 1.f, 1.f, 1.f);
  if (drop) d = __ldg(reinterpret_cast<const float4*>(drop + (int64_t)n * C + w.c));
  float s1 = 0.f, s2 = 0.f;
  double d1 = 0.0, d2 = 0.0;
  int cnt = 0;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t p = w.p; p < HW; p += U * w.step) {
    float4 vg[U], va[U], vh[U], vb[U], vz[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t q = p + u * w.step;
      const bool in = q < HW;
      vg[u] = in ? __ldg(reinterpret_cast<const float4*>(g1 + q * ldg1)) : zero;
      va[u] = (in && a1) ? __ldg(reinterpret_cast<const float4*>(a1 + q * lda1)) : zero;
      vh[u] = (in && g2) ? __ldg(reinterpret_cast<const float4*>(g2 + q * ldg2)) : zero;
      vb[u] = (in && a2) ? __ldg(reinterpret_cast<const float4*>(a2 + q * lda2)) : zero;
      vz[u] = (in && normed) ? __ldg(reinterpret_cast<const float4*>(z + q * ldz)) : zero;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t q = p + u * w.step;
      if (q >= HW) break;
      float4 g = vg[u];
      if (a1) {
        g.x *= act_grad_from_output(va[u].x, act1); g.y *= act_grad_from_output(va[u].y, act1);
        g.z *= act_grad_from_output(va[u].z, act1); g.w *= act_grad_from_output(va[u].w, act1);
      }
      if (g2) {
        float4 h = vh[u];
        if (a2) {
          h.x *= act_grad_from_output(vb[u].x, act2); h.y *= act_grad_from_output(vb[u].y, act2);
          h.z *= act_grad_from_output(vb[u].z, act2); h.w *= act_grad_from_output(vb[u].w, act2);
        }
        g.x += h.x; g.y += h.y; g.z += h.z; g.w += h.w;
      }
      g.x *= d.x; g.y *= d.y; g.z *= d.z; g.w *= d.w;
      *reinterpret_cast<float4*>(dy + q * C) = g;
      if (normed) {
        s1 += (g.x + g.y) + (g.z + g.w);
        s2 += g.x * (vz[u].x - mean) + g.y * (vz[u].y - mean) + g.z * (vz[u].z - mean) + g.w * (vz[u].w - mean);
        if (++cnt == 64) { d1 += s1; d2 += s2; s1 = 0.f; s2 = 0.f; cnt = 0; }
      }
    }
  }
  if (normed) {
    d1 += s1; d2 += s2;
    block_reduce2_atomic(d1, d2 * (double)rstd, sums + 2 * n);
  }
}

template <int U>
__global__ void __launch_bounds__(256)
gn_bwd_apply_kernel(float* __restrict__ dy, const float* __restrict__ z, int ldz, const double* __restrict__ stats,
                    const double* __restrict__ sums, const float* __restrict__ gamma, int N, int64_t HW, int C,
                    float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_trigger();
  const int n = blockIdx.y;
  float mean, rstd;
  const double count = (double)HW * C;
  mean_rstd(stats, n, count, &mean, &rstd);
  const float g = __ldg(gamma);
  const float m1 = (float)(sums[2 * n] / count), m2 = (float)(sums[2 * n + 1] / count);
  const float k = g * rstd;
  const PixWalk w = pix_walk(HW, C >> 2);
  float* dyb = dy + (int64_t)n * HW * C + w.c;
  const float* zb = z + (int64_t)n * HW * ldz + w.c;
  for (int64_t p = w.p; p < HW; p += U * w.step) {
    float4 vd[U], vz[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t q = p + u * w.step;
      if (q < HW) {
        vd[u] = *reinterpret_cast<const float4*>(dyb + q * C);
        vz[u] = __ldg(reinterpret_cast<const float4*>(zb + q * ldz));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t q = p + u * w.step;
      if (q >= HW) break;
      float4 d = vd[u];
      d.x = k * (d.x - m1 - (vz[u].x - mean) * rstd * m2);
      d.y = k * (d.y - m1 - (vz[u].y - mean) * rstd * m2);
      d.z = k * (d.z - m1 - (vz[u].z - mean) * rstd * m2);
      d.w = k * (d.w - m1 - (vz[u].w - mean) * rstd * m2);
      *reinterpret_cast<float4*>(dyb + q * C) = d;
    }
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    double sg = 0.0, sb = 0.0;
    for (int i = 0; i < N; ++i) { sb += sums[2 * i]; sg += sums[2 * i + 1]; }
    *dgamma += (float)sg;
    *dbeta += (float)sb;
  }
}

// grid for the PixWalk kernels: (blocks * 256) % (C / 4) == 0 so that a thread keeps its channel group; about 8 blocks
// of 256 threads per SM over the whole batch, never more threads than U-element work items
static inline dim3 grid_walk(int64_t HW, int C, int N, int U) {
  const int C4 = C / 4;
  int64_t quantum = 1;                       // smallest block count with (blocks * 256) % C4 == 0
  while ((quantum * 256) % C4 != 0) ++quantum;
  int64_t want = (HW * C4 + (int64_t)256 * U - 1) / ((int64_t)256 * U);
  int64_t cap = ((int64_t)num_sms() * 8 + N - 1) / N;
  if (want > cap) want = cap;
  int64_t b = (want + quantum - 1) / quantum * quantum;
  if (b < quantum) b = quantum;
  return dim3((unsigned)b, (unsigned)N);
}

static inline dim3 grid2(int64_t work, int N) {
  int64_t b = (work + 255) / 256;
  int64_t cap = ((int64_t)num_sms() * 8 + N - 1) / N;
  if (cap < 1) cap = 1;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return dim3((unsigned)b, (unsigned)N);
}

}  // namespace ptk

using namespace ptk;

extern "C" int ptk_gn_stats(const float* z, int ld, int N, int64_t HW, int C, double* stats, void* stream) {
  PTK_REQUIRE(N > 0 && N <= 65535 && HW > 0 && C > 0 && HW * C < (1ll << 32), "gn_stats: bad extents");
  if (C % 4 == 0 && ld % 4 == 0 && (reinterpret_cast<uintptr_t>(z) & 15) == 0) {
    gn_stats_kernel<<<grid2(HW * (C / 4), N), 256, 0, (cudaStream_t)stream>>>(z, ld, HW, C, stats);
  } else {
    gn_stats_scalar_kernel<<<grid2(HW * C, N), 256, 0, (cudaStream_t)stream>>>(z, ld, HW, C, stats);
  }
  PTK_LAUNCH_CHECK("gn_stats_kernel");
  return 0;
}

extern "C" int ptk_gn_apply(const float* z, int ldz, const double* stats, const float* gamma, const float* beta,
                            const float* drop, int N, int64_t HW, int C, float* out1, int ld1, int act1,
                            float* out2, int ld2, int act2, void* stream) {
  PTK_REQUIRE(N > 0 && N <= 65535 && HW > 0 && C > 0 && HW * C < (1ll << 32), "gn_apply: bad extents");
  PTK_REQUIRE(C % 4 == 0 && ldz % 4 == 0 && ld1 % 4 == 0 && (!out2 || ld2 % 4 == 0), "gn_apply: C and strides must be multiples of 4");
  PTK_REQUIRE((gamma == nullptr) == (beta == nullptr), "gn_apply: gamma and beta must both be given or both NULL");
  PTK_REQUIRE(!gamma || stats, "gn_apply: stats required");
  PTK_REQUIRE(C / 4 <= 4096, "gn_apply: C too large");
  gn_apply_kernel<4><<<grid_walk(HW, C, N, 4), 256, 0, (cudaStream_t)stream>>>(z, ldz, stats, gamma, beta, drop, HW, C, out1,
                                                                               ld1, act1, out2, ld2, act2);
  PTK_LAUNCH_CHECK("gn_apply_kernel");
  return 0;
}

extern "C" int ptk_gn_bwd_reduce(const float* g1, int ldg1, const float* a1, int lda1, int act1, const float* g2,
                                 int ldg2, const float* a2, int lda2, int act2, const float* drop, const float* z,
                                 int ldz, const double* stats, int N, int64_t HW, int C, float* dy, double* sums,
                                 void* stream) {
  PTK_REQUIRE(N > 0 && N <= 65535 && HW > 0 && C > 0 && C % 4 == 0 && HW * C < (1ll << 32), "gn_bwd_reduce: bad extents (C %% 4 == 0 required)");
  PTK_REQUIRE(ldg1 % 4 == 0 && (!a1 || lda1 % 4 == 0) && (!g2 || ldg2 % 4 == 0) && (!a2 || lda2 % 4 == 0) && (!sums || ldz % 4 == 0),
              "gn_bwd_reduce: strides must be multiples of 4");
  PTK_REQUIRE(!sums || (z && stats), "gn_bwd_reduce: z and stats required with sums");
  PTK_REQUIRE(C / 4 <= 4096, "gn_bwd_reduce: C too large");
  // four resident CTAs per SM (64 registers, a few spilled scalars) stream the large tensors 10-15 % faster; below ~1 M
  // elements per sample the three-CTA build wins (tools/bench_gn.py on B200)
  if (HW * C < (1ll << 20))
    gn_bwd_reduce_kernel<2, 3><<<grid_walk(HW, C, N, 2), 256, 0, (cudaStream_t)stream>>>(g1, ldg1, a1, lda1, act1, g2, ldg2, a2, lda2,
                                                                                         act2, drop, z, ldz, stats, HW, C, dy, sums);
  else
    gn_bwd_reduce_kernel<2, 4><<<grid_walk(HW, C, N, 2), 256, 0, (cudaStream_t)stream>>>(g1, ldg1, a1, lda1, act1, g2, ldg2, a2, lda2,
                                                                                         act2, drop, z, ldz, stats, HW, C, dy, sums);
  PTK_LAUNCH_CHECK("gn_bwd_reduce_kernel");
  return 0;
}

extern "C" int ptk_gn_bwd_apply(float* dy, const float* z, int ldz, const double* stats, const double* sums,
                                const float* gamma, int N, int64_t HW, int C, float* dgamma, float* dbeta,
                                void* stream) {
  PTK_REQUIRE(N > 0 && N <= 65535 && HW > 0 && C > 0 && C % 4 == 0 && ldz % 4 == 0 && HW * C < (1ll << 32), "gn_bwd_apply: bad extents");
  PTK_REQUIRE(C / 4 <= 4096, "gn_bwd_apply: C too large");
  gn_bwd_apply_kernel<4><<<grid_walk(HW, C, N, 4), 256, 0, (cudaStream_t)stream>>>(dy, z, ldz, stats, sums, gamma, N, HW, C,
                                                                                  dgamma, dbeta);
  PTK_LAUNCH_CHECK("gn_bwd_apply_kernel");
  return 0;
}
