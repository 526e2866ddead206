// fp32 CUDA-core implicit-GEMM convolutions ("exact" path, PTK_IMPL_SIMT) for sm_100a.
//
// One gather-GEMM formulation serves Conv2d fprop, ConvTranspose2d fprop and both dgrads:
//     y[n, gy*so+py, gx*so+px, co] = sum_t sum_ci x[n, gy*sm+dy[t], gx*sm+dx[t], ci] * w[wt[t]][ci][co]
// over a base grid (gy,gx) in GH x GW.  A strided conv has sm=stride, so=1 and all k*k taps; a
// transposed conv is split into stride*stride output-parity phases, each a dense stride-1 gather with
// (k/stride)^2 taps (no multiply-by-zero work).  Replaces the cuDNN calls behind
// models/networks.py:154,156-157,186,232,341 of the reference.
#include "common.cuh"

namespace ptk {

struct TapGeom {
  int N, H, W, Cin, ldx;
  int OH, OW, Cout, ldy;
  int GH, GW;
  int sm;
  int so, py, px;
  int ntaps;
  int Cout_pad;  // inner extent of the packed weight [tap][Cin][Cout_pad]
  signed char dy[16], dx[16];
  unsigned char wt[16];
};

// Build the tap table of one phase.  Returns ntaps (0 => phase has no taps: output is bias only).
static void make_geom(const ptk_conv_geom& c, int py, int px, int cout_pad, TapGeom* g) {
  g->N = c.N; g->H = c.H; g->W = c.W; g->Cin = c.Cin; g->ldx = c.ldx;
  g->OH = c.OH; g->OW = c.OW; g->Cout = c.Cout; g->ldy = c.ldy; g->Cout_pad = cout_pad;
  g->ntaps = 0;
  if (!c.transposed) {
    g->GH = c.OH; g->GW = c.OW; g->sm = c.stride; g->so = 1; g->py = 0; g->px = 0;
    for (int kh = 0; kh < c.k; ++kh)
      for (int kw = 0; kw < c.k; ++kw) {
        int t = g->ntaps++;
        g->dy[t] = (signed char)(kh - c.pad); g->dx[t] = (signed char)(kw - c.pad);
        g->wt[t] = (unsigned char)(kh * c.k + kw);
      }
  } else {
    const int s = c.stride;
    g->GH = (c.OH - py + s - 1) / s; g->GW = (c.OW - px + s - 1) / s;
    g->sm = 1; g->so = s; g->py = py; g->px = px;
    // oy = iy*s - pad + kh  =>  iy = (gy*s + py + pad - kh)/s = gy + (py + pad - kh)/s when divisible
    for (int kh = 0; kh < c.k; ++kh) {
      if ((py + c.pad - kh) % s != 0) continue;
      for (int kw = 0; kw < c.k; ++kw) {
        if ((px + c.pad - kw) % s != 0) continue;
        int t = g->ntaps++;
        g->dy[t] = (signed char)((py + c.pad - kh) / s); g->dx[t] = (signed char)((px + c.pad - kw) / s);
        g->wt[t] = (unsigned char)(kh * c.k + kw);
      }
    }
  }
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

template <int BN>
__global__ void __launch_bounds__(256)
conv_gemm_simt_kernel(TapGeom g, const float* __restrict__ x, const float* __restrict__ w,
                      const float* __restrict__ bias, int act, float* __restrict__ y) {
  constexpr int BM = 128, BK = 16, TN = BN / 16, NB4 = BN / 64;
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  __shared__ int s_dy[16], s_dx[16], s_wt[16];
  const int tid = threadIdx.x;
  if (tid < 16) { s_dy[tid] = g.dy[tid]; s_dx[tid] = g.dx[tid]; s_wt[tid] = g.wt[tid]; }
  __syncthreads();
  const int64_t GP = (int64_t)g.GH * g.GW;
  const int64_t M = (int64_t)g.N * GP;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  const int arow = tid & 127, akh = tid >> 7;
  const int64_t am = m0 + arow;
  const bool arow_ok = am < M;
  int an = 0, agy = 0, agx = 0;
  if (arow_ok) {
    an = (int)(am / GP);
    int r = (int)(am - (int64_t)an * GP);
    agy = r / g.GW; agx = r - agy * g.GW;
  }
  const int chunks = (g.Cin + BK - 1) / BK;
  const int T = g.ntaps * chunks;

  float4 a_reg[2];
  float4 b_reg[NB4 > 0 ? NB4 : 1];
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

  auto load_tile = [&](int tile) {
    const int t = tile / chunks;
    const int c0 = (tile - t * chunks) * BK;
    const int iy = agy * g.sm + s_dy[t], ix = agx * g.sm + s_dx[t];
    const bool ok = arow_ok && iy >= 0 && iy < g.H && ix >= 0 && ix < g.W;
    const float* src = x + (((int64_t)an * g.H + iy) * g.W + ix) * g.ldx + c0 + akh * 8;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = c0 + akh * 8 + i * 4;
      a_reg[i] = (ok && c < g.Cin) ? ldg4(src + i * 4) : zero4;
    }
    const float* wsrc = w + ((int64_t)s_wt[t] * g.Cin + c0) * g.Cout_pad + n0;
#pragma unroll
    for (int i = 0; i < NB4; ++i) {
      const int idx = tid + i * 256;
      const int brow = idx / (BN / 4), bc = (idx % (BN / 4)) * 4;
      b_reg[i] = (c0 + brow < g.Cin && n0 + bc < g.Cout_pad) ? ldg4(wsrc + (int64_t)brow * g.Cout_pad + bc) : zero4;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int k = akh * 8 + i * 4;
      As[buf][k + 0][arow] = a_reg[i].x; As[buf][k + 1][arow] = a_reg[i].y;
      As[buf][k + 2][arow] = a_reg[i].z; As[buf][k + 3][arow] = a_reg[i].w;
    }
#pragma unroll
    for (int i = 0; i < NB4; ++i) {
      const int idx = tid + i * 256;
      const int brow = idx / (BN / 4), bc = (idx % (BN / 4)) * 4;
      *reinterpret_cast<float4*>(&Bs[buf][brow][bc]) = b_reg[i];
    }
  };

  const int tx = tid & 15, ty = tid >> 4;
  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  if (T > 0) {
    load_tile(0);
    int buf = 0;
    for (int tile = 0; tile < T; ++tile) {
      store_tile(buf);
      __syncthreads();
      if (tile + 1 < T) load_tile(tile + 1);
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float a[8], b[TN];
        const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
        for (int q = 0; q < TN / 4; ++q) {
          const float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][kk][q * 64 + tx * 4]);
          b[q * 4 + 0] = bv.x; b[q * 4 + 1] = bv.y; b[q * 4 + 2] = bv.z; b[q * 4 + 3] = bv.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      buf ^= 1;
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + ty * 8 + i;
    if (m >= M) continue;
    const int n = (int)(m / GP);
    const int r = (int)(m - (int64_t)n * GP);
    const int gy = r / g.GW, gx = r - gy * g.GW;
    const int oy = gy * g.so + g.py, ox = gx * g.so + g.px;
    float* dst = y + (((int64_t)n * g.OH + oy) * g.OW + ox) * g.ldy;
#pragma unroll
    for (int q = 0; q < TN / 4; ++q) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int co = n0 + q * 64 + tx * 4 + j;
        if (co < g.Cout) {
          float v = acc[i][q * 4 + j];
          if (bias) v += __ldg(bias + co);
          dst[co] = apply_act(v, act);
        }
      }
    }
  }
}

// ---------------------------------------------------------------- narrow-N forward (Cout <= 4)
// One warp per output pixel, lanes split the channel axis (float4), shuffle-reduce.  Used for the
// generator's final Conv(256->3,k3)+tanh (models/networks.py:232 + Tanh) and the PatchGAN head
// Conv(512->1,k4,s2,p1) (models/networks.py:349).
__global__ void __launch_bounds__(256)
conv_narrow_fwd_kernel(TapGeom g, const float* __restrict__ x, const float* __restrict__ w,
                       const float* __restrict__ bias, int act, float* __restrict__ y,
                       float* __restrict__ y_nchw) {
  extern __shared__ float s_w[];  // [ntaps][Cin][Cout]
  const int tid = threadIdx.x;
  const int wcount = g.ntaps * g.Cin * g.Cout;
  for (int i = tid; i < wcount; i += blockDim.x) {
    const int co = i % g.Cout;
    const int r = i / g.Cout;
    const int ci = r % g.Cin, t = r / g.Cin;
    s_w[i] = w[((int64_t)g.wt[t] * g.Cin + ci) * g.Cout_pad + co];
  }
  __syncthreads();
  const int lane = tid & 31;
  const int64_t GP = (int64_t)g.GH * g.GW;
  const int64_t M = (int64_t)g.N * GP;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (tid >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t m = warp0; m < M; m += nwarps) {
    const int n = (int)(m / GP);
    const int r = (int)(m - (int64_t)n * GP);
    const int gy = r / g.GW, gx = r - gy * g.GW;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int t = 0; t < g.ntaps; ++t) {
      const int iy = gy * g.sm + g.dy[t], ix = gx * g.sm + g.dx[t];
      if (iy < 0 || iy >= g.H || ix < 0 || ix >= g.W) continue;
      const float* src = x + (((int64_t)n * g.H + iy) * g.W + ix) * g.ldx;
      const float* wt = s_w + (int64_t)t * g.Cin * g.Cout;
      for (int c = lane * 4; c < g.Cin; c += 128) {
        const float4 v = ldg4(src + c);
        const float* wp = wt + c * g.Cout;
#pragma unroll
        for (int co = 0; co < 4; ++co)
          if (co < g.Cout)
            acc[co] += v.x * wp[co] + v.y * wp[g.Cout + co] + v.z * wp[2 * g.Cout + co] + v.w * wp[3 * g.Cout + co];
      }
    }
#pragma unroll
    for (int co = 0; co < 4; ++co) acc[co] = warp_sum(acc[co]);
    if (lane == 0) {
      const int oy = gy * g.so + g.py, ox = gx * g.so + g.px;
      for (int co = 0; co < g.Cout; ++co) {
        float v = acc[co] + (bias ? bias[co] : 0.f);
        v = apply_act(v, act);
        if (y) y[(((int64_t)n * g.OH + oy) * g.OW + ox) * g.ldy + co] = v;
        if (y_nchw) y_nchw[(((int64_t)n * g.Cout + co) * g.OH + oy) * g.OW + ox] = v;
      }
    }
  }
}


// ---------------------------------------------------------------- k3/s1/p1 narrow-N convolution (generator head)
// Conv(256->3,k3,p1)+bias+tanh (models/networks.py:232 + Tanh): 3.6 GFMA per 8 images but only 3 output channels, so
// it is scheduled as a stencil: 16x16-pixel tiles, 16-channel chunks staged in shared memory (halo 1), weights in
// __constant__ memory so every FFMA takes its weight as a uniform constant operand (no load instruction).
constexpr int kNarrowMaxW = 12288;                 // floats: taps(9) * Cin(<=256) * 4 fits with room to spare
__constant__ float c_narrow_w[kNarrowMaxW];        // [tap][ci][4]

__global__ void __launch_bounds__(256)
conv_k3_narrow_fwd_kernel(const float* __restrict__ x, int ldx, int Cin, int H, int W, int Cout,
                          const float* __restrict__ bias, int act, float* __restrict__ y, int ldy,
                          float* __restrict__ y_nchw) {
  constexpr int T = 16, R = T + 2, CH = 16, LD = 20;   // LD: padded pixel stride in floats (bank-conflict free float4 reads)
  __shared__ __align__(16) float s_in[R * R * LD];
  const int n = blockIdx.z, ty0 = blockIdx.y * T, tx0 = blockIdx.x * T;
  const int tid = threadIdx.x, ly = tid / T, lx = tid % T;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const float* xb = x + (int64_t)n * H * W * ldx;
  for (int c0 = 0; c0 < Cin; c0 += CH) {
    __syncthreads();
    for (int i = tid; i < R * R * (CH / 4); i += 256) {
      const int p = i / (CH / 4), c4 = (i % (CH / 4)) * 4;
      const int yy = ty0 - 1 + p / R, xx = tx0 - 1 + p % R;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = ldg4(xb + ((int64_t)yy * W + xx) * ldx + c0 + c4);
      *reinterpret_cast<float4*>(&s_in[p * LD + c4]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float* ip = &s_in[((ly + ky) * R + lx + kx) * LD];
        const float* wp = &c_narrow_w[((ky * 3 + kx) * Cin + c0) * 4];
#pragma unroll
        for (int c4 = 0; c4 < CH; c4 += 4) {
          const float4 v = *reinterpret_cast<const float4*>(ip + c4);
          const float in[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc[0] = fmaf(in[j], wp[(c4 + j) * 4 + 0], acc[0]);
            acc[1] = fmaf(in[j], wp[(c4 + j) * 4 + 1], acc[1]);
            acc[2] = fmaf(in[j], wp[(c4 + j) * 4 + 2], acc[2]);
            acc[3] = fmaf(in[j], wp[(c4 + j) * 4 + 3], acc[3]);
          }
        }
      }
  }
  const int oy = ty0 + ly, ox = tx0 + lx;
  if (oy < H && ox < W) {
    for (int co = 0; co < Cout; ++co) {
      float v = acc[co] + (bias ? __ldg(bias + co) : 0.f);
      v = apply_act(v, act);
      if (y) y[(((int64_t)n * H + oy) * W + ox) * ldy + co] = v;
      if (y_nchw) y_nchw[(((int64_t)n * Cout + co) * H + oy) * W + ox] = v;
    }
  }
}

// Weight gradient of the same layer: dw[tap][co][ci] += sum_pixels dy[pix][co] * x[pix + tap][ci].
// One thread per input channel, marching along input rows with a sliding 3x3 window of dy (float4 = 3 channels +
// zero pad per pixel) kept in registers: per input pixel 1 coalesced load of x, 3 broadcast loads of dy, 27 FMAs.
__global__ void __launch_bounds__(256)
conv_k3_narrow_wgrad_kernel(const float* __restrict__ x, int ldx, int Cin, const float* __restrict__ dy, int lddy,
                            int N, int H, int W, int Cout, int rows_per_block, float* __restrict__ dw) {
  const int ci = blockIdx.y * 256 + threadIdx.x;
  const bool live = ci < Cin;
  const int64_t row0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t rows = (int64_t)N * H;
  float acc[9][3];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t][0] = acc[t][1] = acc[t][2] = 0.f;
  for (int64_t r = row0; r < row0 + rows_per_block && r < rows; ++r) {
    const int n = (int)(r / H), yi = (int)(r - (int64_t)n * H);
    const float* xrow = x + (((int64_t)n * H + yi) * W) * ldx + ci;
    // dy rows oy = yi - ky + 1  (ky = 0,1,2)
    const float* drow[3];
    bool dok[3];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int oy = yi - ky + 1;
      dok[ky] = oy >= 0 && oy < H;
      drow[ky] = dy + (((int64_t)n * H + (dok[ky] ? oy : 0)) * W) * lddy;
    }
    // window columns: ox = xi - kx + 1 ; win[ky][j] holds dy at column (xi - 1 + j), j = 0..2  => kx = 2 - j
    float4 win[3][3];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      win[ky][0] = make_float4(0.f, 0.f, 0.f, 0.f);
      win[ky][1] = dok[ky] ? ldg4(drow[ky]) : make_float4(0.f, 0.f, 0.f, 0.f);            // column 0
      win[ky][2] = (dok[ky] && W > 1) ? ldg4(drow[ky] + lddy) : make_float4(0.f, 0.f, 0.f, 0.f);  // column 1
    }
    // at step xi the window must hold columns xi-1, xi, xi+1: initial state is for xi = 0
    for (int xi = 0; xi < W; ++xi) {
      const float xv = live ? __ldg(xrow + (int64_t)xi * ldx) : 0.f;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int kx = 2 - j;
          acc[ky * 3 + kx][0] = fmaf(win[ky][j].x, xv, acc[ky * 3 + kx][0]);
          acc[ky * 3 + kx][1] = fmaf(win[ky][j].y, xv, acc[ky * 3 + kx][1]);
          acc[ky * 3 + kx][2] = fmaf(win[ky][j].z, xv, acc[ky * 3 + kx][2]);
        }
      // slide: drop column xi-1, fetch column xi+2
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        win[ky][0] = win[ky][1];
        win[ky][1] = win[ky][2];
        win[ky][2] = (dok[ky] && xi + 2 < W) ? ldg4(drow[ky] + (int64_t)(xi + 2) * lddy) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  if (live) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
      for (int co = 0; co < Cout; ++co) atomicAdd(dw + ((int64_t)t * Cout + co) * Cin + ci, acc[t][co]);
  }
}

// ---------------------------------------------------------------- weight gradient
// dw[t][a][b] += sum_m S[m][a] * Bg[pix(m,t)][b]  (S = "small-grid" tensor, Bg = strided-gather tensor).
struct WgradGeom {
  int N, GH, GW;        // small grid
  int Ca, lda;          // small tensor channels / stride
  int BH, BW, Cb, ldb;  // big tensor extent / channels / stride
  int s;                // big pixel = g*s + d[t]
  int ntaps, Cb_pad;    // dw[t][Ca][Cb_pad]
  int splits;           // split-K factor over pixels
  signed char dy[16], dx[16];
};

__global__ void __launch_bounds__(256)
conv_wgrad_simt_kernel(WgradGeom g, const float* __restrict__ S, const float* __restrict__ Bg,
                       float* __restrict__ dw) {
  constexpr int BA = 64, BB = 64, BK = 16;
  __shared__ __align__(16) float Sa[2][BK][BA];
  __shared__ __align__(16) float Sb[2][BK][BB];
  const int tid = threadIdx.x;
  const int t = blockIdx.z / g.splits, split = blockIdx.z - t * g.splits;
  const int a0 = blockIdx.x * BA, b0 = blockIdx.y * BB;
  const int64_t GP = (int64_t)g.GH * g.GW;
  const int64_t M = (int64_t)g.N * GP;
  const int64_t per = ((M + g.splits - 1) / g.splits + BK - 1) / BK * BK;
  const int64_t mb = (int64_t)split * per;
  const int64_t me = mb + per < M ? mb + per : M;
  const int dyt = g.dy[t], dxt = g.dx[t];
  // load mapping: 16 pixels x 64 channels = 256 float4, one per thread for each operand
  const int lrow = tid >> 4, lc = (tid & 15) * 4;
  float4 ra, rb;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  auto load_tile = [&](int64_t mk) {
    const int64_t m = mk + lrow;
    ra = zero4; rb = zero4;
    if (m < me) {
      const int n = (int)(m / GP);
      const int r = (int)(m - (int64_t)n * GP);
      const int gy = r / g.GW, gx = r - gy * g.GW;
      if (a0 + lc < g.Ca) ra = ldg4(S + m * g.lda + a0 + lc);
      const int by = gy * g.s + dyt, bx = gx * g.s + dxt;
      if (by >= 0 && by < g.BH && bx >= 0 && bx < g.BW && b0 + lc < g.Cb)
        rb = ldg4(Bg + (((int64_t)n * g.BH + by) * g.BW + bx) * g.ldb + b0 + lc);
    }
  };
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  if (mb < me) {
    load_tile(mb);
    int buf = 0;
    for (int64_t mk = mb; mk < me; mk += BK) {
      *reinterpret_cast<float4*>(&Sa[buf][lrow][lc]) = ra;
      *reinterpret_cast<float4*>(&Sb[buf][lrow][lc]) = rb;
      __syncthreads();
      if (mk + BK < me) load_tile(mk + BK);
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 av = *reinterpret_cast<const float4*>(&Sa[buf][kk][ty * 4]);
        const float4 bv = *reinterpret_cast<const float4*>(&Sb[buf][kk][tx * 4]);
        const float a[4] = {av.x, av.y, av.z, av.w};
        const float b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      buf ^= 1;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int a = a0 + ty * 4 + i;
    if (a >= g.Ca) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int b = b0 + tx * 4 + j;
      if (b < g.Cb) atomicAdd(dw + ((int64_t)t * g.Ca + a) * g.Cb_pad + b, acc[i][j]);
    }
  }
}

// Narrow wgrad: the small-channel operand has <= 4 channels (final conv: dy has 3; PatchGAN head: dy has 1).
// dw[t][a][b] += sum_m S[m][a] * Bg[pix(m,t)][b],  a < Ca <= 4, b over blockDim-strided channels.
// `a_is_small`: 1 => S is the narrow tensor; 0 => Bg is narrow (not needed for this network).
__global__ void __launch_bounds__(256)
conv_wgrad_narrow_kernel(WgradGeom g, const float* __restrict__ S, const float* __restrict__ Bg,
                         float* __restrict__ dw) {
  const int b = blockIdx.y * blockDim.x + threadIdx.x;
  const int64_t GP = (int64_t)g.GH * g.GW;
  const int64_t M = (int64_t)g.N * GP;
  const int64_t per = (M + gridDim.x - 1) / gridDim.x;
  const int64_t mb = (int64_t)blockIdx.x * per;
  const int64_t me = mb + per < M ? mb + per : M;
  float acc[16][4];
#pragma unroll
  for (int t = 0; t < 16; ++t)
#pragma unroll
    for (int a = 0; a < 4; ++a) acc[t][a] = 0.f;
  if (b < g.Cb) {
    for (int64_t m = mb; m < me; ++m) {
      const int n = (int)(m / GP);
      const int r = (int)(m - (int64_t)n * GP);
      const int gy = r / g.GW, gx = r - gy * g.GW;
      float sv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) sv[a] = a < g.Ca ? __ldg(S + m * g.lda + a) : 0.f;
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        if (t < g.ntaps) {
          const int by = gy * g.s + g.dy[t], bx = gx * g.s + g.dx[t];
          if (by >= 0 && by < g.BH && bx >= 0 && bx < g.BW) {
            const float v = __ldg(Bg + (((int64_t)n * g.BH + by) * g.BW + bx) * g.ldb + b);
#pragma unroll
            for (int a = 0; a < 4; ++a) acc[t][a] = fmaf(sv[a], v, acc[t][a]);
          }
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 16; ++t)
      if (t < g.ntaps)
#pragma unroll
        for (int a = 0; a < 4; ++a)
          if (a < g.Ca) atomicAdd(dw + ((int64_t)t * g.Ca + a) * g.Cb_pad + b, acc[t][a]);
  }
}

__global__ void bias_grad_kernel(const float* __restrict__ dy, int ld, int64_t pixels, int C,
                                 float* __restrict__ db) {
  // blockDim.x = 32 channels-lanes x 8 pixel-lanes; grid.x over channel groups, grid.y over pixel chunks
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int pl = threadIdx.x >> 5;
  const int64_t per = (pixels + gridDim.y - 1) / gridDim.y;
  const int64_t pb = (int64_t)blockIdx.y * per;
  const int64_t pe = pb + per < pixels ? pb + per : pixels;
  float s = 0.f;
  if (c < C)
    for (int64_t p = pb + pl; p < pe; p += 8) s += __ldg(dy + p * ld + c);
  __shared__ float red[8][32];
  red[pl][threadIdx.x & 31] = s;
  __syncthreads();
  if (pl == 0 && c < C) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += red[i][threadIdx.x & 31];
    atomicAdd(db + c, tot);
  }
}

// Host-side dispatch for the SIMT path --------------------------------------------------------------
int conv_forward_simt(const ptk_conv_geom& c, const float* x, const float* w_t, int cout_pad,
                      const float* bias, int act, float* y, float* y_nchw, cudaStream_t st) {
  PTK_REQUIRE(c.Cin % 4 == 0 && c.ldx % 4 == 0, "conv: Cin (%d) and ldx (%d) must be multiples of 4", c.Cin, c.ldx);
  PTK_REQUIRE(cout_pad % 4 == 0, "conv: packed Cout (%d) must be a multiple of 4", cout_pad);
  PTK_REQUIRE(c.k * c.k <= 16, "conv: k*k > 16 unsupported");
  PTK_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_t) & 15) == 0,
              "conv: x / w must be 16-byte aligned");
  const int phases = c.transposed ? c.stride : 1;
  for (int py = 0; py < phases; ++py)
    for (int px = 0; px < phases; ++px) {
      TapGeom g;
      make_geom(c, py, px, cout_pad, &g);
      const int64_t M = (int64_t)g.N * g.GH * g.GW;
      if (M == 0) continue;
      if (c.Cout <= 4 && !c.transposed && c.k == 3 && c.stride == 1 && c.pad == 1 && c.Cin % 16 == 0 &&
          9 * c.Cin * 4 <= kNarrowMaxW && c.N <= 65535) {
        // weights [tap][Cin][cout_pad(=4)] -> constant memory (device-to-device, stream ordered)
        PTK_REQUIRE(cout_pad == 4, "k3 narrow conv: packed Cout must be 4");
        cudaError_t ce = cudaMemcpyToSymbolAsync(c_narrow_w, w_t, sizeof(float) * 9 * c.Cin * 4, 0, cudaMemcpyDeviceToDevice, st);
        if (ce != cudaSuccess) return fail(3, "cudaMemcpyToSymbolAsync: %s", cudaGetErrorString(ce));
        dim3 grid((c.W + 15) / 16, (c.H + 15) / 16, c.N);
        conv_k3_narrow_fwd_kernel<<<grid, 256, 0, st>>>(x, c.ldx, c.Cin, c.H, c.W, c.Cout, bias, act, y, c.ldy, y_nchw);
        PTK_LAUNCH_CHECK("conv_k3_narrow_fwd_kernel");
      } else if (c.Cout <= 4) {
        const int smem = g.ntaps * g.Cin * g.Cout * (int)sizeof(float);
        PTK_REQUIRE(smem <= 200 * 1024, "narrow conv: weights do not fit in shared memory");
        cudaFuncSetAttribute(conv_narrow_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        int blocks = (int)((M + 7) / 8);
        const int cap = num_sms() * 8;
        if (blocks > cap) blocks = cap;
        conv_narrow_fwd_kernel<<<blocks, 256, smem, st>>>(g, x, w_t, bias, act, y, y_nchw);
        PTK_LAUNCH_CHECK("conv_narrow_fwd_kernel");
      } else {
        PTK_REQUIRE(y_nchw == nullptr, "conv: NCHW copy only supported for Cout <= 4");
        const unsigned gm = (unsigned)((M + 127) / 128);
        if (c.Cout > 64) {
          dim3 grid(gm, (c.Cout + 127) / 128);
          conv_gemm_simt_kernel<128><<<grid, 256, 0, st>>>(g, x, w_t, bias, act, y);
        } else {
          dim3 grid(gm, (c.Cout + 63) / 64);
          conv_gemm_simt_kernel<64><<<grid, 256, 0, st>>>(g, x, w_t, bias, act, y);
        }
        PTK_LAUNCH_CHECK("conv_gemm_simt_kernel");
      }
    }
  return 0;
}

int conv_wgrad_simt(const ptk_conv_geom& c, const float* x, const float* dy, float* dw, cudaStream_t st) {
  WgradGeom g;
  const float *S, *Bg;
  int cb_pad;
  if (!c.transposed) {  // small = dy (OH x OW), big = x ; dw[t][Cout][Cin]
    g.N = c.N; g.GH = c.OH; g.GW = c.OW; g.Ca = c.Cout; g.lda = c.ldy;
    g.BH = c.H; g.BW = c.W; g.Cb = c.Cin; g.ldb = c.ldx; S = dy; Bg = x; cb_pad = c.Cin;
  } else {              // small = x (H x W), big = dy ; dw[t][Cin][Cout]
    g.N = c.N; g.GH = c.H; g.GW = c.W; g.Ca = c.Cin; g.lda = c.ldx;
    g.BH = c.OH; g.BW = c.OW; g.Cb = c.Cout; g.ldb = c.ldy; S = x; Bg = dy; cb_pad = c.Cout;
  }
  g.s = c.stride; g.Cb_pad = cb_pad; g.ntaps = c.k * c.k;
  PTK_REQUIRE(g.ntaps <= 16, "wgrad: k*k > 16 unsupported");
  for (int kh = 0; kh < c.k; ++kh)
    for (int kw = 0; kw < c.k; ++kw) {
      g.dy[kh * c.k + kw] = (signed char)(kh - c.pad);
      g.dx[kh * c.k + kw] = (signed char)(kw - c.pad);
    }
  const int64_t M = (int64_t)g.N * g.GH * g.GW;
  if (g.Ca <= 3 && !c.transposed && c.k == 3 && c.stride == 1 && c.pad == 1 && c.ldy % 4 == 0 &&
      (reinterpret_cast<uintptr_t>(dy) & 15) == 0) {
    const int64_t rows = (int64_t)c.N * c.H;
    int rpb = (int)((rows + (int64_t)num_sms() * 4 - 1) / ((int64_t)num_sms() * 4));
    if (rpb < 1) rpb = 1;
    dim3 grid((unsigned)((rows + rpb - 1) / rpb), (c.Cin + 255) / 256);
    conv_k3_narrow_wgrad_kernel<<<grid, 256, 0, st>>>(x, c.ldx, c.Cin, dy, c.ldy, c.N, c.H, c.W, c.Cout, rpb, dw);
    PTK_LAUNCH_CHECK("conv_k3_narrow_wgrad_kernel");
    return 0;
  }
  if (g.Ca <= 4) {
    g.splits = 1;
    int chunks = (int)((M + 15) / 16);
    const int cap = num_sms() * 4;
    if (chunks > cap) chunks = cap;
    if (chunks < 1) chunks = 1;
    const int threads = g.Cb >= 256 ? 256 : ((g.Cb + 31) / 32) * 32;
    dim3 grid(chunks, (g.Cb + threads - 1) / threads);
    conv_wgrad_narrow_kernel<<<grid, threads, 0, st>>>(g, S, Bg, dw);
    PTK_LAUNCH_CHECK("conv_wgrad_narrow_kernel");
    return 0;
  }
  PTK_REQUIRE(g.Ca % 4 == 0 && g.lda % 4 == 0 && g.Cb % 4 == 0 && g.ldb % 4 == 0,
              "wgrad: channel counts / strides must be multiples of 4 (Ca=%d lda=%d Cb=%d ldb=%d)", g.Ca, g.lda, g.Cb, g.ldb);
  const int ta = (g.Ca + 63) / 64, tb = (g.Cb + 63) / 64;
  int splits = (num_sms() * 4 + ta * tb * g.ntaps - 1) / (ta * tb * g.ntaps);
  const int64_t max_splits = (M + 255) / 256;
  if (splits > max_splits) splits = (int)max_splits;
  if (splits < 1) splits = 1;
  if (g.ntaps * splits > 65535) splits = 65535 / g.ntaps;
  g.splits = splits;
  dim3 grid(ta, tb, g.ntaps * splits);
  conv_wgrad_simt_kernel<<<grid, 256, 0, st>>>(g, S, Bg, dw);
  PTK_LAUNCH_CHECK("conv_wgrad_simt_kernel");
  return 0;
}

}  // namespace ptk

extern "C" int ptk_bias_grad(const float* dy, int ld, int64_t pixels, int C, float* dbias, void* stream) {
  if (pixels <= 0 || C <= 0) return 0;
  int chunks = (int)((pixels + 511) / 512);
  if (chunks > 512) chunks = 512;
  dim3 grid((C + 31) / 32, chunks);
  ptk::bias_grad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dy, ld, pixels, C, dbias);
  PTK_LAUNCH_CHECK("bias_grad_kernel");
  return 0;
}
