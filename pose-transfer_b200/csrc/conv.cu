// C-ABI dispatch for the convolution family: chooses the fp32 CUDA-core path or the tcgen05 TF32 path.
#include "common.cuh"

namespace ptk {
int conv_forward_simt(const ptk_conv_geom& c, const float* x, const float* w_t, int cout_pad, const float* bias, int act,
                      float* y, float* y_nchw, cudaStream_t st);
int conv_wgrad_simt(const ptk_conv_geom& c, const float* x, const float* dy, float* dw, cudaStream_t st);
bool conv_tc_supported(const ptk_conv_geom& c);
int conv_forward_tc(const ptk_conv_geom& c, const float* x, const float* w_k, const float* bias, int act, float* y,
                    double* stats, float* scratch, int64_t scratch_floats, cudaStream_t st);
bool conv_wgrad_tc_supported(const ptk_conv_geom& c);
int conv_wgrad_tc(const ptk_conv_geom& c, const float* x, const float* dy, float* dw, int64_t dw_capacity, int* nparts,
                  cudaStream_t st, bool plan_only = false);
}  // namespace ptk

using namespace ptk;

extern "C" int ptk_conv_tc_supported(const ptk_conv_geom* g) { return conv_tc_supported(*g) ? 1 : 0; }

extern "C" int ptk_conv_forward(const ptk_conv_geom* g, const float* x, const float* w_t, const float* w_k,
                                const float* bias, int act, float* y, float* y_nchw, double* stats, void* stream) {
  return ptk_conv_forward_ws(g, x, w_t, w_k, bias, act, y, y_nchw, stats, nullptr, 0, stream);
}

extern "C" int ptk_conv_forward_ws(const ptk_conv_geom* g, const float* x, const float* w_t, const float* w_k,
                                   const float* bias, int act, float* y, float* y_nchw, double* stats, float* scratch,
                                   int64_t scratch_floats, void* stream) {
  PTK_REQUIRE(g && x && (y || y_nchw), "conv_forward: null argument");
  PTK_REQUIRE(g->N > 0 && g->H > 0 && g->W > 0 && g->OH > 0 && g->OW > 0 && g->Cin > 0 && g->Cout > 0, "conv_forward: bad extents");
  PTK_REQUIRE(g->ldx >= g->Cin && (!y || g->ldy >= g->Cout), "conv_forward: ld smaller than channel count");
  cudaStream_t st = (cudaStream_t)stream;
  bool use_tc = false;
  if (g->impl == PTK_IMPL_TC) {
    PTK_REQUIRE(conv_tc_supported(*g), "conv_forward: geometry not supported by the tcgen05 path");
    use_tc = true;
  } else if (g->impl == PTK_IMPL_AUTO) {
    use_tc = w_k != nullptr && y != nullptr && y_nchw == nullptr &&
             (act == PTK_ACT_NONE || act == PTK_ACT_LEAKY || act == PTK_ACT_RELU) && conv_tc_supported(*g);
  }
  if (use_tc) {
    PTK_REQUIRE(w_k != nullptr && y != nullptr && y_nchw == nullptr, "conv_forward(tc): needs w_k and an NHWC destination");
    return conv_forward_tc(*g, x, w_k, bias, act, y, stats, scratch, scratch_floats, st);
  }
  PTK_REQUIRE(w_t != nullptr, "conv_forward(simt): w_t is NULL");
  const int cout_pad = (g->Cout + 3) & ~3;
  int rc = conv_forward_simt(*g, x, w_t, cout_pad, bias, act, y, y_nchw, st);
  if (rc) return rc;
  if (stats) {
    PTK_REQUIRE(y != nullptr && act == PTK_ACT_NONE, "conv_forward: stats need a raw NHWC output");
    rc = ptk_gn_stats(y, g->ldy, g->N, (int64_t)g->OH * g->OW, g->Cout, stats, stream);
  }
  return rc;
}

extern "C" int ptk_conv_wgrad(const ptk_conv_geom* g, const float* x, const float* dy, float* dw, void* stream) {
  PTK_REQUIRE(g && x && dy && dw, "conv_wgrad: null argument");
  if (g->impl == PTK_IMPL_TC) PTK_REQUIRE(conv_wgrad_tc_supported(*g), "conv_wgrad: geometry not supported by the tcgen05 path");
  if (g->impl != PTK_IMPL_SIMT && conv_wgrad_tc_supported(*g)) return conv_wgrad_tc(*g, x, dy, dw, 0, nullptr, (cudaStream_t)stream);
  int rc = ptk_fill(dw, (int64_t)g->k * g->k * (g->transposed ? (int64_t)g->Cin * g->Cout : (int64_t)g->Cout * g->Cin), 0.f, stream);
  if (rc) return rc;
  return conv_wgrad_simt(*g, x, dy, dw, (cudaStream_t)stream);
}

extern "C" int ptk_conv_wgrad_parts(const ptk_conv_geom* g, const float* x, const float* dy, float* dw, int64_t dw_capacity,
                                    int* nparts, void* stream) {
  PTK_REQUIRE(g && x && dy && dw && nparts, "conv_wgrad_parts: null argument");
  const int64_t part = (int64_t)g->k * g->k * (int64_t)g->Cin * g->Cout;
  PTK_REQUIRE(dw_capacity >= part, "conv_wgrad_parts: dw_capacity (%lld) smaller than one gradient (%lld floats)",
              (long long)dw_capacity, (long long)part);
  if (g->impl == PTK_IMPL_TC) PTK_REQUIRE(conv_wgrad_tc_supported(*g), "conv_wgrad: geometry not supported by the tcgen05 path");
  if (g->impl != PTK_IMPL_SIMT && conv_wgrad_tc_supported(*g))
    return conv_wgrad_tc(*g, x, dy, dw, dw_capacity, nparts, (cudaStream_t)stream);
  *nparts = 1;
  int rc = ptk_fill(dw, part, 0.f, stream);
  if (rc) return rc;
  return conv_wgrad_simt(*g, x, dy, dw, (cudaStream_t)stream);
}

extern "C" int ptk_conv_wgrad_plan(const ptk_conv_geom* g, int64_t dw_capacity, int* nparts) {
  PTK_REQUIRE(g && nparts, "conv_wgrad_plan: null argument");
  *nparts = 1;
  if (g->impl != PTK_IMPL_SIMT && conv_wgrad_tc_supported(*g))
    return conv_wgrad_tc(*g, nullptr, nullptr, nullptr, dw_capacity, nparts, nullptr, true);
  return 0;
}
