// Shared host/device helpers for the ptk kernel library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "ptk.h"

namespace ptk {

extern thread_local char g_err[512];
extern std::atomic<int64_t> g_launches;

int fail(int code, const char* fmt, ...);

#define PTK_REQUIRE(cond, ...)                                  \
  do {                                                          \
    if (!(cond)) return ::ptk::fail(2, __VA_ARGS__);            \
  } while (0)

// Call after every kernel launch: counts it and converts launch errors into a return code.
#define PTK_LAUNCH_CHECK(name)                                                    \
  do {                                                                            \
    ::ptk::g_launches.fetch_add(1, std::memory_order_relaxed);                    \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess) return ::ptk::fail(3, "%s: %s", name, cudaGetErrorString(e__)); \
  } while (0)

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case PTK_ACT_LEAKY: return v > 0.f ? v : 0.2f * v;
    case PTK_ACT_RELU: return v > 0.f ? v : 0.f;
    case PTK_ACT_TANH: return tanhf(v);
    case PTK_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    default: return v;
  }
}

// derivative of LeakyReLU / ReLU expressed through the stored activated value `a`
// (torch: grad * (x > 0 ? 1 : slope); sign(a) == sign(x)).
__device__ __forceinline__ float act_grad_from_output(float a, int act) {
  if (act == PTK_ACT_LEAKY) return a > 0.f ? 1.f : 0.2f;
  if (act == PTK_ACT_RELU) return a > 0.f ? 1.f : 0.f;
  return 1.f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- programmatic dependent launch
// pdl_trigger(): the next kernel of the stream (if it was launched with the programmatic-serialization attribute) may
// start occupying SM slots as soon as every CTA of this grid has passed this point or exited -- i.e. while the last wave
// of this grid is still running.  pdl_wait(): block until the preceding grid has completed and its writes are visible;
// everything before it (barrier init, TMEM allocation, descriptor prefetch) overlaps the predecessor's tail.  Both are
// no-ops for launches without the attribute.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

static inline int num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

}  // namespace ptk
