// torch.optim.Adam(lr, betas=(0.5,0.999)) (models/pose_gan.py:49-51) as ONE launch over a flat parameter
// arena (params, grads, exp_avg, exp_avg_sq are four flat fp32 buffers of equal length).
#include "common.cuh"

namespace ptk {
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
            float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt, float gscale) {
  const float step_size = lr / bc1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = m[i] + (gi - m[i]) * (1.f - b1);           // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = v[i] * b2 + (1.f - b2) * gi * gi;          // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= step_size * (mi / denom);
  }
}
}  // namespace ptk

extern "C" int ptk_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                             float beta2, float eps, int step, float grad_scale, void* stream) {
  PTK_REQUIRE(n > 0 && step >= 1, "adam: bad arguments");
  const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
  int64_t blocks = (n + 255) / 256;
  if (blocks > (int64_t)ptk::num_sms() * 16) blocks = (int64_t)ptk::num_sms() * 16;
  ptk::adam_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, (float)bc1,
                                                                     (float)sqrt(bc2), grad_scale);
  PTK_LAUNCH_CHECK("adam_kernel");
  return 0;
}
