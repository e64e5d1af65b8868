// Fused Adam over ONE flat fp32 parameter buffer (all ~130 parameter tensors of the model are views into it, so the
// optimiser is a single bandwidth-bound launch instead of ~130 tiny ones; the same flat gradient buffer is what the
// data-parallel allreduce buckets slice).  Update rule identical to torch.optim.Adam (amsgrad=False, L2 weight decay)
// as the reference configures it (models/avmnist.py:413-415, cfg train.optimizer):
//   g' = g*grad_scale + wd*p ; m = b1 m + (1-b1) g' ; v = b2 v + (1-b2) g'^2
//   p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// lr and the step counter may live on the device (state_dev = {lr, step}) so that a ReduceLROnPlateau-style scheduler
// can change lr and a captured CUDA graph can replay the step without re-capturing.
#include "common.cuh"
#include "kernels.h"

namespace m2 {
namespace {

__global__ void adam_tick_kernel(float* state) { state[1] += 1.f; }

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long long n, float lr, float beta1, float beta2,
                                                   float eps, float wd, float stepf, float grad_scale,
                                                   const float* __restrict__ state) {
  if (state) { lr = state[0]; stepf = state[1]; }
  const float bc1 = 1.f - powf(beta1, stepf), bc2 = 1.f - powf(beta2, stepf);
  const float step_size = lr / bc1, rsq_bc2 = rsqrtf(bc2);
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
#define M2_ADAM1(c)                                              \
  {                                                              \
    const float gr = gg.c * grad_scale + wd * pp.c;              \
    mm.c = beta1 * mm.c + (1.f - beta1) * gr;                    \
    vv.c = beta2 * vv.c + (1.f - beta2) * gr * gr;               \
    pp.c -= step_size * mm.c / (sqrtf(vv.c) * rsq_bc2 + eps);    \
  }
    M2_ADAM1(x) M2_ADAM1(y) M2_ADAM1(z) M2_ADAM1(w)
#undef M2_ADAM1
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  // tail
  for (long long i = (n4 << 2) + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float gr = g[i] * grad_scale + wd * p[i];
    const float mi = beta1 * m[i] + (1.f - beta1) * gr, vi = beta2 * v[i] + (1.f - beta2) * gr * gr;
    m[i] = mi; v[i] = vi;
    p[i] -= step_size * mi / (sqrtf(vi) * rsq_bc2 + eps);
  }
}

}  // namespace

int adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
              float weight_decay, int step, float grad_scale, float* state_dev, cudaStream_t s) {
  if (!p || !g || !m || !v || n <= 0) return M2_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
       reinterpret_cast<uintptr_t>(v)) & 15)
    return M2_ERR_ALIGN;
  // step < 0 with a device state: read {lr, step} as they are (further ranges of a step whose first launch advanced them)
  const bool tick = state_dev && step >= 0;
  LaunchScope scope("adam", s, tick ? 2 : 1);
  if (tick) adam_tick_kernel<<<1, 1, 0, s>>>(state_dev);
  long long blocks = (n / 4 + 255) / 256;
  const int grid = static_cast<int>(blocks < 1 ? 1 : (blocks > 148 * 8 ? 148 * 8 : blocks));
  adam_kernel<<<grid, 256, 0, s>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, static_cast<float>(step), grad_scale, state_dev);
  M2_LAUNCH_CHECK();
  return M2_OK;
}

}  // namespace m2
