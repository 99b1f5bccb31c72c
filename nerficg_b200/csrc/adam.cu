// adam.cu -- K7: Adam on the flat fp32 parameter buffers (reference src/Methods/NeRF/Trainer.py:32-37 uses
// torch.optim.Adam(lr=1.0, betas=(0.9, 0.999), eps=1e-8) x LambdaLR).
// torch's fused multi-tensor Adam needs 70 us per 48-tensor block (it is bound by its per-tensor chunk table);
// on the flat buffers the same update is one 7 MB stream: read p, g, m, v, write p, m, v (28 B per parameter).
// The step counter and the bias corrections live on the device so a captured CUDA graph advances them on replay.
#include "common.cuh"
#include "../../include/nerf_b200.h"

namespace nerf {

// state[0] = step (as float, like torch's capturable Adam), state[1] = 1 - beta1^step, state[2] = sqrt(1 - beta2^step)
__global__ void adam_tick_kernel(float* __restrict__ state, float beta1, float beta2) {
  const float step = state[0] + 1.f;
  state[0] = step;
  state[1] = (float)(1.0 - pow((double)beta1, (double)step));
  state[2] = (float)sqrt(1.0 - pow((double)beta2, (double)step));
}

// kZero: the gradient buffer is cleared behind the read (optimizer.zero_grad() of the reference's iteration, Trainer.py:62,
// folded into the same pass: the captured step loses its two memset nodes); grad_mult scales the gradient on the fly
// (1/world_size of the data-parallel mean, so the all-reduce is a plain SUM with no separate scaling launch).
template <bool kZero>
__global__ void __launch_bounds__(256) adam_update_kernel(float4* __restrict__ p, float4* __restrict__ m, float4* __restrict__ v,
                                                          float4* __restrict__ g, const float* __restrict__ lr,
                                                          const float* __restrict__ state, float beta1, float beta2, float eps,
                                                          float grad_mult, int64_t n4) {
  const float step_size = __ldg(lr) / __ldg(state + 1);
  const float inv_bc2_sqrt = 1.f / __ldg(state + 2);
  auto upd = [&](float& pp, float& mm, float& vv, float gg) {
    mm = mm + (gg - mm) * (1.f - beta1);            // exp_avg.lerp_(grad, 1 - beta1)
    vv = beta2 * vv + (1.f - beta2) * gg * gg;      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(vv) * inv_bc2_sqrt + eps;
    pp -= step_size * (mm / denom);
  };
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = p[i], mm = m[i], vv = v[i];
    float4 gg = g[i];
    if (kZero) g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (grad_mult != 1.f) {
      gg.x *= grad_mult;
      gg.y *= grad_mult;
      gg.z *= grad_mult;
      gg.w *= grad_mult;
    }
    upd(pp.x, mm.x, vv.x, gg.x);
    upd(pp.y, mm.y, vv.y, gg.y);
    upd(pp.z, mm.z, vv.z, gg.z);
    upd(pp.w, mm.w, vv.w, gg.w);
    p[i] = pp;
    m[i] = mm;
    v[i] = vv;
  }
}

}  // namespace nerf

extern "C" int nerf_adam_tick(float* state, float beta1, float beta2, void* stream) {
  using namespace nerf;
  NERF_CHECK_ARG(state != nullptr, "adam_tick: null state");
  adam_tick_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(state, beta1, beta2);
  NERF_CHECK_LAUNCH("adam_tick_kernel");
  return 0;
}

extern "C" int nerf_adam_update(float* params, float* exp_avg, float* exp_avg_sq, const float* grads, const float* lr,
                                const float* state, float beta1, float beta2, float eps, int64_t n, void* stream) {
  return nerf_adam_update_ex(params, exp_avg, exp_avg_sq, const_cast<float*>(grads), lr, state, beta1, beta2, eps, 1.f, 0, n, stream);
}

extern "C" int nerf_adam_update_ex(float* params, float* exp_avg, float* exp_avg_sq, float* grads, const float* lr, const float* state,
                                   float beta1, float beta2, float eps, float grad_mult, int zero_grads, int64_t n, void* stream) {
  using namespace nerf;
  if (n <= 0) return 0;
  NERF_CHECK_ARG(params && exp_avg && exp_avg_sq && grads && lr && state, "adam_update: null pointer");
  NERF_CHECK_ARG(n % 4 == 0, "adam_update: n must be a multiple of 4 (flat block buffers are), got %lld", (long long)n);
  NERF_CHECK_ARG(((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq) |
                   reinterpret_cast<uintptr_t>(grads)) & 15) == 0, "adam_update: buffers must be 16-byte aligned");
  const int64_t n4 = n / 4;
  const int64_t want = (n4 + 255) / 256, cap = (int64_t)kNumSMs * 8;
  const int grid = (int)(want < cap ? want : cap);
  if (zero_grads)
    adam_update_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<float4*>(params), reinterpret_cast<float4*>(exp_avg), reinterpret_cast<float4*>(exp_avg_sq),
        reinterpret_cast<float4*>(grads), lr, state, beta1, beta2, eps, grad_mult, n4);
  else
    adam_update_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<float4*>(params), reinterpret_cast<float4*>(exp_avg), reinterpret_cast<float4*>(exp_avg_sq),
        reinterpret_cast<float4*>(grads), lr, state, beta1, beta2, eps, grad_mult, n4);
  NERF_CHECK_LAUNCH("adam_update_kernel");
  return 0;
}
