// loss.cu -- K8: NeRFLoss forward and its gradient in one launch (reference src/Methods/NeRF/Loss.py:26-43,
// src/Datasets/utils.py:185-189 apply_background_color, src/Optim/Losses/utils.py:54-57 mse).
//   gt    = clamp(lerp(background, rgb_gt, alpha_gt), 0, 1)
//   loss  = lambda_c * [mse(rgb, gt) + mse(rgb_coarse, gt)] + lambda_a * [mse(alpha, alpha_gt) + mse(alpha_coarse, alpha_gt)]
//   g_rgb = 2 lambda_c (rgb - gt) / (3 n),  g_alpha = 2 lambda_a (alpha - alpha_gt) / n        (same for the coarse pass)
// The captured training step spent a dozen 3-9 us torch launches on this (lerp, clamp, sub, square, mean, mul, add...);
// here it is a single block that walks the rays once.  One block keeps the reduction order fixed (bit-reproducible loss).
#include "common.cuh"
#include "../../include/nerf_b200.h"

namespace nerf {

constexpr int kLossThreads = 1024;

// torch.lerp's two-sided formula (ATen Lerp.h), with the contractions nvcc applies to it
__device__ __forceinline__ float torch_lerp(float start, float end, float w) {
  const float diff = end - start;
  return fabsf(w) < 0.5f ? fmaf(w, diff, start) : fmaf(-diff, 1.f - w, end);
}

__global__ void __launch_bounds__(kLossThreads) loss_kernel(float* __restrict__ loss, float* __restrict__ g_rgb, float* __restrict__ g_rgb_c,
                                                            float* __restrict__ g_alpha, float* __restrict__ g_alpha_c,
                                                            const float* __restrict__ rgb, const float* __restrict__ rgb_c,
                                                            const float* __restrict__ alpha, const float* __restrict__ alpha_c,
                                                            const float* __restrict__ rgb_gt, const float* __restrict__ alpha_gt,
                                                            const float* __restrict__ background, int n, float lambda_c, float lambda_a) {
  __shared__ float red[4][kLossThreads / 32];
  const float b0 = background ? __ldg(background) : 0.f, b1 = background ? __ldg(background + 1) : 0.f,
              b2 = background ? __ldg(background + 2) : 0.f;
  const float kc = 2.f * lambda_c / (3.f * (float)n), ka = 2.f * lambda_a / (float)n;
  const bool with_alpha = lambda_a > 0.f && alpha != nullptr;
  float s_f = 0.f, s_c = 0.f, s_af = 0.f, s_ac = 0.f;
  for (int r = threadIdx.x; r < n; r += kLossThreads) {
    const float a_gt = alpha_gt ? __ldg(alpha_gt + r) : 1.f;
    float gt[3];
    gt[0] = fminf(fmaxf(torch_lerp(b0, __ldg(rgb_gt + 3 * r), a_gt), 0.f), 1.f);
    gt[1] = fminf(fmaxf(torch_lerp(b1, __ldg(rgb_gt + 3 * r + 1), a_gt), 0.f), 1.f);
    gt[2] = fminf(fmaxf(torch_lerp(b2, __ldg(rgb_gt + 3 * r + 2), a_gt), 0.f), 1.f);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float d = __ldg(rgb + 3 * r + c) - gt[c];
      s_f = fmaf(d, d, s_f);
      g_rgb[3 * r + c] = d * kc;
      if (rgb_c != nullptr) {
        const float dc = __ldg(rgb_c + 3 * r + c) - gt[c];
        s_c = fmaf(dc, dc, s_c);
        g_rgb_c[3 * r + c] = dc * kc;
      }
    }
    if (with_alpha) {
      const float d = __ldg(alpha + r) - a_gt;
      s_af = fmaf(d, d, s_af);
      g_alpha[r] = d * ka;
      if (alpha_c != nullptr) {
        const float dc = __ldg(alpha_c + r) - a_gt;
        s_ac = fmaf(dc, dc, s_ac);
        g_alpha_c[r] = dc * ka;
      }
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  s_f = warp_sum(s_f);
  s_c = warp_sum(s_c);
  s_af = warp_sum(s_af);
  s_ac = warp_sum(s_ac);
  if (lane == 0) {
    red[0][warp] = s_f;
    red[1][warp] = s_c;
    red[2][warp] = s_af;
    red[3][warp] = s_ac;
  }
  __syncthreads();
  if (warp == 0) {
    float t0 = red[0][lane], t1 = red[1][lane], t2 = red[2][lane], t3 = red[3][lane];  // kLossThreads / 32 == 32 partials each
    t0 = warp_sum(t0);
    t1 = warp_sum(t1);
    t2 = warp_sum(t2);
    t3 = warp_sum(t3);
    if (lane == 0) *loss = lambda_c * (t0 + t1) / (3.f * (float)n) + lambda_a * (t2 + t3) / (float)n;
  }
}

}  // namespace nerf

extern "C" int nerf_loss_mse(float* loss, float* g_rgb, float* g_rgb_coarse, float* g_alpha, float* g_alpha_coarse, const float* rgb,
                             const float* rgb_coarse, const float* alpha, const float* alpha_coarse, const float* rgb_gt,
                             const float* alpha_gt, const float* background, int n_rays, float lambda_color, float lambda_alpha,
                             void* stream) {
  using namespace nerf;
  NERF_CHECK_ARG(n_rays > 0, "loss_mse: n_rays must be positive");
  NERF_CHECK_ARG(loss && g_rgb && rgb && rgb_gt, "loss_mse: null pointer");
  NERF_CHECK_ARG(rgb_coarse == nullptr || g_rgb_coarse != nullptr, "loss_mse: coarse colours need a coarse gradient buffer");
  if (lambda_alpha > 0.f) {
    NERF_CHECK_ARG(alpha && g_alpha, "loss_mse: the alpha term needs alpha and its gradient buffer");
    NERF_CHECK_ARG(alpha_coarse == nullptr || g_alpha_coarse != nullptr, "loss_mse: coarse alpha needs a coarse gradient buffer");
  }
  loss_kernel<<<1, kLossThreads, 0, static_cast<cudaStream_t>(stream)>>>(loss, g_rgb, g_rgb_coarse, g_alpha, g_alpha_coarse, rgb, rgb_coarse,
                                                                         alpha, alpha_coarse, rgb_gt, alpha_gt, background, n_rays,
                                                                         lambda_color, lambda_alpha);
  NERF_CHECK_LAUNCH("loss_kernel");
  return 0;
}
