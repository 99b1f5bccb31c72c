// sampling.cu -- K1 stratified sampling and K2 inverse-CDF importance sampling + merge.
// Both are HBM-bound, warp-level kernels (no tensor cores): K1 is elementwise, K2 assigns one
// warp per ray (segmented prefix sum over the ray's weights, binary search per fine sample,
// in-shared-memory bitonic merge of coarse + fine depths).
#include "common.cuh"
#include "../../include/nerf_b200.h"

namespace nerf {

// ---- K1 -------------------------------------------------------------------------------
// t_k follows torch.linspace's two-sided formula so deterministic depths match bit for bit.
__device__ __forceinline__ float linspace_at(int k, int n, float start, float end, float step) {
  return (k < n / 2) ? __fadd_rn(start, __fmul_rn(step, (float)k)) : __fsub_rn(end, __fmul_rn(step, (float)(n - k - 1)));
}

__global__ void __launch_bounds__(256) stratified_kernel(float* __restrict__ z, const float* __restrict__ u, int64_t total,
                                                         int n_samples, float near_plane, float far_plane) {
  const float step = n_samples > 1 ? (far_plane - near_plane) / (float)(n_samples - 1) : 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % n_samples);
    const float t = n_samples > 1 ? linspace_at(k, n_samples, near_plane, far_plane, step) : near_plane;
    if (u == nullptr) {
      z[i] = t;
      continue;
    }
    float lo = t, hi = t;
    if (k > 0) lo = __fmul_rn(0.5f, __fadd_rn(t, linspace_at(k - 1, n_samples, near_plane, far_plane, step)));
    if (k < n_samples - 1) hi = __fmul_rn(0.5f, __fadd_rn(linspace_at(k + 1, n_samples, near_plane, far_plane, step), t));
    z[i] = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), __ldg(u + i)));
  }
}

// ---- K2 -------------------------------------------------------------------------------
constexpr int kWarpsPerBlock = 4;

__global__ void __launch_bounds__(kWarpsPerBlock * 32) importance_kernel(
    float* __restrict__ z_merged, float* __restrict__ z_fine_out, const float* __restrict__ z_coarse,
    const float* __restrict__ w_coarse, const float* __restrict__ u, int n_rays, int nc, int nf, int s_pad) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ray = blockIdx.x * kWarpsPerBlock + warp;
  if (ray >= n_rays) return;  // whole warp exits together; only __syncwarp is used below
  const int nb = nc - 1;      // number of bin edges == number of cdf entries
  float* cdf = sm + warp * (2 * nb + s_pad);
  float* edges = cdf + nb;
  float* zall = edges + nb;
  const float* zc = z_coarse + (int64_t)ray * nc;
  const float* wc = w_coarse + (int64_t)ray * nc;

  // bin edges (midpoints) and coarse depths into the merge buffer
  for (int j = lane; j < nc; j += 32) {
    const float zj = __ldg(zc + j);
    zall[j] = zj;
    if (j < nb) edges[j] = __fmul_rn(0.5f, __fadd_rn(zj, __ldg(zc + j + 1)));
  }
  // pdf over the nc-2 interior weights.  The normaliser is a warp reduction; the running sum is
  // taken sequentially by one lane in the reference's order (torch.cumsum), because samples that
  // land in near-empty bins amplify cdf rounding by 1/pdf (SURVEY hard part 7).
  const int nv = nc - 2;
  float part = 0.f;
  for (int j = lane; j < nv; j += 32) part += __fadd_rn(__ldg(wc + j + 1), 1e-5f);
  const float total = warp_sum(part);
  if (lane == 0) {
    float run = 0.f;
    cdf[0] = 0.f;
    for (int j = 0; j < nv; ++j) {
      run = __fadd_rn(run, __fdiv_rn(__fadd_rn(__ldg(wc + j + 1), 1e-5f), total));
      cdf[j + 1] = run;
    }
  }
  __syncwarp();

  // invert the cdf for every fine sample
  const float inv_nf1 = nf > 1 ? 1.f / (float)(nf - 1) : 0.f;
  for (int j = lane; j < nf; j += 32) {
    float uj;
    if (u != nullptr) {
      uj = __ldg(u + (int64_t)ray * nf + j);
    } else {  // torch.linspace(0, 1, nf), two-sided
      const float st = inv_nf1;
      uj = (j < nf / 2) ? __fmul_rn(st, (float)j) : __fsub_rn(1.f, __fmul_rn(st, (float)(nf - j - 1)));
    }
    // first index with cdf > u  (searchsorted right=True); cdf[0] = 0 <= u
    int a = 0, b = nb;
    while (a < b) {
      const int mid = (a + b) >> 1;
      if (cdf[mid] <= uj) a = mid + 1; else b = mid;
    }
    const int below = max(a - 1, 0), above = min(a, nb - 1);
    const float c0 = cdf[below], c1 = cdf[above];
    float den = __fsub_rn(c1, c0);
    if (den < 1e-5f) den = 1.f;
    const float t = __fdiv_rn(__fsub_rn(uj, c0), den);
    const float e0 = edges[below], e1 = edges[above];
    const float zf = __fadd_rn(e0, __fmul_rn(t, __fsub_rn(e1, e0)));
    zall[nc + j] = zf;
    if (z_fine_out != nullptr) z_fine_out[(int64_t)ray * nf + j] = zf;
  }
  const int s = nc + nf;
  for (int j = s + lane; j < s_pad; j += 32) zall[j] = __int_as_float(0x7f800000);  // +inf padding
  __syncwarp();

  // bitonic sort of the padded buffer (ascending)
  for (int k = 2; k <= s_pad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (s_pad >> 1); t += 32) {
        const int pos = 2 * t - (t & (j - 1));
        const int partner = pos + j;
        const float x = zall[pos], y = zall[partner];
        const bool up = (pos & k) == 0;
        if ((x > y) == up) {
          zall[pos] = y;
          zall[partner] = x;
        }
      }
      __syncwarp();
    }
  }
  for (int j = lane; j < s; j += 32) z_merged[(int64_t)ray * s + j] = zall[j];
}

}  // namespace nerf

extern "C" int nerf_sample_stratified(float* z, const float* u, int n_rays, int n_samples, float near_plane, float far_plane,
                                      void* stream) {
  using namespace nerf;
  if (n_rays <= 0) return 0;
  NERF_CHECK_ARG(z != nullptr && n_samples >= 1, "sample_stratified: bad arguments");
  const int64_t total = (int64_t)n_rays * n_samples;
  const int64_t want = (total + 255) / 256, cap = (int64_t)kNumSMs * 16;
  const int blocks = (int)(want < cap ? want : cap);
  stratified_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(z, u, total, n_samples, near_plane, far_plane);
  NERF_CHECK_LAUNCH("stratified_kernel");
  return 0;
}

extern "C" int nerf_sample_importance(float* z_merged, float* z_fine, const float* z_coarse, const float* w_coarse,
                                      const float* u, int n_rays, int n_coarse, int n_fine, void* stream) {
  using namespace nerf;
  if (n_rays <= 0) return 0;
  NERF_CHECK_ARG(z_merged && z_coarse && w_coarse, "sample_importance: null pointer");
  NERF_CHECK_ARG(n_coarse >= 3 && n_coarse <= 512, "sample_importance: n_coarse must be in [3,512], got %d", n_coarse);
  NERF_CHECK_ARG(n_fine >= 1 && n_coarse + n_fine <= 2048, "sample_importance: n_coarse+n_fine must be <= 2048");
  if (n_rays == 0) return 0;
  int s_pad = 2;
  while (s_pad < n_coarse + n_fine) s_pad <<= 1;
  const size_t smem = sizeof(float) * kWarpsPerBlock * (2 * (n_coarse - 1) + s_pad);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(importance_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    NERF_CHECK_ARG(e == cudaSuccess, "sample_importance: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  }
  const int blocks = (n_rays + kWarpsPerBlock - 1) / kWarpsPerBlock;
  importance_kernel<<<blocks, kWarpsPerBlock * 32, smem, static_cast<cudaStream_t>(stream)>>>(
      z_merged, z_fine, z_coarse, w_coarse, u, n_rays, n_coarse, n_fine, s_pad);
  NERF_CHECK_LAUNCH("importance_kernel");
  return 0;
}
