// rays.cu -- K0: device-side ray generation and ray-batch assembly (the step right before the path).
//   nerf_generate_rays : View.get_rays (reference src/Datasets/utils.py:1053-1074) for all pixels of a view or for a list
//                        of pixel ids: local direction ((x + .5 - cx) / fx, (y + .5 - cy) / fy, 1) taken from the two
//                        torch.linspace tables of PerspectiveCamera.compute_local_ray_directions (Cameras/Perspective.py:64-94,
//                        two-sided linspace formula), rotated by c2w (View.cam_to_world, utils.py:1033-1038), NOT normalised;
//                        view_direction = F.normalize(direction); origin = camera position.
//   nerf_gather_rays   : RayBatch.__getitem__ with an index tensor (src/Datasets/utils.py:598-613): origin, direction,
//                        view_direction, rgb (3 floats each) and alpha (1 float) of the selected rays in ONE launch
//                        (the reference / torch path issues one gather per field).
// Both are pure streaming kernels: 36 B written per generated ray, 52 B read + 52 B written per gathered ray.
#include "common.cuh"
#include "../../include/nerf_b200.h"

namespace nerf {

struct RayCamera {
  float x0, x1, y0, y1;  // linspace end points of the local x / y tables (host: double arithmetic, then float, like torch)
  float r[9];            // rotation, row-major (world <- camera)
  float t[3];            // camera position
  int width, height;
};

// torch.linspace(start, end, n)[k]: step = (end - start) / (n - 1); first half counts up from start, second half down from end
__device__ __forceinline__ float linspace_two_sided(int k, int n, float start, float end) {
  if (n == 1) return start;
  const float step = __fdiv_rn(__fsub_rn(end, start), (float)(n - 1));
  return (k < n / 2) ? __fadd_rn(start, __fmul_rn(step, (float)k)) : __fsub_rn(end, __fmul_rn(step, (float)(n - k - 1)));
}

__global__ void __launch_bounds__(256) generate_rays_kernel(float* __restrict__ origin, float* __restrict__ direction,
                                                            float* __restrict__ view_direction, const int64_t* __restrict__ pixel_ids,
                                                            int64_t n, const RayCamera cam) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = pixel_ids ? __ldg(pixel_ids + i) : i;
    const int px = (int)(pix % cam.width), py = (int)(pix / cam.width);
    const float lx = linspace_two_sided(px, cam.width, cam.x0, cam.x1);
    const float ly = linspace_two_sided(py, cam.height, cam.y0, cam.y1);
    // (lx, ly, 1) @ R^T, accumulated in index order like a 3-term dot product
    float d[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) d[j] = fmaf(1.f, cam.r[3 * j + 2], fmaf(ly, cam.r[3 * j + 1], __fmul_rn(lx, cam.r[3 * j])));
    const float norm = fmaxf(sqrtf(fmaf(d[2], d[2], fmaf(d[1], d[1], __fmul_rn(d[0], d[0])))), 1e-12f);  // F.normalize eps
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      origin[3 * i + j] = cam.t[j];
      direction[3 * i + j] = d[j];
      view_direction[3 * i + j] = __fdiv_rn(d[j], norm);
    }
  }
}

__global__ void __launch_bounds__(256) gather_rays_kernel(float* __restrict__ o_dst, float* __restrict__ d_dst, float* __restrict__ v_dst,
                                                          float* __restrict__ c_dst, float* __restrict__ a_dst,
                                                          const float* __restrict__ o_src, const float* __restrict__ d_src,
                                                          const float* __restrict__ v_src, const float* __restrict__ c_src,
                                                          const float* __restrict__ a_src, const int64_t* __restrict__ ids, int64_t n) {
  // one thread per (ray, field): fields 0..3 are 3-float rows, field 4 is the alpha scalar
  const int64_t total = n * 5;
  for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < total; w += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = w / 5;
    const int f = (int)(w - 5 * i);
    const int64_t src = __ldg(ids + i);
    const float* s = f == 0 ? o_src : (f == 1 ? d_src : (f == 2 ? v_src : (f == 3 ? c_src : a_src)));
    float* dst = f == 0 ? o_dst : (f == 1 ? d_dst : (f == 2 ? v_dst : (f == 3 ? c_dst : a_dst)));
    if (s == nullptr || dst == nullptr) continue;
    if (f == 4) {
      dst[i] = __ldg(s + src);
    } else {
      dst[3 * i] = __ldg(s + 3 * src);
      dst[3 * i + 1] = __ldg(s + 3 * src + 1);
      dst[3 * i + 2] = __ldg(s + 3 * src + 2);
    }
  }
}

}  // namespace nerf

extern "C" int nerf_generate_rays(float* origin, float* direction, float* view_direction, const int64_t* pixel_ids, int64_t n_rays,
                                  const double* c2w_host, int width, int height, double focal_x, double focal_y, double center_x,
                                  double center_y, void* stream) {
  using namespace nerf;
  if (n_rays <= 0) return 0;
  NERF_CHECK_ARG(origin && direction && view_direction && c2w_host, "generate_rays: null pointer");
  NERF_CHECK_ARG(width >= 1 && height >= 1 && focal_x != 0.0 && focal_y != 0.0, "generate_rays: bad camera");
  NERF_CHECK_ARG(pixel_ids != nullptr || n_rays == (int64_t)width * height, "generate_rays: without pixel ids n_rays must be width*height");
  RayCamera cam;
  // the reference computes these in Python floats (double) and torch.linspace casts them to float32
  cam.x0 = (float)((0.5 - center_x) / focal_x);
  cam.x1 = (float)(((double)(width - 1) + 0.5 - center_x) / focal_x);
  cam.y0 = (float)((0.5 - center_y) / focal_y);
  cam.y1 = (float)(((double)(height - 1) + 0.5 - center_y) / focal_y);
  for (int j = 0; j < 3; ++j) {  // c2w: 3x4 or 4x4 row-major with a row stride of 4 doubles; float32 like View.rotation / position
    for (int k = 0; k < 3; ++k) cam.r[3 * j + k] = (float)c2w_host[4 * j + k];
    cam.t[j] = (float)c2w_host[4 * j + 3];
  }
  cam.width = width;
  cam.height = height;
  const int64_t want = (n_rays + 255) / 256, cap = (int64_t)kNumSMs * 16;
  generate_rays_kernel<<<(int)(want < cap ? want : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(origin, direction, view_direction,
                                                                                                      pixel_ids, n_rays, cam);
  NERF_CHECK_LAUNCH("generate_rays_kernel");
  return 0;
}

extern "C" int nerf_gather_rays(float* origin_dst, float* direction_dst, float* view_direction_dst, float* rgb_dst, float* alpha_dst,
                                const float* origin_src, const float* direction_src, const float* view_direction_src,
                                const float* rgb_src, const float* alpha_src, const int64_t* ids, int64_t n_rays, void* stream) {
  using namespace nerf;
  if (n_rays <= 0) return 0;
  NERF_CHECK_ARG(ids != nullptr, "gather_rays: null index pointer");
  const int64_t want = (5 * n_rays + 255) / 256, cap = (int64_t)kNumSMs * 16;
  gather_rays_kernel<<<(int)(want < cap ? want : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      origin_dst, direction_dst, view_direction_dst, rgb_dst, alpha_dst, origin_src, direction_src, view_direction_src, rgb_src, alpha_src,
      ids, n_rays);
  NERF_CHECK_LAUNCH("gather_rays_kernel");
  return 0;
}
