// api.cu -- error reporting, device gate and parameter layout of the C-ABI (include/nerf_b200.h).
#include <cstdarg>
#include <cstring>

#include "common.cuh"
#include "../../include/nerf_b200.h"

namespace nerf {
static thread_local char g_error[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
static uint64_t* g_timing = nullptr;
uint64_t* timing_buffer() { return g_timing; }
}  // namespace nerf

extern "C" int nerf_debug_set_timing(void* device_buffer) {
  nerf::g_timing = static_cast<uint64_t*>(device_buffer);
  return 0;
}

extern "C" int nerf_abi_version(void) { return NERF_ABI_VERSION; }
extern "C" const char* nerf_last_error(void) { return nerf::g_error; }

extern "C" int nerf_device_check(int device) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    nerf::set_error("device_check: %s", cudaGetErrorString(e));
    return -2;
  }
  if (prop.major != 10) {
    nerf::set_error("device_check: device %d is sm_%d%d; this library only contains sm_100a code and has no fallback",
                    device, prop.major, prop.minor);
    return -1;
  }
  return 0;
}

extern "C" int nerf_param_layout(int64_t* offsets, int64_t* sizes, int64_t* total) {
  using L = nerf::ParamLayout;
  int t = 0;
  for (int l = 0; l < 8; ++l) {
    offsets[t] = L::hidden_w(l);
    sizes[t++] = 256 * L::hidden_in(l);
    offsets[t] = L::hidden_b(l);
    sizes[t++] = 256;
  }
  const int64_t rest[8][2] = {{L::kWF, 65536}, {L::kBF, 256}, {L::kWS, 256}, {L::kBS, 1},
                              {L::kWC0, 128 * 283}, {L::kBC0, 128}, {L::kWC1, 384}, {L::kBC1, 3}};
  for (int i = 0; i < 8; ++i) {
    offsets[t] = rest[i][0];
    sizes[t++] = rest[i][1];
  }
  if (total) *total = L::kTotal;
  return t == NERF_N_PARAM_TENSORS ? 0 : -1;
}
