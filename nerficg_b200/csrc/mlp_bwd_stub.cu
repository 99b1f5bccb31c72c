// temporary: backward entry points not yet implemented
#include "common.cuh"
#include "../../include/nerf_b200.h"
extern "C" size_t nerf_mlp_backward_workspace_bytes(int64_t) { return 0; }
extern "C" int nerf_mlp_backward(float*, const float*, const float*, const void*, void*, const void*, const float*, int, int, float, void*) { nerf::set_error("not implemented"); return -9; }
