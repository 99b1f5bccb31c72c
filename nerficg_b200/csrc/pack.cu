// pack.cu -- fp32 parameters of one NeRFBlock -> fp16 UMMA weight images (see mlp_layout.cuh).
// One thread per 16-byte chunk (8 halves) of the image.  Runs once per optimiser step (2.3 MB).
#include "common.cuh"
#include "mlp_layout.cuh"
#include "tc.cuh"
#include "../../include/nerf_b200.h"

namespace nerf {

struct PanelSrc {
  int64_t base;   // float offset of the source matrix in the flat parameter buffer
  int ld;         // source row length (in-features)
  int col0;       // first source column (forward) / first source row (transposed) of this panel
  int valid;      // number of valid panel columns (rest are zero)
  int rows;       // panel rows
};

__device__ __forceinline__ PanelSrc fwd_panel_src(int panel) {
  using L = ParamLayout;
  PanelSrc s;
  s.rows = 256;
  s.valid = 64;
  if (panel == 0) {
    s = {L::kW0, 63, 0, 63, 256};
  } else if (panel < 17) {
    const int l = 1 + (panel - 1) / 4;
    s = {L::hidden_w(l), 256, 64 * ((panel - 1) % 4), 64, 256};
  } else if (panel < 22) {
    const int pp = panel - 17;
    s = {L::kW5, 319, pp < 4 ? 64 * pp : 256, pp < 4 ? 64 : 63, 256};
  } else if (panel < 30) {
    const int l = 6 + (panel - 22) / 4;
    s = {L::hidden_w(l), 256, 64 * ((panel - 22) % 4), 64, 256};
  } else if (panel < 34) {
    s = {L::kWF, 256, 64 * (panel - 30), 64, 256};
  } else {
    const int pp = panel - 34;
    s = {L::kWC0, 283, pp < 4 ? 64 * pp : 256, pp < 4 ? 64 : 27, 128};
  }
  return s;
}

// transposed panels: row = input feature k (256 rows), column c -> output neuron col0 + c
__device__ __forceinline__ PanelSrc bwd_panel_src(int panel) {
  using L = ParamLayout;
  if (panel < 2) return {L::kWC0, 283, 64 * panel, 64, 256};
  const int j = (panel - 2) / 4, pp = (panel - 2) % 4;
  if (j == 0) return {L::kWF, 256, 64 * pp, 64, 256};
  const int l = 8 - j;  // j = 1..7 -> layers 7..1
  return {L::hidden_w(l), L::hidden_in(l), 64 * pp, 64, 256};
}

__global__ void __launch_bounds__(256) pack_kernel(uint8_t* __restrict__ packed, const float* __restrict__ params,
                                                   int with_backward) {
  const uint32_t n_chunks_fwd = kFwdImageBytes / 16, n_chunks_bwd = kBwdImageBytes / 16;
  const uint32_t total = n_chunks_fwd + (with_backward ? n_chunks_bwd : 0);
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
    float v[8];
    uint32_t dst;
    if (q < n_chunks_fwd) {
      uint32_t byte = q * 16;
      int panel;
      uint32_t in_panel;
      if (byte < kFwdPanels256 * kPanelBytes256) {
        panel = byte / kPanelBytes256;
        in_panel = byte % kPanelBytes256;
      } else {
        byte -= kFwdPanels256 * kPanelBytes256;
        panel = kFwdPanels256 + byte / kPanelBytes128;
        in_panel = byte % kPanelBytes128;
      }
      const PanelSrc s = fwd_panel_src(panel);
      const uint32_t row = in_panel / 128, chunk = (in_panel % 128) / 16;  // logical (row, chunk)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = chunk * 8 + e;
        v[e] = c < s.valid ? __ldg(params + s.base + (int64_t)row * s.ld + s.col0 + c) : 0.f;
      }
      dst = fwd_panel_offset(panel) + tc::panel_chunk_offset(row, chunk);
    } else {
      const uint32_t byte = (q - n_chunks_fwd) * 16;
      const int panel = byte / kPanelBytes256;
      const uint32_t in_panel = byte % kPanelBytes256;
      const PanelSrc s = bwd_panel_src(panel);
      const uint32_t row = in_panel / 128, chunk = (in_panel % 128) / 16;  // row = input feature k
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int n = s.col0 + chunk * 8 + e;  // output neuron
        v[e] = __ldg(params + s.base + (int64_t)n * s.ld + row);
      }
      dst = kBwdImageOffset + panel * kPanelBytes256 + tc::panel_chunk_offset(row, chunk);
    }
    uint4 o;
    o.x = tc::pack_half2(v[0], v[1]);
    o.y = tc::pack_half2(v[2], v[3]);
    o.z = tc::pack_half2(v[4], v[5]);
    o.w = tc::pack_half2(v[6], v[7]);
    *reinterpret_cast<uint4*>(packed + dst) = o;
  }
}

}  // namespace nerf

extern "C" size_t nerf_mlp_packed_bytes(void) { return nerf::kPackedBytes; }

extern "C" int nerf_mlp_pack(void* packed, const float* params, int with_backward, void* stream) {
  using namespace nerf;
  NERF_CHECK_ARG(packed && params, "mlp_pack: null pointer");
  NERF_CHECK_ARG((reinterpret_cast<uintptr_t>(packed) & 127) == 0, "mlp_pack: packed buffer must be 128-byte aligned");
  pack_kernel<<<kNumSMs * 2, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<uint8_t*>(packed), params, with_backward);
  NERF_CHECK_LAUNCH("pack_kernel");
  return 0;
}
