// mlp_layout.cuh -- static geometry of the fused NeRF MLP: chain stages, weight-image layout,
// activation-stash layout.  Shared by the pack, forward, backward and wgrad kernels.
//
// Forward chain (one 128-sample tile walks all stages without leaving the SM):
//   stage 0      : enc(x) [K=64]            -> h0   (N=256)
//   stage 1..4   : h_{l-1} [K=256]          -> h_l
//   stage 5      : [h4, enc(x)] [K=256+64]  -> h5
//   stage 6, 7   : h_{l-1}                  -> h_l       (stage 7 also yields sigma on CUDA cores)
//   stage 8 (F)  : h7                       -> f  (no activation)
//   stage 9 (C0) : [f, enc(dir)] [K=256+32] -> g  (N=128); rgb = sigmoid(W_c1 g) on CUDA cores
// Backward chain (dgrad): C0 -> F -> L7 -> ... -> L1 (10 stages with transposed weights).
#pragma once
#include <cstdint>

namespace nerf {

constexpr int kTile = 128;
constexpr uint32_t kPanelBytes256 = 256 * 128;  // weight panel with 256 rows
constexpr uint32_t kPanelBytes128 = 128 * 128;  // 128-row panel (activations, C0 weights)
constexpr uint32_t kActBytes = 4 * kPanelBytes128;  // one 128 x 256 activation tile image

// ---- forward weight image: panels in consumption order --------------------------------
constexpr int kFwdStages = 10;
__host__ __device__ constexpr int fwd_panels(int stage) { return stage == 0 ? 1 : ((stage == 5 || stage == 9) ? 5 : 4); }
__host__ __device__ constexpr int fwd_first_panel(int stage) {
  // 0 | 1..16 | 17..21 | 22..29 | 30..33 | 34..38
  return stage == 0 ? 0 : (stage <= 5 ? 1 + 4 * (stage - 1) : (stage <= 8 ? 22 + 4 * (stage - 6) : 34));
}
constexpr int kFwdPanels256 = 34;  // stages 0..8
constexpr int kFwdPanels128 = 5;   // stage 9
constexpr uint32_t kFwdImageBytes = kFwdPanels256 * kPanelBytes256 + kFwdPanels128 * kPanelBytes128;
__host__ __device__ constexpr uint32_t fwd_panel_offset(int panel) {
  return panel < kFwdPanels256 ? panel * kPanelBytes256 : kFwdPanels256 * kPanelBytes256 + (panel - kFwdPanels256) * kPanelBytes128;
}
__host__ __device__ constexpr uint32_t fwd_panel_bytes(int panel) { return panel < kFwdPanels256 ? kPanelBytes256 : kPanelBytes128; }

// ---- backward (dgrad) weight image: W^T panels, 256 rows (input feature k) each ---------
// dgrad stage j: 0 = C0 (K'=128 -> 2 panels), 1 = F, 2..8 = L7..L1 (K'=256 -> 4 panels)
constexpr int kBwdStages = 9;
__host__ __device__ constexpr int bwd_panels(int stage) { return stage == 0 ? 2 : 4; }
__host__ __device__ constexpr int bwd_first_panel(int stage) { return stage == 0 ? 0 : 2 + 4 * (stage - 1); }
constexpr int kBwdPanels = 34;
constexpr uint32_t kBwdImageOffset = kFwdImageBytes;
constexpr uint32_t kBwdImageBytes = kBwdPanels * kPanelBytes256;
constexpr uint32_t kPackedBytes = kFwdImageBytes + kBwdImageBytes;

// ---- activation stash written by the training forward (per tile, region-major) -----------
// regions: ENC (x encoding, 1 panel), H0..H7 (4 panels each), F (4), DIR (1), G (2),
//          MASK (9 x 32 B per row: ReLU bits of h0..h7 (256 each) and, in slot 8, of g (128 bits, first 16 bytes))
enum StashRegion { kStashEnc = 0, kStashH0 = 1, kStashF = 9, kStashDir = 10, kStashG = 11, kStashMask = 12, kStashRegions = 13 };
__host__ __device__ constexpr uint32_t stash_region_tile_bytes(int r) {
  return r == kStashEnc || r == kStashDir ? kPanelBytes128 : (r == kStashG ? 2 * kPanelBytes128 : (r == kStashMask ? 9 * 128 * 32 : kActBytes));
}
__host__ __device__ constexpr uint64_t stash_tile_bytes_total() {
  uint64_t t = 0;
  for (int r = 0; r < kStashRegions; ++r) t += stash_region_tile_bytes(r);
  return t;
}
// byte offset of region r for `n_tiles` tiles
__host__ __device__ constexpr uint64_t stash_region_offset(int r, uint64_t n_tiles) {
  uint64_t t = 0;
  for (int i = 0; i < r; ++i) t += stash_region_tile_bytes(i) * n_tiles;
  return t;
}

// ---- gradient stash written by the dgrad chain (per tile, region-major) -------------------
// regions: DC0 (dL/dg pre-activation, 2 panels), DF (dL/df, 4), D7..D0 (dL/d pre-activation of layer l, 4 each),
//          DHEAD (1 panel: cols 0..2 = dL/d rgb pre-sigmoid, col 3 = dL/d sigma_raw, rest 0)
enum GradRegion { kGradC0 = 0, kGradF = 1, kGradL7 = 2, kGradL0 = 9, kGradHead = 10, kGradRegions = 11 };
__host__ __device__ constexpr uint32_t grad_region_tile_bytes(int r) {
  return r == kGradC0 ? 2 * kPanelBytes128 : (r == kGradHead ? kPanelBytes128 : kActBytes);
}
__host__ __device__ constexpr uint64_t grad_tile_bytes_total() {
  uint64_t t = 0;
  for (int r = 0; r < kGradRegions; ++r) t += grad_region_tile_bytes(r);
  return t;
}
__host__ __device__ constexpr uint64_t grad_region_offset(int r, uint64_t n_tiles) {
  uint64_t t = 0;
  for (int i = 0; i < r; ++i) t += grad_region_tile_bytes(i) * n_tiles;
  return t;
}

}  // namespace nerf
