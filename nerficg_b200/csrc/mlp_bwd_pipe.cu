// mlp_bwd_pipe.cu -- K4 fused: data gradients AND weight gradients of the NeRF MLP in one persistent kernel whose CTA
// pairs are LAYER-STATIONARY pipeline stages.
//
// Why.  In the tile-major chain (mlp_bwd.cu) the per-layer output gradients dY_l leave the SM for HBM (624 KB per
// 128-sample tile) and the layer-major weight-gradient kernel (mlp_wgrad.cu) reads them back together with the activation
// stash (1,424 KB per tile): 10.4 of the step's 23.1 GB are that round trip, and both kernels are bound by memory paths
// (measured, tools/l2_probe.py: an SM's store path to L2 carries 27-32 B/clk whoever issues the store, and a queued 64 KB
// image store delays the weight-ring loads queued behind it).  A weight gradient needs an accumulator per LAYER
// (256 x 256 fp32 = all 512 TMEM columns of one SM), so it cannot live inside a tile-major chain -- but it can if the CHAIN
// is laid across SMs instead of across time:
//
//   pipeline p (8 of them) = 9 CTA pairs: [head] -> [F] -> [L7] -> [L6] -> [L5] -> [L4] -> [L3] -> [L2] -> [L1]
//   head      : dL/d(rgb, sigma_raw) -> dG (CUDA cores, as the chain's prologue), dF = dG W_c0[:, :256]      (tcgen05, 2 slots)
//   stage st  : receives dY_in (dF, dY7, ... dY1), keeps ITS layer's W^T half resident in shared memory and ITS layer's
//               dW half (128 x 256 fp32 = 256 TMEM columns) resident in tensor memory for the whole kernel:
//                 dgrad  acc[256 smp x 256]  = dY_in [K-major A]  x W^T            (cta_group::2, M = 256 samples)
//                 wgrad  dW [256 x 256]     += dY_in^T [MN-major A] x H [MN-major B, activation stash]   (M = 256 features)
//               the epilogue (ReLU mask of the forward, fp16) of group t overlaps the wgrad MMAs of group t, so the tensor
//               pipe never waits for a TMEM drain; dY_out travels to the next pair through a 4-deep ring in L2.
//   Groups of 256 samples (one 128-sample tile per CTA of a pair) are dealt round-robin to the 8 pipelines; the 72 pairs
//   run concurrently (74 two-SM clusters fit on 148 SMs; 2 pairs stay idle) and synchronise through monotone
//   counters in global memory (release/acquire at gpu scope; the producer never overtakes the 4-slot ring).
//
// HBM traffic per tile: the activation stash is read ONCE (H0..H7: 512 KB + masks), and only dG, dY5, dY0 and the head
// panel (176 KB) are still written for the small residual weight-gradient kernel (mlp_wgrad.cu, jobs 8-12: encodings,
// colour head) -- instead of 624 KB written + 1,424 KB read.  Bias gradients are column sums of dY_in on CUDA cores
// (reducer warps), the density head's weight gradient is sum_m dsigma_raw[m] h7[m][:] on the same warps of stage F.
#include "common.cuh"
#include "mlp_layout.cuh"
#include "tc.cuh"
#include "../../include/nerf_b200.h"

namespace nerf {
using namespace tc;

namespace pipe {
constexpr int kThreads = 640;                 // warps 0-3 service, 4-19 epilogue / reducer / store roles
constexpr int kPipelines = 8;
constexpr int kRoles = 9;                     // head + 8 stages
constexpr int kLinks = 8;                     // link k feeds stage k + 1
constexpr int kLinkSlots = 4;
constexpr uint32_t kStageBytes = kPanelBytes128;   // 16 KB ring stage
constexpr uint32_t kHalfPanel = 8192;
constexpr int kRing = 6;
constexpr uint32_t kGroupBytes = 2 * kActBytes;    // one link slot: two tile images
// shared memory map of a stage CTA
constexpr uint32_t kOffW = 0;                                   // 4 x 16 KB: this CTA's half of the layer's W^T panels
constexpr uint32_t kOffRing = kOffW + kActBytes;                // 6 x 16 KB
constexpr uint32_t kOffOut = kOffRing + kRing * kStageBytes;    // 64 KB: dY_out image (4 panels)
constexpr uint32_t kOffBars = kOffOut + kActBytes;
constexpr uint32_t kOffDsig = kOffBars + 512;                   // 64 floats: dsigma_raw of the half tile in flight (stage F)
constexpr uint32_t kSmemBytes = kOffDsig + 512;
static_assert(kSmemBytes <= 232448, "shared memory budget exceeded");
// shared memory map of a head CTA: W (2 x 16 KB) | dG operand of slot 0, 1 (2 x 32 KB) | dF image of slot 0, 1 (2 x 64 KB)
constexpr uint32_t kHeadOffA = 2 * kPanelBytes128;
constexpr uint32_t kHeadOffOut = kHeadOffA + 2 * (2 * kPanelBytes128);
static_assert(kHeadOffOut + 2 * kActBytes <= kOffBars, "head layout overlaps the barriers");
constexpr int kRegsEpilogue = 112, kRegsOther = 32;
constexpr size_t kLinkBytes = (size_t)kPipelines * kLinks * kLinkSlots * kGroupBytes;   // 32 MB
constexpr size_t kFlagStride = 32;                                                      // uint32 per flag, 128 B apart
constexpr size_t kFlagBytes = 2 * (size_t)kPipelines * kLinks * kLinkSlots * kFlagStride * 4;
}  // namespace pipe

struct PipeParams {
  float* grads;
  const float4* d_rgbsigma;
  const float4* rgbsigma;
  const uint8_t* stash;
  uint8_t* gstash;
  uint8_t* links;
  uint32_t* flags;          // [2][pipelines][links][slots][kFlagStride]: ready, then freed
  const uint8_t* packed;
  const float* params;
  int64_t n_evals;
  int n_tiles;
  float inv_scale;
};

// ---- flags (gpu scope) ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void spin_until(const uint32_t* p, uint32_t target) {
  while (ld_acquire_gpu(p) < target) __nanosleep(64);
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// 16 columns of a dgrad epilogue (as in mlp_bwd.cu): t = acc (+ dsigma_raw * w_sigma), masked by the forward's ReLU bits
template <bool kSig>
__device__ __forceinline__ void pipe_dgrad16(const uint32_t (&v)[16], uint32_t m, int qbase, const float* __restrict__ ws, float dsr,
                                             uint32_t dst0, uint32_t dst1) {
  uint32_t w[8];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float t0 = __uint_as_float(v[4 * q + 0]), t1 = __uint_as_float(v[4 * q + 1]);
    float t2 = __uint_as_float(v[4 * q + 2]), t3 = __uint_as_float(v[4 * q + 3]);
    if (kSig) {
      const float4 wq = __ldg(reinterpret_cast<const float4*>(ws) + q);
      t0 = fmaf(dsr, wq.x, t0);
      t1 = fmaf(dsr, wq.y, t1);
      t2 = fmaf(dsr, wq.z, t2);
      t3 = fmaf(dsr, wq.w, t3);
    }
    t0 = (m & (1u << (qbase + 2 * q))) ? t0 : 0.f;
    t1 = (m & (1u << (16 + qbase + 2 * q))) ? t1 : 0.f;
    t2 = (m & (1u << (qbase + 2 * q + 1))) ? t2 : 0.f;
    t3 = (m & (1u << (16 + qbase + 2 * q + 1))) ? t3 : 0.f;
    w[2 * q] = pack_half2(t0, t1);
    w[2 * q + 1] = pack_half2(t2, t3);
  }
  st_shared_v4(dst0, w[0], w[1], w[2], w[3]);
  st_shared_v4(dst1, w[4], w[5], w[6], w[7]);
}

// accumulator (this warp's 32 lanes x 128 columns at t_acc) -> masked fp16 -> the two panels of this column half at out_h
template <bool kSig>
__device__ __forceinline__ void pipe_drain_acc(uint32_t t_acc, const uint32_t (&mk)[4], const float* __restrict__ wsp, float dsr,
                                               uint32_t out_h, uint32_t xr) {
  uint32_t va[16], vb[16];
  tmem_ld16(t_acc, va);
#pragma unroll
  for (int s = 0; s < 8; s += 2) {
    const uint32_t pbase = out_h + (uint32_t)(s >> 2) * kPanelBytes128;
    const uint32_t c0 = (uint32_t)(s & 3) * 32u;
    tmem_ld_wait16(va);
    tmem_ld16(t_acc + 16 * (s + 1), vb);
    pipe_dgrad16<kSig>(va, mk[s >> 1], 0, wsp + 16 * s, dsr, pbase + (c0 ^ xr), pbase + ((c0 + 16u) ^ xr));
    tmem_ld_wait16(vb);
    if (s + 2 < 8) tmem_ld16(t_acc + 16 * (s + 2), va);
    pipe_dgrad16<kSig>(vb, mk[s >> 1], 8, wsp + 16 * (s + 1), dsr, pbase + ((c0 + 32u) ^ xr), pbase + ((c0 + 48u) ^ xr));
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(pipe::kThreads, 1) mlp_bwd_pipe_kernel(const PipeParams p) {
  using namespace pipe;
  using L = ParamLayout;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0) __trap();
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = (int)blockIdx.x >> 1;
  if (cluster_id >= kPipelines * kRoles) return;          // the two spare pairs
  const int pl = cluster_id / kRoles, role = cluster_id % kRoles;   // role 0 = head, 1..8 = stage st
  const int n_groups = (p.n_tiles + 1) / 2;
  const int n_mine = n_groups > pl ? (n_groups - pl + kPipelines - 1) / kPipelines : 0;   // groups pl, pl + 8, ...
  if (n_mine == 0) return;                                 // (identical in both CTAs and in every role of the pipeline)

  const uint32_t bars = smem_base + kOffBars;
  const uint32_t bar_full = bars;                          // [kRing]
  const uint32_t bar_empty = bars + 8 * kRing;             // [kRing] count 2: MMA commit + reducers
  const uint32_t bar_peer = bars + 16 * kRing;             // [kRing] leader only: the peer's stage has landed
  const uint32_t bar_w_full = bars + 24 * kRing;           // resident weights landed
  const uint32_t bar_w_peer = bar_w_full + 8;              // leader only
  const uint32_t bar_acc_ready = bar_w_peer + 8;           // [2] accumulator complete (multicast commit); stages use [0]
  const uint32_t bar_acc_free = bar_acc_ready + 16;        // leader only, count 16: accumulator drained by both CTAs
  const uint32_t bar_a_ready = bar_acc_free + 8;           // [2] head, leader only, count 16: dG operands written
  const uint32_t bar_img_full = bar_a_ready + 16;          // count 8: dY_out image complete in shared memory
  const uint32_t bar_img_empty = bar_img_full + 8;         // count 2: copied out by the two store warps
  const uint32_t bar_done = bar_img_empty + 8;             // every MMA of this pair has completed (multicast commit)
  const uint32_t tmem_slot = bar_done + 8;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kRing; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 2);
      mbar_init(bar_peer + 8 * i, 1);
    }
    mbar_init(bar_w_full, 1);
    mbar_init(bar_w_peer, 1);
    mbar_init(bar_acc_ready, 1);
    mbar_init(bar_acc_ready + 8, 1);
    mbar_init(bar_acc_free, 16);
    mbar_init(bar_a_ready, 16);
    mbar_init(bar_a_ready + 8, 16);
    mbar_init(bar_img_full, 8);
    mbar_init(bar_img_empty, 2);
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc2(tmem_slot, 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const uint64_t n_tiles64 = (uint64_t)p.n_tiles;
  const uint8_t* wimg = p.packed + kBwdImageOffset;
  auto link_base = [&](int k, int slot) -> uint8_t* { return p.links + ((size_t)(pl * kLinks + k) * kLinkSlots + slot) * kGroupBytes; };
  auto flag_ready = [&](int k, int slot) -> uint32_t* { return p.flags + ((size_t)(pl * kLinks + k) * kLinkSlots + slot) * kFlagStride; };
  auto flag_freed = [&](int k, int slot) -> uint32_t* {
    return p.flags + ((size_t)kPipelines * kLinks * kLinkSlots + (size_t)(pl * kLinks + k) * kLinkSlots + slot) * kFlagStride;
  };

  if (role == 0) {
    // ==================================================================================================================
    // HEAD: prologue on CUDA cores + the C0 dgrad stage, two ping-pong slots (groups t = slot, slot + 2, ...)
    // ==================================================================================================================
    if (warp < 4) {
      setmaxnreg_dec<kRegsOther>();
      if (warp == 0) {
        // resident weights: rows [128 rank, +128) of the two W_c0^T panels (dgrad stage 0 of mlp_layout.cuh)
        if (elect_one()) {
          mbar_arrive_expect_tx(bar_w_full, 2 * kStageBytes);
          for (int pp = 0; pp < 2; ++pp)
            bulk_g2s_hint(smem_base + kOffW + pp * kStageBytes, wimg + (uint32_t)(bwd_first_panel(0) + pp) * kPanelBytes256 + rank * kStageBytes,
                          kStageBytes, bar_w_full, l2_evict_last());
        }
        __syncwarp();
        if (rank == 1) {   // relay: my weights have landed
          mbar_wait(bar_w_full, 0);
          if (elect_one()) mbar_arrive_cluster(mapa(bar_w_peer, 0));
          __syncwarp();
        }
      } else if (warp == 1 && rank == 0) {
        constexpr uint32_t idesc = make_idesc(256, 256, kF16, kF16, 0, 0);
        mbar_wait(bar_w_full, 0);
        mbar_wait_cluster(bar_w_peer, 0);
        uint32_t a_phase[2] = {0, 0};
        for (int t = 0; t < n_mine; ++t) {
          const int slot = t & 1;
          mbar_wait_cluster(bar_a_ready + 8 * slot, a_phase[slot]);
          a_phase[slot] ^= 1;
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_smem = smem_base + kHeadOffA + slot * (2 * kPanelBytes128);
#pragma unroll
            for (int pp = 0; pp < 2; ++pp) {
              const uint64_t da = make_smem_desc(a_smem + pp * kPanelBytes128, 16u, kAtomBytes);
              const uint64_t db = make_smem_desc(smem_base + kOffW + pp * kStageBytes, 16u, kAtomBytes);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) umma2(tmem_base + slot * 256, da + 2u * ks, db + 2u * ks, idesc, (pp | ks) != 0);
            }
            umma_commit2(bar_acc_ready + 8 * slot, 3);
          }
          __syncwarp();
        }
      }
    } else {
      setmaxnreg_inc<kRegsEpilogue>();
      const int ew = warp - 4;
      const int slot = ew >> 3;
      const int half = (ew >> 2) & 1;
      const int wq = warp & 3;
      const int row = wq * 32 + lane;
      const int tg = threadIdx.x - 128 - slot * 256;
      const uint32_t a_smem = smem_base + kHeadOffA + slot * (2 * kPanelBytes128);
      const uint32_t out_smem = smem_base + kHeadOffOut + slot * kActBytes;
      const uint32_t t_acc = tmem_base + slot * 256 + half * 128 + (static_cast<uint32_t>(wq * 32) << 16);
      const uint32_t bar_id = 1 + slot;
      const uint32_t row_off = (uint32_t)(row >> 3) * kAtomBytes + (uint32_t)(row & 7) * kPanelRowBytes;
      const uint32_t xr = (uint32_t)(row & 7) << 4;
      const uint32_t out_h = out_smem + 2 * half * kPanelBytes128 + row_off;
      const uint32_t a_ready_leader = mapa(bar_a_ready + 8 * slot, 0);
      uint32_t acc_phase = 0;
      for (int t = slot; t < n_mine; t += 2) {
        const int group = pl + kPipelines * t;
        const int tile = group * 2 + (int)rank;
        const bool tile_ok = tile < p.n_tiles;
        const int64_t e = (int64_t)tile * kTile + row;
        const bool valid = tile_ok && e < p.n_evals;
        // ---- prologue (mlp_bwd.cu): dL/dg for this half's 64 colour-layer neurons, masked by the forward's g > 0 bits ----
        uint2 gm = make_uint2(0u, 0u);
        if (tile_ok)
          gm = __ldg(reinterpret_cast<const uint2*>(p.stash + stash_region_offset(kStashMask, n_tiles64) +
                                                    (uint64_t)tile * stash_region_tile_bytes(kStashMask) + 8 * (128 * 32) + row * 32 + half * 8));
        float dp0 = 0.f, dp1 = 0.f, dp2 = 0.f, dsr = 0.f;
        if (valid) {
          const float4 dd = __ldg(p.d_rgbsigma + e);
          const float4 o = __ldg(p.rgbsigma + e);
          dp0 = dd.x * o.x * (1.f - o.x);
          dp1 = dd.y * o.y * (1.f - o.y);
          dp2 = dd.z * o.z * (1.f - o.z);
          dsr = dd.w;
        }
        uint32_t outw[2][16];
#pragma unroll
        for (int ci = 0; ci < 2; ++ci) {
          const int c0 = 64 * half + 32 * ci;
          const uint32_t gbits = ci == 0 ? gm.x : gm.y;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.params + L::kWC1 + c0) + q);
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.params + L::kWC1 + 128 + c0) + q);
            const float4 w2 = __ldg(reinterpret_cast<const float4*>(p.params + L::kWC1 + 256 + c0) + q);
            const float d0 = dp0 * w0.x + dp1 * w1.x + dp2 * w2.x;
            const float d1 = dp0 * w0.y + dp1 * w1.y + dp2 * w2.y;
            const float d2 = dp0 * w0.z + dp1 * w1.z + dp2 * w2.z;
            const float d3 = dp0 * w0.w + dp1 * w1.w + dp2 * w2.w;
            const bool m0 = (gbits >> (2 * q)) & 1u, m1 = (gbits >> (16 + 2 * q)) & 1u;
            const bool m2 = (gbits >> (2 * q + 1)) & 1u, m3 = (gbits >> (16 + 2 * q + 1)) & 1u;
            outw[ci][2 * q] = pack_half2(m0 ? d0 : 0.f, m1 ? d1 : 0.f);
            outw[ci][2 * q + 1] = pack_half2(m2 ? d2 : 0.f, m3 ? d3 : 0.f);
          }
        }
        // the previous group's image stores of this slot (dG from a_smem, dF from out_smem) must have left shared memory
        if (tg == 0) bulk_wait_read<0>();
        named_bar_sync(bar_id, 256);
        {
          const uint32_t dpanel = a_smem + half * kPanelBytes128;
#pragma unroll
          for (int ci = 0; ci < 2; ++ci)
#pragma unroll
            for (int q = 0; q < 4; ++q)
              st_shared_v4(dpanel + panel_chunk_offset(row, 4 * ci + q), outw[ci][4 * q], outw[ci][4 * q + 1], outw[ci][4 * q + 2],
                           outw[ci][4 * q + 3]);
        }
        fence_proxy_async_smem();
        named_bar_sync(bar_id, 256);
        if (tg == 0 && tile_ok) {   // dG image for the residual weight-gradient kernel (colour layer 0)
          bulk_s2g_hint(p.gstash + grad_region_offset(kGradC0, n_tiles64) + (uint64_t)tile * grad_region_tile_bytes(kGradC0), a_smem,
                        2 * kPanelBytes128, l2_evict_first());
          bulk_commit();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(a_ready_leader);
        if (tile_ok) {   // head-gradient panel: cols 0..2 = dL/d(rgb pre-sigmoid), col 3 = dL/dsigma_raw, rest zero
          uint8_t* hd = p.gstash + grad_region_offset(kGradHead, n_tiles64) + (uint64_t)tile * grad_region_tile_bytes(kGradHead);
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            uint4 v = make_uint4(0, 0, 0, 0);
            if (ch == 0 && half == 0) {
              v.x = pack_half2(dp0, dp1);
              v.y = pack_half2(dp2, dsr);
            }
            *reinterpret_cast<uint4*>(hd + panel_chunk_offset(row, 4 * half + ch)) = v;
          }
        }
        // ---- dF = dG W_c0[:, :256] (no mask: f is linear) -> fp16 image -> link 0 ----
        mbar_wait(bar_acc_ready + 8 * slot, acc_phase);
        acc_phase ^= 1;
        tc_fence_after();
        {
          const uint32_t mk[4] = {~0u, ~0u, ~0u, ~0u};
          pipe_drain_acc<false>(t_acc, mk, nullptr, 0.f, out_h, xr);
        }
        tc_fence_before();
        fence_proxy_async_smem();
        named_bar_sync(bar_id, 256);
        if (tg == 0) {
          const int ls = t % kLinkSlots, use = t / kLinkSlots;
          spin_until(flag_freed(0, ls), 2u * (uint32_t)use);          // both consumer CTAs have released the slot's previous group
          bulk_s2g(link_base(0, ls) + rank * kActBytes, out_smem, kActBytes);
          bulk_commit();
          bulk_wait_all<0>();                                           // writes complete (not only read) before the flag
          fence_proxy_async_all();
          __threadfence();
          red_release_gpu(flag_ready(0, ls), 2u);                       // (stage producers add 1 per store warp: 4 per group either way)
        }
      }
      if (tg == 0) bulk_wait_all<0>();
    }
  } else {
    // ==================================================================================================================
    // STAGE st = role: dY_in -> dY_out (dgrad) and dW += dY_in^T H (wgrad), layer-stationary
    // ==================================================================================================================
    const int st = role;                         // 1 = F, 2..8 = hidden layers 7..1 (chain stage numbering of mlp_bwd.cu)
    const int k_in = st - 1;                     // input link
    const int h_region = kStashH0 + (8 - st);    // activations that fed this layer: h7, h6, ..., h0
    if (warp < 4) {
      setmaxnreg_dec<kRegsOther>();
      if (warp == 0) {
        // ------------------------------------------------ loader ------------------------------------------------
        if (elect_one()) {
          mbar_arrive_expect_tx(bar_w_full, 4 * kStageBytes);
          for (int pp = 0; pp < 4; ++pp)
            bulk_g2s_hint(smem_base + kOffW + pp * kStageBytes, wimg + (uint32_t)(bwd_first_panel(st) + pp) * kPanelBytes256 + rank * kStageBytes,
                          kStageBytes, bar_w_full, l2_evict_last());
        }
        __syncwarp();
        const uint64_t pol_stream = l2_evict_first();
        uint32_t stage = 0, phase = 0;
        auto next_stage = [&]() {
          if (++stage == (uint32_t)kRing) {
            stage = 0;
            phase ^= 1;
          }
        };
        for (int t = 0; t < n_mine; ++t) {
          const int group = pl + kPipelines * t;
          const int ls = t % kLinkSlots, use = t / kLinkSlots;
          if (elect_one()) {
            spin_until(flag_ready(k_in, ls), 4u * (uint32_t)(use + 1));
            fence_proxy_async_all();                 // the bulk loads below (async proxy) must observe the producer's data
          }
          __syncwarp();
          const uint8_t* lbase = link_base(k_in, ls);
          // dgrad operand: this CTA's tile, K panels 0..3
          for (int pp = 0; pp < 4; ++pp) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(bar_full + 8 * stage, kStageBytes);
              bulk_g2s(smem_base + kOffRing + stage * kStageBytes, lbase + rank * kActBytes + pp * kPanelBytes128, kStageBytes, bar_full + 8 * stage);
            }
            __syncwarp();
            next_stage();
          }
          // wgrad operands: K = the 64-sample halves of tile q; this CTA's 128 feature columns = panels 2 rank, 2 rank + 1
          for (int q = 0; q < 2; ++q) {
            int tile_q = group * 2 + q;
            if (tile_q >= p.n_tiles) tile_q = p.n_tiles - 1;   // (its dY is all zero: any finite activations will do)
            const uint8_t* hbase = p.stash + stash_region_offset(h_region, n_tiles64) + (uint64_t)tile_q * kActBytes;
            for (int h = 0; h < 2; ++h) {
              for (int xy = 0; xy < 2; ++xy) {
                mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                if (elect_one()) {
                  mbar_arrive_expect_tx(bar_full + 8 * stage, kStageBytes);
                  const uint32_t dst = smem_base + kOffRing + stage * kStageBytes;
                  for (int i = 0; i < 2; ++i) {
                    const uint32_t off = (uint32_t)(2 * rank + i) * kPanelBytes128 + (uint32_t)h * kHalfPanel;
                    if (xy == 0) bulk_g2s(dst + i * kHalfPanel, lbase + q * kActBytes + off, kHalfPanel, bar_full + 8 * stage);
                    else bulk_g2s_hint(dst + i * kHalfPanel, hbase + off, kHalfPanel, bar_full + 8 * stage, pol_stream);
                  }
                }
                __syncwarp();
                next_stage();
              }
            }
          }
        }
      } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer (leader) / relay (peer) ------------------------------------------------
        const bool leader = rank == 0;
        constexpr uint32_t idesc_d = make_idesc(256, 256, kF16, kF16, 0, 0);
        constexpr uint32_t idesc_w = make_idesc(256, 256, kF16, kF16, 1, 1);
        const uint32_t acc = tmem_base, dw = tmem_base + 256;
        mbar_wait(bar_w_full, 0);
        if (leader) mbar_wait_cluster(bar_w_peer, 0);
        else {
          if (elect_one()) mbar_arrive_cluster(mapa(bar_w_peer, 0));
          __syncwarp();
        }
        uint32_t stage = 0, phase = 0, free_phase = 0;
        auto next_stage = [&]() {
          if (++stage == (uint32_t)kRing) {
            stage = 0;
            phase ^= 1;
          }
        };
        // the stage's operands have landed in BOTH CTAs (leader) / in this CTA, announced to the leader (peer)
        auto stage_ready = [&]() {
          mbar_wait(bar_full + 8 * stage, phase);
          if (leader) mbar_wait_cluster(bar_peer + 8 * stage, phase);
          else {
            if (elect_one()) mbar_arrive_cluster(mapa(bar_peer + 8 * stage, 0));
            __syncwarp();
          }
        };
        for (int t = 0; t < n_mine; ++t) {
          const int ls = t % kLinkSlots;
          if (leader && t > 0) {   // the accumulator of the previous group has been drained by both CTAs
            mbar_wait_cluster(bar_acc_free, free_phase);
            free_phase ^= 1;
          }
          for (int pp = 0; pp < 4; ++pp) {
            stage_ready();
            if (leader) {
              tc_fence_after();
              if (elect_one()) {
                const uint64_t da = make_smem_desc(smem_base + kOffRing + stage * kStageBytes, 16u, kAtomBytes);
                const uint64_t db = make_smem_desc(smem_base + kOffW + pp * kStageBytes, 16u, kAtomBytes);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma2(acc, da + 2u * ks, db + 2u * ks, idesc_d, (pp | ks) != 0);
                umma_commit2(bar_empty + 8 * stage, 3);
                if (pp == 3) umma_commit2(bar_acc_ready, 3);
              }
              __syncwarp();
            }
            next_stage();
          }
          for (int c = 0; c < 4; ++c) {
            stage_ready();
            const uint32_t sx = stage;
            next_stage();
            stage_ready();
            const uint32_t sy = stage;
            next_stage();
            if (c == 3) {   // every load from the input link slot has landed in this CTA: give the slot back to the producer
              if (elect_one()) red_release_gpu(flag_freed(k_in, ls), 1u);
              __syncwarp();
            }
            if (leader) {
              tc_fence_after();
              if (elect_one()) {
                const uint32_t ax = smem_base + kOffRing + sx * kStageBytes, by = smem_base + kOffRing + sy * kStageBytes;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  umma2(dw, desc_mnmajor(ax, ks, kHalfPanel), desc_mnmajor(by, ks, kHalfPanel), idesc_w, (t | c | ks) != 0);
                umma_commit2(bar_empty + 8 * sx, 3);
                umma_commit2(bar_empty + 8 * sy, 3);
              }
              __syncwarp();
            }
          }
        }
        if (leader) {
          if (elect_one()) umma_commit2(bar_done, 3);
          __syncwarp();
        }
      }
    } else if (warp < 12) {
      // ------------------------------------------------ epilogue: acc -> mask -> fp16 image ------------------------------------------------
      setmaxnreg_inc<kRegsEpilogue>();
      const int ew = warp - 4;
      const int half = (ew >> 2) & 1;
      const int wq = warp & 3;
      const int row = wq * 32 + lane;
      const uint32_t t_acc = tmem_base + half * 128 + (static_cast<uint32_t>(wq * 32) << 16);
      const uint32_t row_off = (uint32_t)(row >> 3) * kAtomBytes + (uint32_t)(row & 7) * kPanelRowBytes;
      const uint32_t xr = (uint32_t)(row & 7) << 4;
      const uint32_t out_h = smem_base + kOffOut + 2 * half * kPanelBytes128 + row_off;
      const uint32_t acc_free_leader = mapa(bar_acc_free, 0);
      const int mask_layer = 8 - st;
      const float* wsp = p.params + L::kWS + 128 * half;
      uint32_t acc_phase = 0;
      for (int t = 0; t < n_mine; ++t) {
        const int group = pl + kPipelines * t;
        const int tile = group * 2 + (int)rank;
        const bool tile_ok = tile < p.n_tiles;
        const int64_t e = (int64_t)tile * kTile + row;
        uint4 mk4 = make_uint4(~0u, ~0u, ~0u, ~0u);
        if (tile_ok)
          mk4 = __ldg(reinterpret_cast<const uint4*>(p.stash + stash_region_offset(kStashMask, n_tiles64) +
                                                     (uint64_t)tile * stash_region_tile_bytes(kStashMask) + mask_layer * (128 * 32) + row * 32 + half * 16));
        float dsr = 0.f;
        if (st == 1 && tile_ok && e < p.n_evals) dsr = __ldg(p.d_rgbsigma + e).w;
        mbar_wait(bar_acc_ready, acc_phase);
        acc_phase ^= 1;
        tc_fence_after();
        if (t > 0) mbar_wait(bar_img_empty, (uint32_t)(t - 1) & 1u);   // the previous image has been copied out of kOffOut
        const uint32_t mk[4] = {mk4.x, mk4.y, mk4.z, mk4.w};
        if (st == 1) pipe_drain_acc<true>(t_acc, mk, wsp, dsr, out_h, xr);
        else pipe_drain_acc<false>(t_acc, mk, wsp, dsr, out_h, xr);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_cluster(acc_free_leader);   // accumulator drained (tcgen05.wait::ld of every chunk has retired)
          mbar_arrive(bar_img_full);
        }
      }
      // ---- flush this CTA's half of dW (rows 128 rank .. +128 = output neurons, 256 columns = input features) ----
      mbar_wait(bar_done, 0);
      tc_fence_after();
      {
        const int layer = 8 - st + 1;   // st = 1 -> feature layer, st = 2..8 -> hidden layers 7..1
        const int64_t w_off = st == 1 ? L::kWF : L::hidden_w(layer);
        const int ld = st == 1 ? 256 : L::hidden_in(layer);
        float* wrow = p.grads + w_off + (int64_t)(128 * rank + row) * ld + 128 * half;
        const uint32_t t_dw = tmem_base + 256 + half * 128 + (static_cast<uint32_t>(wq * 32) << 16);
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(t_dw + c0, v);
          tmem_ld_wait32(v);
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(wrow + c0 + j, __uint_as_float(v[j]) * p.inv_scale);
        }
      }
      tc_fence_before();
    } else if (warp < 16) {
      // ------------------------------------------------ reducers: bias (and density-head) gradients ------------------------------------------------
      // Every ring stage is visited in the loader's order; the X stages (dY_in: 64 samples x this CTA's 128 columns) feed the
      // column sums, and in stage F the Y stages (h7) feed dL/dw_sigma.  A consumer arrives on `empty` only after the stage's
      // `full` phase: the barrier counts two arrivals per use (MMA commit + this group) and must never see two of one kind.
      const int tr = threadIdx.x - 12 * 32;      // 0..127
      const int cp = tr & 63;                    // column pair 2 cp, 2 cp + 1 of the 128
      const int rh = tr >> 6;                    // rows [32 rh, 32 rh + 32) of the 64
      float s0 = 0.f, s1 = 0.f, d0 = 0.f, d1 = 0.f, dsum = 0.f;
      float* dsig = reinterpret_cast<float*>(smem_raw + kOffDsig);
      uint32_t stage = 0, phase = 0;
      auto next_stage = [&]() {
        if (++stage == (uint32_t)kRing) {
          stage = 0;
          phase ^= 1;
        }
      };
      auto col_base = [&](uint32_t s) { return smem_base + kOffRing + s * kStageBytes + (uint32_t)(cp >> 5) * kHalfPanel; };
      for (int t = 0; t < n_mine; ++t) {
        const int group = pl + kPipelines * t;
        for (int pp = 0; pp < 4; ++pp) {
          mbar_wait(bar_full + 8 * stage, phase);
          named_bar_sync(3, 128);
          if (tr == 0) mbar_arrive(bar_empty + 8 * stage);
          next_stage();
        }
        for (int c = 0; c < 4; ++c) {
          // X: dY_in
          mbar_wait(bar_full + 8 * stage, phase);
          {
            const uint32_t base = col_base(stage);
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
              uint32_t w;
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(base + panel_offset(32 * rh + r, (2 * cp) & 63)));
              const float2 f = __half22float2(*reinterpret_cast<__half2*>(&w));
              s0 += f.x;
              s1 += f.y;
            }
          }
          if (st == 1) {   // dsigma_raw of the 64 samples of this half tile (zero beyond the batch)
            if (tr < 64) {
              const int64_t em = (int64_t)(group * 2 + (c >> 1)) * kTile + 64 * (c & 1) + tr;
              dsig[tr] = em < p.n_evals ? __ldg(p.d_rgbsigma + em).w : 0.f;
            }
          }
          named_bar_sync(3, 128);
          if (tr == 0) mbar_arrive(bar_empty + 8 * stage);
          next_stage();
          // Y: activations
          mbar_wait(bar_full + 8 * stage, phase);
          if (st == 1) {
            const uint32_t base = col_base(stage);
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
              uint32_t w;
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(base + panel_offset(32 * rh + r, (2 * cp) & 63)));
              const float2 f = __half22float2(*reinterpret_cast<__half2*>(&w));
              const float ds = dsig[32 * rh + r];
              d0 = fmaf(ds, f.x, d0);
              d1 = fmaf(ds, f.y, d1);
              if (cp == 0) dsum += ds;
            }
          }
          named_bar_sync(3, 128);
          if (tr == 0) mbar_arrive(bar_empty + 8 * stage);
          next_stage();
        }
      }
      {
        const int layer = 8 - st + 1;
        const int64_t b_off = st == 1 ? L::kBF : L::hidden_b(layer);
        atomicAdd(p.grads + b_off + 128 * rank + 2 * cp, s0 * p.inv_scale);
        atomicAdd(p.grads + b_off + 128 * rank + 2 * cp + 1, s1 * p.inv_scale);
        if (st == 1) {
          atomicAdd(p.grads + L::kWS + 128 * rank + 2 * cp, d0 * p.inv_scale);
          atomicAdd(p.grads + L::kWS + 128 * rank + 2 * cp + 1, d1 * p.inv_scale);
          if (cp == 0 && rank == 0) atomicAdd(p.grads + L::kBS, dsum * p.inv_scale);   // (both CTAs see all 256 samples)
        }
      }
    } else if (warp < 18) {
      // ------------------------------------------------ store warps: dY_out image -> next link (and HBM where the residual kernel needs it) -------
      const int sw = warp - 16;                  // each copies one half (two panels) of the image
      const uint8_t* src = smem_raw + kOffOut + sw * (kActBytes / 2);
      const int gregion = st == 3 ? kGradL7 + 2 : (st == 8 ? kGradL0 : -1);   // dY5 (layer-5 encoding columns) / dY0 (layer 0)
      for (int t = 0; t < n_mine; ++t) {
        const int group = pl + kPipelines * t;
        const int tile = group * 2 + (int)rank;
        const bool tile_ok = tile < p.n_tiles;
        const int ls = t % kLinkSlots, use = t / kLinkSlots;
        mbar_wait(bar_img_full, (uint32_t)t & 1u);
        const uint4* sp = reinterpret_cast<const uint4*>(src);
        if (st < 8) {
          if (lane == 0) spin_until(flag_freed(st, ls), 2u * (uint32_t)use);
          __syncwarp();
          uint4* dst = reinterpret_cast<uint4*>(link_base(st, ls) + rank * kActBytes + sw * (kActBytes / 2));
#pragma unroll 4
          for (int q = lane; q < (int)(kActBytes / 32); q += 32) dst[q] = sp[q];
        }
        if (gregion >= 0 && tile_ok) {
          uint4* dst = reinterpret_cast<uint4*>(p.gstash + grad_region_offset(gregion, n_tiles64) + (uint64_t)tile * kActBytes + sw * (kActBytes / 2));
#pragma unroll 4
          for (int q = lane; q < (int)(kActBytes / 32); q += 32) __stcs(dst + q, sp[q]);
        }
        __threadfence();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar_img_empty);
          if (st < 8) red_release_gpu(flag_ready(st, ls), 1u);
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc2(tmem_base, 512);
}

int launch_wgrad_residual(float* grads, const uint8_t* stash, const uint8_t* gstash, int n_tiles, float inv_scale, cudaStream_t stream);

}  // namespace nerf

extern "C" size_t nerf_mlp_backward_pipe_workspace_bytes(void) { return nerf::pipe::kLinkBytes + nerf::pipe::kFlagBytes; }

// Fused backward: `workspace` = gradient stash (nerf_mlp_backward_workspace_bytes) followed by the pipeline's links and flags.
extern "C" int nerf_mlp_backward_pipe(float* grads, const float* d_rgbsigma, const float* rgbsigma, const void* stash, void* workspace,
                                      const void* packed, const float* params, int n_rays, int n_samples, float grad_scale, void* stream) {
  using namespace nerf;
  if (n_rays <= 0) return 0;
  NERF_CHECK_ARG(grads && d_rgbsigma && rgbsigma && stash && workspace && packed && params, "mlp_backward_pipe: null pointer");
  NERF_CHECK_ARG(grad_scale > 0.f, "mlp_backward_pipe: grad_scale must be positive");
  const int64_t n_evals = (int64_t)n_rays * n_samples;
  NERF_CHECK_ARG(n_evals < (int64_t(1) << 31) - kTile, "mlp_backward_pipe: n_rays*n_samples must be < 2^31 per call");
  NERF_CHECK_ARG(((reinterpret_cast<uintptr_t>(stash) | reinterpret_cast<uintptr_t>(workspace) | reinterpret_cast<uintptr_t>(packed)) & 127) == 0,
                 "mlp_backward_pipe: stash, workspace and packed must be 128-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  PipeParams p;
  p.grads = grads;
  p.d_rgbsigma = reinterpret_cast<const float4*>(d_rgbsigma);
  p.rgbsigma = reinterpret_cast<const float4*>(rgbsigma);
  p.stash = static_cast<const uint8_t*>(stash);
  p.gstash = static_cast<uint8_t*>(workspace);
  p.n_evals = n_evals;
  p.n_tiles = (int)((n_evals + kTile - 1) / kTile);
  const size_t gbytes = (size_t)(grad_tile_bytes_total() * (uint64_t)p.n_tiles);
  p.links = p.gstash + ((gbytes + 1023) & ~size_t(1023));
  p.flags = reinterpret_cast<uint32_t*>(p.links + pipe::kLinkBytes);
  p.packed = static_cast<const uint8_t*>(packed);
  p.params = params;
  p.inv_scale = 1.f / grad_scale;
  // a pipeline whose consumer waits for a producer needs every pair resident at once: 74 clusters on 148 SMs
  static int clusters_ok[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!clusters_ok[dev & 63]) {
    cudaError_t e1 = cudaFuncSetAttribute(mlp_bwd_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pipe::kSmemBytes);
    NERF_CHECK_ARG(e1 == cudaSuccess, "mlp_backward_pipe: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kNumSMs);
    cfg.blockDim = dim3(pipe::kThreads);
    cfg.dynamicSmemBytes = pipe::kSmemBytes;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int max_clusters = 0;
    cudaError_t e2 = cudaOccupancyMaxActiveClusters(&max_clusters, mlp_bwd_pipe_kernel, &cfg);
    NERF_CHECK_ARG(e2 == cudaSuccess && max_clusters >= pipe::kPipelines * pipe::kRoles,
                   "mlp_backward_pipe: the device cannot keep %d CTA pairs resident at once (%d): %s", pipe::kPipelines * pipe::kRoles, max_clusters,
                   cudaGetErrorString(e2));
    clusters_ok[dev & 63] = 1;
  }
  cudaError_t em = cudaMemsetAsync(p.flags, 0, pipe::kFlagBytes, s);
  NERF_CHECK_ARG(em == cudaSuccess, "mlp_backward_pipe: cudaMemsetAsync failed: %s", cudaGetErrorString(em));
  mlp_bwd_pipe_kernel<<<kNumSMs, pipe::kThreads, pipe::kSmemBytes, s>>>(p);
  NERF_CHECK_LAUNCH("mlp_bwd_pipe_kernel");
  return launch_wgrad_residual(grads, p.stash, p.gstash, p.n_tiles, p.inv_scale, s);
}
