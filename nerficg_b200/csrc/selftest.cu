// selftest.cu -- one-tile exerciser of the tcgen05 building blocks in tc.cuh.
// D[128][n] = A[128][k] * B[n][k]^T, operands converted on device into the panel format.
// mode 0: fp16 x fp16, K-major panels (forward / dgrad flavour)
// mode 1: fp16 x fp16, MN-major view of reduction-major panels (wgrad flavour)
// mode 2: bf16 x bf16, K-major.  (Mixed bf16 x fp16 operands are an illegal instruction on sm_100a --
// measured in round 1 -- which is why the backward uses loss-scaled fp16 gradients.)
#include "common.cuh"
#include "tc.cuh"
#include "../../include/nerf_b200.h"

namespace nerf {
using namespace tc;

__device__ __forceinline__ uint16_t to_bits(float v, bool bf16) {
  if (bf16) {
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    return *reinterpret_cast<uint16_t*>(&h);
  }
  __half h = __float2half_rn(v);
  return *reinterpret_cast<uint16_t*>(&h);
}

__global__ void __launch_bounds__(128, 1) selftest_kernel(float* __restrict__ out, const float* __restrict__ a,
                                                          const float* __restrict__ b, int n, int k, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;

  const bool mn_major = (mode == 1);
  const bool a_bf16 = (mode == 2);
  const uint32_t a_bytes = 128u * k * 2u;
  uint8_t* sa = smem;
  uint8_t* sb = smem + a_bytes;
  const int tid = threadIdx.x;

  if (!mn_major) {
    // K-major: panel p holds columns [64p, 64p+64) of every row; A panels have 128 rows, B panels n rows
    for (int i = tid; i < 128 * k; i += 128) {
      int r = i / k, c = i % k;
      *reinterpret_cast<uint16_t*>(sa + (c / 64) * (128 * 128) + panel_offset(r, c % 64)) = to_bits(a[i], a_bf16);
    }
    for (int i = tid; i < n * k; i += 128) {
      int r = i / k, c = i % k;
      *reinterpret_cast<uint16_t*>(sb + (c / 64) * (n * 128) + panel_offset(r, c % 64)) = to_bits(b[i], a_bf16);
    }
  } else {
    // reduction-major: panel g holds M/N indices [64g, 64g+64) as columns, rows = k index
    for (int i = tid; i < 128 * k; i += 128) {
      int m = i / k, kk = i % k;
      *reinterpret_cast<uint16_t*>(sa + (m / 64) * (k * 128) + panel_offset(kk, m % 64)) = to_bits(a[i], a_bf16);
    }
    for (int i = tid; i < n * k; i += 128) {
      int nn = i / k, kk = i % k;
      *reinterpret_cast<uint16_t*>(sb + (nn / 64) * (k * 128) + panel_offset(kk, nn % 64)) = to_bits(b[i], a_bf16);
    }
  }
  uint32_t ncols = 32;
  while (ncols < static_cast<uint32_t>(n)) ncols <<= 1;
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_barrier_init();
  }
  if (tid < 32) {
    tmem_alloc(smem_u32(&tmem_base_s), ncols);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, n, a_bf16 ? kBF16 : kF16, a_bf16 ? kBF16 : kF16, mn_major ? 1 : 0, mn_major ? 1 : 0);
    uint32_t acc = 0;
    for (int ks = 0; ks < k / 16; ++ks) {
      uint64_t da, db;
      if (!mn_major) {
        da = desc_kmajor(smem_u32(sa) + (ks / 4) * (128 * 128), ks % 4);
        db = desc_kmajor(smem_u32(sb) + (ks / 4) * (n * 128), ks % 4);
      } else {
        da = desc_mnmajor(smem_u32(sa), ks, k * 128);
        db = desc_mnmajor(smem_u32(sb), ks, k * 128);
      }
      umma(tmem, da, db, idesc, acc);
      acc = 1;
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  const int warp = tid / 32, lane = tid % 32;
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < n; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j)
      if (c0 + j < n) out[row * n + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, ncols);
}

// ---- CTA-pair flavour: D[256][n] = A[256][k] * B[n][k]^T with one cta_group::2 instruction stream ----------
// CTA r of the 2-CTA cluster holds A rows [128r, 128r+128) and B rows [n/2 * r, n/2 * (r+1)); the leader issues
// M = 256 MMAs, the peer announces its operands with a remote mbarrier arrive (the relay used by the MLP kernels)
// and both CTAs are released by ONE multicast tcgen05.commit.  Checks the M / N split conventions of tc.cuh.
// kMn: both operands are stored reduction-major (rows = k index, 64-wide column groups = M / N index, group stride k * 128 bytes)
// and read as MN-major operands -- the cta_group::2 flavour of the weight-gradient MMAs (dW += dY^T X over the samples).
template <bool kMn>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
    selftest2_kernel(float* __restrict__ out, const float* __restrict__ a, const float* __restrict__ b, int n, int k) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bar_done, bar_peer;
  __shared__ uint32_t tmem_base_s;
  const uint32_t rank = cluster_ctarank();
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  const int nh = n / 2;
  uint8_t* sa = smem;
  uint8_t* sb = smem + 128u * k * 2u;
  if (tid == 0) {
    mbar_init(smem_u32(&bar_done), 1);
    mbar_init(smem_u32(&bar_peer), 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc2(smem_u32(&tmem_base_s), 256);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  for (int i = tid; i < 128 * k; i += 128) {
    int r = i / k, c = i % k;
    uint8_t* dst = kMn ? sa + (r / 64) * (k * 128) + panel_offset(c, r % 64) : sa + (c / 64) * (128 * 128) + panel_offset(r, c % 64);
    *reinterpret_cast<uint16_t*>(dst) = to_bits(a[(size_t)(128 * rank + r) * k + c], false);
  }
  for (int i = tid; i < nh * k; i += 128) {
    int r = i / k, c = i % k;
    uint8_t* dst = kMn ? sb + (r / 64) * (k * 128) + panel_offset(c, r % 64) : sb + (c / 64) * (nh * 128) + panel_offset(r, c % 64);
    *reinterpret_cast<uint16_t*>(dst) = to_bits(b[(size_t)(nh * rank + r) * k + c], false);
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (rank == 1 && tid == 0) mbar_arrive_cluster(mapa(smem_u32(&bar_peer), 0));
  if (rank == 0 && warp == 0) {
    mbar_wait_cluster(smem_u32(&bar_peer), 0);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t idesc = make_idesc(256, n, kF16, kF16, kMn ? 1 : 0, kMn ? 1 : 0);
      for (int ks = 0; ks < k / 16; ++ks) {
        if (kMn)
          umma2(tmem, desc_mnmajor(smem_u32(sa), ks, k * 128), desc_mnmajor(smem_u32(sb), ks, k * 128), idesc, ks != 0);
        else
          umma2(tmem, desc_kmajor(smem_u32(sa) + (ks / 4) * (128 * 128), ks % 4), desc_kmajor(smem_u32(sb) + (ks / 4) * (nh * 128), ks % 4),
                idesc, ks != 0);
      }
      umma_commit2(smem_u32(&bar_done), 3);
    }
    __syncwarp();
  }
  mbar_wait_cluster(smem_u32(&bar_done), 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < n; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    tmem_ld_wait32(v);
    for (int j = 0; j < 32; ++j) out[(size_t)(128 * rank + row) * n + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc2(tmem, 256);
}

// ---- TMEM read bandwidth probe ---------------------------------------------------------------------
// n_warps warps (warp w: lane quarter w % 4, columns 32 * (w / 4)) each issue `iters` accumulator loads back to back:
// mode 0 = one 32x32b.x32 load per wait, 1 = two x32 loads in flight per wait, 2 = one x16 load per wait.
// out[0] = cycles between the two block barriers, out[1] = bytes read from TMEM.  Sizes the epilogue floor of the chain
// kernels: every stage must read its 128 x 256 fp32 accumulator (128 KB) out of TMEM.
__global__ void __launch_bounds__(512, 1) tmem_read_probe_kernel(unsigned long long* __restrict__ out, int mode, int iters) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(smem_u32(&tmem_base_s), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t = tmem_base_s + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64) % 448u;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (mode == 0) {
      uint32_t v[32];
      tmem_ld32(t, v);
      tmem_ld_wait32(v);
      acc += v[0] ^ v[31];
    } else if (mode == 1) {
      uint32_t v[32], w[32];
      tmem_ld32(t, v);
      tmem_ld32(t + 32, w);
      tmem_ld_wait32(v);
      asm volatile("" : "+r"(w[0]), "+r"(w[31]));
      acc += v[0] ^ w[31];
    } else {
      uint32_t v[16];
      tmem_ld16(t, v);
      tmem_ld_wait16(v);
      acc += v[0] ^ v[15];
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    out[0] = (unsigned long long)(t1 - t0);
    const unsigned long long per = mode == 0 ? 4096ull : (mode == 1 ? 8192ull : 2048ull);
    out[1] = per * (unsigned long long)iters * (blockDim.x >> 5);
  }
  if (acc == 0x12345678u) out[2] = acc;  // keep the loads alive
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}

}  // namespace nerf

// ---- tensor-pipe rate probe: `iters` back-to-back cta_group::2 MMAs (M = 256, N = 256, K = 16) on resident operands ------------
// mode 0: K-major A and B (the chains' flavour); 1: MN-major A and B (weight gradients).  out[0] = cycles from the first issue to the
// completion barrier.  Answers: do the MN-major MMAs of the fused backward run at the K-major rate (128 cycles each)?
namespace nerf {
template <bool kMn>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) umma2_rate_kernel(unsigned long long* __restrict__ out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ uint32_t tmem_base_s;
  const uint32_t rank = cluster_ctarank();
  const int tid = threadIdx.x, warp = tid / 32;
  for (int i = tid; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw + (smem_base - smem_u32(smem_raw)))[i] = 0x3c003c00u;  // fp16 1.0
  if (tid == 0) {
    mbar_init(smem_u32(&bar_done), 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc2(smem_u32(&tmem_base_s), 256);
    tmem_relinquish2();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  long long t0 = 0;
  if (rank == 0 && warp == 0) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc(256, 256, kF16, kF16, kMn ? 1 : 0, kMn ? 1 : 0);
      const uint32_t sa = smem_base, sb = smem_base + 32768;
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const int ks = i & 3;
        if (kMn) umma2(tmem, desc_mnmajor(sa, ks, 16384), desc_mnmajor(sb, ks, 16384), idesc, 1);
        else umma2(tmem, desc_kmajor(sa, ks), desc_kmajor(sb, ks), idesc, 1);
      }
      umma_commit2(smem_u32(&bar_done), 3);
    }
    __syncwarp();
  }
  mbar_wait_cluster(smem_u32(&bar_done), 0);
  if (rank == 0 && tid == 0) out[0] = (unsigned long long)(clock64() - t0);
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc2(tmem, 256);
}
}  // namespace nerf

extern "C" int nerf_selftest_umma2_rate(unsigned long long* out, int mn_major, int iters, void* stream) {
  using namespace nerf;
  NERF_CHECK_ARG(out != nullptr && iters > 0, "selftest_umma2_rate: bad arguments");
  const int smem = 65536 + 1024;
  cudaError_t e = mn_major ? cudaFuncSetAttribute(umma2_rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
                           : cudaFuncSetAttribute(umma2_rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  NERF_CHECK_ARG(e == cudaSuccess, "selftest_umma2_rate: %s", cudaGetErrorString(e));
  if (mn_major) umma2_rate_kernel<true><<<2, 128, smem, static_cast<cudaStream_t>(stream)>>>(out, iters);
  else umma2_rate_kernel<false><<<2, 128, smem, static_cast<cudaStream_t>(stream)>>>(out, iters);
  NERF_CHECK_LAUNCH("umma2_rate_kernel");
  return 0;
}

static int run_selftest2(float* d_out, const float* a, const float* b, int n, int k, bool mn, void* stream) {
  using namespace nerf;
  NERF_CHECK_ARG(n >= 64 && n <= 256 && n % (mn ? 128 : 64) == 0, "selftest2: n must be a multiple of %d in [64,256], got %d", mn ? 128 : 64, n);
  NERF_CHECK_ARG(k >= 64 && k <= 256 && k % 64 == 0, "selftest2: k must be a multiple of 64 in [64,256], got %d", k);
  size_t smem = 1024 + size_t(128) * k * 2 + size_t(n / 2) * k * 2;
  cudaError_t e = mn ? cudaFuncSetAttribute(selftest2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                     : cudaFuncSetAttribute(selftest2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  NERF_CHECK_ARG(e == cudaSuccess, "selftest2: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  if (mn)
    selftest2_kernel<true><<<2, 128, smem, static_cast<cudaStream_t>(stream)>>>(d_out, a, b, n, k);
  else
    selftest2_kernel<false><<<2, 128, smem, static_cast<cudaStream_t>(stream)>>>(d_out, a, b, n, k);
  NERF_CHECK_LAUNCH("selftest2_kernel");
  return 0;
}

extern "C" int nerf_selftest_umma2(float* d_out, const float* a, const float* b, int n, int k, void* stream) {
  return run_selftest2(d_out, a, b, n, k, false, stream);
}
extern "C" int nerf_selftest_umma2_mn(float* d_out, const float* a, const float* b, int n, int k, void* stream) {
  return run_selftest2(d_out, a, b, n, k, true, stream);
}

extern "C" int nerf_selftest_umma(float* d_out, const float* a, const float* b, int n, int k, int mode, void* stream) {
  using namespace nerf;
  NERF_CHECK_ARG(n >= 32 && n <= 256 && n % 32 == 0, "selftest: n must be a multiple of 32 in [32,256], got %d", n);
  NERF_CHECK_ARG(k >= 64 && k <= 256 && k % 64 == 0, "selftest: k must be a multiple of 64 in [64,256], got %d", k);
  NERF_CHECK_ARG(mode >= 0 && mode <= 2, "selftest: bad mode %d", mode);
  size_t smem = 1024 + size_t(128) * k * 2 + size_t(n) * k * 2;
  cudaError_t e = cudaFuncSetAttribute(selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  NERF_CHECK_ARG(e == cudaSuccess, "selftest: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  selftest_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(d_out, a, b, n, k, mode);
  NERF_CHECK_LAUNCH("selftest_kernel");
  return 0;
}

extern "C" int nerf_selftest_tmem_read(unsigned long long* out, int n_warps, int mode, int iters, void* stream) {
  using namespace nerf;
  NERF_CHECK_ARG(out != nullptr && n_warps >= 1 && n_warps <= 16 && mode >= 0 && mode <= 2 && iters >= 1,
                 "selftest_tmem_read: bad arguments");
  tmem_read_probe_kernel<<<1, 32 * n_warps, 0, static_cast<cudaStream_t>(stream)>>>(out, mode, iters);
  NERF_CHECK_LAUNCH("tmem_read_probe_kernel");
  return 0;
}

// ---- L2 -> SM streaming probe (development aid) -----------------------------------------------------------------
// Every CTA streams `iters` chunks of 16 KB from an L2-resident window into an 8-stage shared-memory ring with 1-D bulk
// copies (mode bit 0) and/or bulk-stores 16 KB chunks back into the window (mode bit 1); out[0] = cycles of the slowest
// CTA, out[1] = bytes moved per CTA.  Answers: how many bytes per cycle and SM does the L2 deliver to 148 TMA engines?
namespace nerf {
__global__ void __launch_bounds__(64, 1) l2_stream_kernel(unsigned long long* out, uint8_t* window, uint32_t window_bytes, int mode, int iters) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  const int csel = (mode >> 4) & 7;                          // chunk size: 0 -> 16 KB (default), else 1 KB << csel (2 KB .. 128 KB ring of 128 KB)
  const uint32_t kChunk = csel == 0 ? 16384u : (1024u << csel);
  const int kStages = (int)(131072u / kChunk) < 32 ? (int)(131072u / kChunk) : 32;
  const uint32_t bars = smem_base + 131072u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) tc::mbar_init(bars + 8 * i, 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const uint32_t n_chunks = window_bytes / kChunk;
  uint32_t c = (blockIdx.x * 977u) % n_chunks;
  const long long t0 = clock64();
  uint32_t phase = 0;
  int stage = 0;
  for (int i = 0; i < iters; ++i) {
    if (i >= kStages && (mode & 1)) tc::mbar_wait(bars + 8 * stage, phase ^ 1);   // the load that used this stage has landed
    if (mode & 2) {
      if (i >= kStages) tc::bulk_wait_read<7>();
      tc::bulk_s2g(window + (uint64_t)c * kChunk, smem_base + stage * kChunk, kChunk);
      tc::bulk_commit();
    }
    if (mode & 1) {
      tc::mbar_arrive_expect_tx(bars + 8 * stage, kChunk);
      tc::bulk_g2s(smem_base + stage * kChunk, window + (uint64_t)((c + n_chunks / 2) % n_chunks) * kChunk, kChunk, bars + 8 * stage);
    }
    c = (c + 148u) % n_chunks;
    if (++stage == kStages) {
      stage = 0;
      phase ^= 1;
    }
  }
  if (mode & 1)
    for (int i = 0; i < kStages && i < iters; ++i) {   // drain
      int sidx = (stage + i) % kStages;
      uint32_t ph = (sidx >= stage) ? (phase ^ 1) : phase;
      tc::mbar_wait(bars + 8 * sidx, ph);
    }
  if (mode & 2) tc::bulk_wait_all<0>();
  const unsigned long long dt = (unsigned long long)(clock64() - t0);
  atomicMax(out, dt);
  if (blockIdx.x == 0) out[1] = (unsigned long long)iters * kChunk * (((mode & 1) ? 1 : 0) + ((mode & 2) ? 1 : 0));
}
}  // namespace nerf

// LSU flavour of the probe: `n_warps` warps copy 16 KB chunks between the window and shared memory with 16-byte
// ld.shared + st.global (mode bit 2, "stores") and/or ld.global + st.shared (mode bit 3, "loads"), optionally while thread 0
// keeps the TMA ring of bulk LOADS busy (mode bit 0) -- can the LSU path carry the stash stores next to the weight ring?
namespace nerf {
__global__ void __launch_bounds__(288, 1) l2_stream_lsu_kernel(unsigned long long* out, uint8_t* window, uint32_t window_bytes, int mode, int iters,
                                                              int n_warps) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - tc::smem_u32(smem_raw));
  constexpr int kStages = 8;
  constexpr uint32_t kChunk = 16384;
  const uint32_t bars = smem_base + kStages * kChunk;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) tc::mbar_init(bars + 8 * i, 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  const uint32_t n_chunks = window_bytes / kChunk;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t0 = clock64();
  if (warp == 0) {
    if (lane == 0 && (mode & 1)) {   // TMA load ring
      uint32_t c = (blockIdx.x * 977u) % n_chunks, phase = 0;
      int stage = 0;
      for (int i = 0; i < iters; ++i) {
        if (i >= kStages) tc::mbar_wait(bars + 8 * stage, phase ^ 1);
        tc::mbar_arrive_expect_tx(bars + 8 * stage, kChunk);
        tc::bulk_g2s(smem_base + stage * kChunk, window + (uint64_t)c * kChunk, kChunk, bars + 8 * stage);
        c = (c + 148u) % n_chunks;
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      for (int i = 0; i < kStages && i < iters; ++i) {
        int sidx = (stage + i) % kStages;
        tc::mbar_wait(bars + 8 * sidx, (sidx >= stage) ? (phase ^ 1) : phase);
      }
    }
  } else if (warp - 1 < n_warps) {
    // each LSU warp moves its share of every chunk: 16 B per lane per instruction, fully coalesced (512 B per warp instruction)
    const int w = warp - 1;
    uint32_t c = (blockIdx.x * 977u + 31u) % n_chunks;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int i = 0; i < iters; ++i) {
      uint4* g = reinterpret_cast<uint4*>(window + (uint64_t)c * kChunk);
      uint4* sp = reinterpret_cast<uint4*>(smem_gen + (i % kStages) * kChunk);
      for (int q = w * 32 + lane; q < (int)(kChunk / 16); q += n_warps * 32) {
        if (mode & 4) g[q] = sp[q];
        if (mode & 8) { uint4 v = __ldcg(g + q); acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w; }
      }
      c = (c + 148u) % n_chunks;
    }
    if (acc.x == 0x12345678u && acc.y == 1u) out[2] = acc.z;   // keep the loads alive
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicMax(out, (unsigned long long)(clock64() - t0));
    if (blockIdx.x == 0) out[1] = (unsigned long long)iters * kChunk * (((mode & 1) ? 1 : 0) + ((mode & 4) ? 1 : 0) + ((mode & 8) ? 1 : 0));
  }
}
}  // namespace nerf

extern "C" int nerf_selftest_l2_stream_lsu(unsigned long long* out, void* window, uint32_t window_bytes, int mode, int iters, int n_ctas,
                                           int n_warps, void* stream) {
  using namespace nerf;
  NERF_CHECK_ARG(out && window && window_bytes >= (1u << 20) && iters > 0 && n_ctas > 0 && n_warps >= 1 && n_warps <= 8, "selftest_l2_stream_lsu: bad arguments");
  const int smem = 8 * 16384 + 256 + 1024;
  cudaError_t e = cudaFuncSetAttribute(l2_stream_lsu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  NERF_CHECK_ARG(e == cudaSuccess, "selftest_l2_stream_lsu: %s", cudaGetErrorString(e));
  l2_stream_lsu_kernel<<<n_ctas, 288, smem, static_cast<cudaStream_t>(stream)>>>(out, static_cast<uint8_t*>(window), window_bytes, mode, iters, n_warps);
  NERF_CHECK_LAUNCH("l2_stream_lsu_kernel");
  return 0;
}

extern "C" int nerf_selftest_l2_stream(unsigned long long* out, void* window, uint32_t window_bytes, int mode, int iters, int n_ctas, void* stream) {
  using namespace nerf;
  NERF_CHECK_ARG(out && window && window_bytes >= (1u << 20) && iters > 0 && n_ctas > 0, "selftest_l2_stream: bad arguments");
  const int smem = 8 * 16384 + 512 + 1024;
  cudaError_t e = cudaFuncSetAttribute(l2_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  NERF_CHECK_ARG(e == cudaSuccess, "selftest_l2_stream: %s", cudaGetErrorString(e));
  l2_stream_kernel<<<n_ctas, 64, smem, static_cast<cudaStream_t>(stream)>>>(out, static_cast<uint8_t*>(window), window_bytes, mode, iters);
  NERF_CHECK_LAUNCH("l2_stream_kernel");
  return 0;
}
