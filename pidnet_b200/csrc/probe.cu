// Hardware probe (kept as a documented experiment, not on the product path): does tcgen05.mma read a
// K-major SWIZZLE_128B operand correctly when the descriptor start address is shifted by whole
// 128-byte rows inside a TMA-written halo patch and the 8-row groups are SBO = 1280 B apart
// (10-pixel patch rows), i.e. is the swizzle XOR a function of absolute smem address bits?
// If yes, one (TH+2)x(TW+2) patch load can feed all nine 3x3 taps.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "ptx.cuh"

namespace pidnet {

struct ProbeParams {
  CUtensorMap tmA;  // [1,18,10,64] bf16 NHWC, box {64,10,18,1}, SWIZZLE_128B
  CUtensorMap tmB;  // [64][64] bf16, box {64,64}, SWIZZLE_128B
  int r, s, mode;   // mode 1: also program base_offset = (start >> 7) & 7
  float* out;       // [128][64]
};

__global__ void __launch_bounds__(128) halo_probe_kernel(const __grid_constant__ ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t patch = base;            // 180 rows x 128 B = 23040 B
  const uint32_t wts = base + 23552;      // 1024-aligned
  const uint32_t bar = wts + 8192;
  const uint32_t bar2 = bar + 8;
  const uint32_t slot = bar + 16;
  volatile uint32_t* slot_gen = reinterpret_cast<volatile uint32_t*>(gen + 23552 + 8192 + 16);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar2, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<64>(slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_gen;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 23040 + 8192);
    tma_load_4d(patch, &p.tmA, bar, 0, 0, 0, 0);
    tma_load_2d(wts, &p.tmB, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc_bf16(128, 64);
    const uint32_t a_start = patch + (p.r * 10 + p.s) * 128;
    uint64_t a_desc = 0;
    a_desc |= static_cast<uint64_t>((a_start & 0x3FFFF) >> 4);
    a_desc |= static_cast<uint64_t>(1) << 16;
    a_desc |= static_cast<uint64_t>(1280 >> 4) << 32;  // SBO: one tile row (8 px) per 10-px patch row
    a_desc |= static_cast<uint64_t>(1) << 46;
    if (p.mode == 1) a_desc |= static_cast<uint64_t>((a_start >> 7) & 7) << 49;
    a_desc |= static_cast<uint64_t>(2) << 61;
    const uint64_t b_desc = make_kmajor_desc(wts, 128);
    for (int k = 0; k < 4; ++k) umma_bf16(tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, k != 0);
    umma_commit(bar2);
  }
  __syncwarp();
  mbar_wait(bar2, 0);
  tc_fence_after();
  const uint32_t t_row = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  for (int g = 0; g < 2; ++g) {
    uint32_t v[32];
    tmem_ld32(t_row + g * 32, v);
    tmem_ld_wait();
    for (int e = 0; e < 32; ++e) p.out[threadIdx.x * 64 + g * 32 + e] = __uint_as_float(v[e]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<64>(tmem);
}

cudaError_t halo_probe_launch(const ProbeParams& p, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(halo_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
  if (e != cudaSuccess) return e;
  halo_probe_kernel<<<1, 128, 40960, st>>>(p);
  return cudaGetLastError();
}


// ---- MMA rate probe: one CTA per SM issues `iters` back-to-back tcgen05.mma (M=128, N, K=16, SS operands in
// SWIZZLE_128B smem, garbage data) and reports cycles per MMA (clock64 around issue + final commit wait).
template <int N>
__global__ void __launch_bounds__(128) mma_rate_kernel(int iters, int distinct, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a_base = base;                 // 9 x 16 KB "taps"
  const uint32_t b_base = base + 9 * 16384;     // 9 x N*128 B
  const uint32_t bar = b_base + 9 * N * 128;
  const uint32_t slot = bar + 8;
  volatile uint32_t* slot_gen = reinterpret_cast<volatile uint32_t*>(gen + 9 * 16384 + 9 * N * 128 + 8);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc<(N < 32 ? 32 : N)>(slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_gen;
  if (warp == 0) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, N);
      const long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const int t = distinct ? (i % 9) : 0;
        const uint64_t a_desc = make_kmajor_desc(a_base + t * 16384, 128);
        const uint64_t b_desc = make_kmajor_desc(b_base + t * N * 128, 128);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, 1u);
      }
      umma_commit(bar);
      mbar_wait(bar, 0);
      const long long t1 = clock64();
      out[blockIdx.x] = t1 - t0;
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<(N < 32 ? 32 : N)>(tmem);
}

// ---- MN-major operand probe (wgrad shape): D[co, ci] = sum_p A[p][co] * B[p][ci], both operands loaded by TMA
// from row-major [pixels][channels] matrices (channel-contiguous == "MN-major"), SWIZZLE_128B.
struct MnProbeParams {
  CUtensorMap tmA;  // [64 pixels][128 co] bf16, box {64 ch, 64 px}
  CUtensorMap tmB;  // [64 pixels][64 ci]  bf16, box {64 ch, 64 px}
  int lbo_a, sbo, variant;
  float* out;       // [128][64]
};
__global__ void __launch_bounds__(128) mn_probe_kernel(const __grid_constant__ MnProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a0 = base, a1 = base + 8192, b0 = base + 16384;   // three 64 x 128 B tiles
  const uint32_t bar = base + 24576, bar2 = bar + 8, slot = bar + 16;
  volatile uint32_t* slot_gen = reinterpret_cast<volatile uint32_t*>(gen + 24576 + 16);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc<64>(slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_gen;
  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(bar, 3 * 8192);
      tma_load_2d(a0, &p.tmA, bar, 0, 0);
      tma_load_2d(a1, &p.tmA, bar, 64, 0);
      tma_load_2d(b0, &p.tmB, bar, 0, 0);
      mbar_wait(bar, 0);
      tc_fence_after();
      // idesc: f32 accum, bf16 x bf16, A and B MN-major (bits 15, 16)
      const uint32_t idesc = make_idesc_bf16(128, 64) | (1u << 15) | (1u << 16);
      auto mk = [&](uint32_t addr, uint32_t lbo, uint32_t sbo) {
        uint64_t d = 0;
        d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
        d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
        d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
        d |= 1ull << 46;
        d |= 2ull << 61;   // SWIZZLE_128B
        return d;
      };
      for (int k = 0; k < 4; ++k) {
        const uint32_t koff = k * 2048;   // 16 pixel rows x 128 B
        umma_bf16(tmem, mk(a0 + koff, p.lbo_a, p.sbo), mk(b0 + koff, p.lbo_a, p.sbo), idesc, k != 0);
      }
      umma_commit(bar2);
    }
    __syncwarp();
  }
  mbar_wait(bar2, 0);
  tc_fence_after();
  const uint32_t t_row = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  for (int g = 0; g < 2; ++g) {
    uint32_t v[32];
    tmem_ld32(t_row + g * 32, v);
    tmem_ld_wait();
    for (int e = 0; e < 32; ++e) p.out[threadIdx.x * 64 + g * 32 + e] = __uint_as_float(v[e]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<64>(tmem);
}
cudaError_t mn_probe_launch(const MnProbeParams& p, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(mn_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  if (e != cudaSuccess) return e;
  mn_probe_kernel<<<1, 128, 32768, st>>>(p);
  return cudaGetLastError();
}

// ---- CTA-pair (cta_group::2) probes.  (1) correctness: D[256][64] = A[256][64] * B[64][64]^T with each CTA of the pair
// TMA-loading its own 128 rows of A and HALF (32 rows) of B, both signalling the leader's mbarrier; documents which half
// of B each CTA supplies and that the accumulator rows of CTA r are rows r*128.. of D.  (2) rate: cycles per M256 x N x K16.
struct PairProbeParams {
  CUtensorMap tmA;  // [256][64] bf16, box {64, 128}, SWIZZLE_128B
  CUtensorMap tmB;  // [64][64]  bf16, box {64, 32},  SWIZZLE_128B
  int swap_b;       // 1: CTA r loads B rows (1-r)*32.. (to test the convention)
  float* out;       // [256][64]
};
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) pair_probe_kernel(const __grid_constant__ PairProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a_s = base, b_s = base + 16384;
  const uint32_t bar_full = base + 20480, bar_done = bar_full + 8, slot = bar_full + 16;
  volatile uint32_t* slot_gen = reinterpret_cast<volatile uint32_t*>(gen + 20480 + 16);
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) { mbar_init(bar_full, 2); mbar_init(bar_done, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc_pair<64>(slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem = *slot_gen;
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t leader_full = mapa_u32(bar_full, 0);
      mbar_arrive_expect_tx_cluster(leader_full, 16384 + 4096);
      tma_load_2d_pair(a_s, &p.tmA, leader_full, 0, static_cast<int>(rank) * 128);
      tma_load_2d_pair(b_s, &p.tmB, leader_full, 0, static_cast<int>(p.swap_b ? 1 - rank : rank) * 32);
      if (rank == 0) {
        mbar_wait(bar_full, 0);
        tc_fence_after();
        constexpr uint32_t idesc = make_idesc_bf16(256, 64);
        const uint64_t a_desc = make_kmajor_desc(a_s, 128), b_desc = make_kmajor_desc(b_s, 128);
        for (int k = 0; k < 4; ++k) umma_bf16_pair(tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, k != 0);
        umma_commit_pair(bar_done, 3);
      }
    }
    __syncwarp();
  }
  mbar_wait(bar_done, 0);
  tc_fence_after();
  const uint32_t t_row = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  for (int g = 0; g < 2; ++g) {
    uint32_t v[32];
    tmem_ld32(t_row + g * 32, v);
    tmem_ld_wait();
    for (int e = 0; e < 32; ++e) p.out[(rank * 128 + threadIdx.x) * 64 + g * 32 + e] = __uint_as_float(v[e]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 1) tmem_dealloc_pair<64>(tmem);
}
cudaError_t pair_probe_launch(const PairProbeParams& p, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(pair_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  if (e != cudaSuccess) return e;
  pair_probe_kernel<<<2, 128, 32768, st>>>(p);
  return cudaGetLastError();
}

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) mma_rate_pair_kernel(int iters, int distinct, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a_base = base;                 // 9 x 16 KB "taps" (this CTA's 128 rows)
  const uint32_t b_base = base + 9 * 16384;     // 9 x (N/2)*128 B (this CTA's half of B)
  const uint32_t bar = b_base + 9 * (N / 2) * 128;
  const uint32_t slot = bar + 8;
  volatile uint32_t* slot_gen = reinterpret_cast<volatile uint32_t*>(gen + 9 * 16384 + 9 * (N / 2) * 128 + 8);
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc_pair<(N < 32 ? 32 : N)>(slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem = *slot_gen;
  long long t0 = 0;
  if (warp == 0 && rank == 0) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(256, N);
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const int t = distinct ? (i % 9) : 0;
        const uint64_t a_desc = make_kmajor_desc(a_base + t * 16384, 128);
        const uint64_t b_desc = make_kmajor_desc(b_base + t * (N / 2) * 128, 128);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_pair(tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, 1u);
      }
      umma_commit_pair(bar, 3);
    }
    __syncwarp();
  }
  if (warp == 0) {
    mbar_wait(bar, 0);
    if (rank == 0 && t0 != 0) out[blockIdx.x >> 1] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 1) tmem_dealloc_pair<(N < 32 ? 32 : N)>(tmem);
}
cudaError_t mma_rate_pair_launch(int N, int iters, int distinct, int pairs, long long* out, cudaStream_t st) {
  const int smem = 9 * 16384 + 9 * (N / 2) * 128 + 64 + 1024;
#define PIDNET_RATE2(NN)                                                                                           \
  if (N == NN) {                                                                                                   \
    cudaError_t e = cudaFuncSetAttribute(mma_rate_pair_kernel<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
    if (e != cudaSuccess) return e;                                                                                \
    mma_rate_pair_kernel<NN><<<2 * pairs, 128, smem, st>>>(iters, distinct, out);                                  \
    return cudaGetLastError();                                                                                     \
  }
  PIDNET_RATE2(32) PIDNET_RATE2(64) PIDNET_RATE2(128) PIDNET_RATE2(256)
#undef PIDNET_RATE2
  return cudaErrorInvalidValue;
}

cudaError_t mma_rate_launch(int N, int iters, int distinct, int blocks, long long* out, cudaStream_t st) {
  const int smem = 9 * 16384 + 9 * N * 128 + 64 + 1024;
#define PIDNET_RATE(NN)                                                                                       \
  if (N == NN) {                                                                                              \
    cudaError_t e = cudaFuncSetAttribute(mma_rate_kernel<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
    if (e != cudaSuccess) return e;                                                                           \
    mma_rate_kernel<NN><<<blocks, 128, smem, st>>>(iters, distinct, out);                                     \
    return cudaGetLastError();                                                                                \
  }
  PIDNET_RATE(32) PIDNET_RATE(64) PIDNET_RATE(128) PIDNET_RATE(256)
#undef PIDNET_RATE
  return cudaErrorInvalidValue;
}

}  // namespace pidnet
