// Weight-gradient GEMM on tcgen05 with MN-major operands:
//     dW[co][ci][r][s] += sum over output pixels p of  dY[p][co] * X[tap_{r,s}(p)][ci]
// Per filter tap this is D[co, ci] = A^T B with A = dY tile [64 pixels][128 co] and B = X tile [64 pixels][64 ci]:
// both operands are channel-contiguous ("MN-major") exactly as TMA delivers an 8x8-pixel box of an NHWC tensor
// (64 rows of 128 B, SWIZZLE_128B; descriptor: SBO = 1024 B per 8 pixel rows, LBO = 8192 B between 64-channel
// blocks -- verified by the probe in probe.cu).  The pixel dimension is the GEMM K: each CTA owns a (co tile,
// ci tile, tap group) and a strided subset of the pixel tiles (split-K), keeps one fp32 accumulator per tap in
// TMEM for its whole life, and finally stores its partial tile to a workspace with plain vector stores; a small
// second kernel sums the split-K partials into the torch-layout gradient (scattered fp32 atomics from every CTA
// were the bottleneck of the first version: up to 7 M atomics per layer).  Stride-2 convs use the same parity
// tensor maps as the forward kernel; zero padding / ragged tiles are TMA out-of-bounds fills.
#include <cstdlib>
#include <cstring>
#include "train_kernels.cuh"
#include "ptx.cuh"

namespace pidnet {
namespace {

constexpr int kWgThreads = 192;  // warp 0: TMA producer, warp 1: MMA issuer (+TMEM), warps 2..5: epilogue
// pipeline depth per taps-per-CTA variant: stage = (2 + T) * 8 KB, kept under the 227 KB limit
template <int T> struct WgStages { static constexpr int value = T == 1 ? 6 : 4; };   // <= 160 KB: leaves room for the BatchNorm blocks of the main stream

template <int T>  // taps per CTA (1 or 3)
__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tc_kernel(const __grid_constant__ WgradParams p) {
  constexpr int kWgStages = WgStages<T>::value;
  constexpr int kABytes = 2 * 8192, kBBytes = T * 8192, kStageBytes = kABytes + kBBytes;
  constexpr int kCols = T == 1 ? 64 : 256;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_base = base + kWgStages * kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kWgStages + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * kWgStages);
  const uint32_t tmem_slot = done_bar + 8u;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen + kWgStages * kStageBytes + 8 * (2 * kWgStages + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ngroups = (p.ntaps + T - 1) / T;
  const int ci_tile = blockIdx.z / ngroups, grp = blockIdx.z % ngroups;
  const int co0 = blockIdx.y * 128, ci0 = ci_tile * 64;
  const int tap0 = grp * T;
  const int ntap = min(T, p.ntaps - tap0);
  const int per_img = p.tiles_w * p.tiles_h;
  const int m_tiles = p.N * per_img;
  const bool second_a = co0 + 64 < p.Cout;   // the upper 64-channel block exists (else rows 64..127 are unused)

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_gen;

  int my_tiles = 0;
  for (int t = blockIdx.x; t < m_tiles; t += gridDim.x) ++my_tiles;

  if (warp == 0) {
    if (elect_one()) {
      int it = 0;
      for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++it) {
        const int n = tile / per_img;
        const int rem = tile - n * per_img;
        const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
        const int w0 = tw * 8, h0 = th * 8;
        const int st = it % kWgStages;
        const uint32_t ph = (it / kWgStages) & 1;
        mbar_wait(empty_bar(st), ph ^ 1);
        mbar_arrive_expect_tx(full_bar(st), (second_a ? 2 : 1) * 8192 + ntap * 8192);
        const uint32_t sb = base + st * kStageBytes;
        tma_load_4d(sb, &p.tmY, full_bar(st), co0, w0, h0, n);
        if (second_a) tma_load_4d(sb + 8192, &p.tmY, full_bar(st), co0 + 64, w0, h0, n);
        for (int j = 0; j < ntap; ++j) {
          const uint32_t tp = p.taps[tap0 + j];
          const int dh = static_cast<int>((tp >> 8) & 0xFF) - 8, dw = static_cast<int>((tp >> 16) & 0xFF) - 8;
          tma_load_4d(sb + kABytes + j * 8192, &p.tmX[tp & 0xFF], full_bar(st), ci0, w0 + dw, h0 + dh, n);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // idesc: fp32 accumulate, bf16 x bf16, A and B MN-major
    constexpr uint32_t idesc = make_idesc_bf16(128, 64) | (1u << 15) | (1u << 16);
    constexpr uint64_t hi = (static_cast<uint64_t>(8192 >> 4) << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) |
                            (1ull << 46) | (2ull << 61);
    for (int it = 0; it < my_tiles; ++it) {
      const int st = it % kWgStages;
      const uint32_t ph = (it / kWgStages) & 1;
      mbar_wait(full_bar(st), ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sb = base + st * kStageBytes;
#pragma unroll
        for (int j = 0; j < T; ++j) {
          if (j < ntap) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t a_desc = hi | static_cast<uint64_t>(((sb + k * 2048) & 0x3FFFF) >> 4);
              const uint64_t b_desc = hi | static_cast<uint64_t>(((sb + kABytes + j * 8192 + k * 2048) & 0x3FFFF) >> 4);
              umma_bf16(tmem + j * 64, a_desc, b_desc, idesc, (it | k) != 0 ? 1u : 0u);
            }
          }
        }
        umma_commit(empty_bar(st));
        if (it == my_tiles - 1) umma_commit(done_bar);
      }
      __syncwarp();
    }
  } else if (my_tiles > 0) {
    // epilogue: lane quarter q of TMEM = co rows 32q..32q+31
    const int q = warp & 3;
    const int co = co0 + q * 32 + lane;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    // the split-K reduction may be scheduled from here on (it waits for this grid's completion itself): its launch latency
    // hides behind the accumulator read-out.  Not earlier -- waiting reduction blocks would hold the thread slots the
    // BatchNorm kernels of the main stream need.
    pdl_launch_dependents();
    // partial tile of this CTA: [row = co within the tile][tap j][64 ci]
    float* tile = p.ws + ((static_cast<size_t>(blockIdx.x) * gridDim.y + blockIdx.y) * gridDim.z + blockIdx.z) *
                             (static_cast<size_t>(T) * kWgradTileFloats);
    const int nci = min(64, p.Cin - ci0);
    if (T == 1 && p.direct) {
      // 1x1 conv: dW[co][ci_off + ci0 ..] is contiguous along ci -> the split-K sum is 16-byte vector reductions at L2 (the
      // separate reduction launch cost as much as these small GEMMs themselves: 7-21 us against 7-23 us)
      float* row = p.dW + static_cast<size_t>(co) * p.Cin_total + p.ci_off + ci0;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if (g * 32 >= nci) break;   // warp-uniform
        uint32_t v[32];
        tmem_ld32(tmem + g * 32 + (static_cast<uint32_t>(q * 32) << 16), v);
        tmem_ld_wait();
        if (co < p.Cout) {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (g * 32 + 4 * e < nci)   // (channel counts are multiples of 8)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(row + g * 32 + 4 * e), "f"(__uint_as_float(v[4 * e])),
                           "f"(__uint_as_float(v[4 * e + 1])), "f"(__uint_as_float(v[4 * e + 2])), "f"(__uint_as_float(v[4 * e + 3]))
                           : "memory");
        }
      }
    } else
    for (int j = 0; j < ntap; ++j) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if (g * 32 >= nci) break;   // warp-uniform
        uint32_t v[32];
        tmem_ld32(tmem + j * 64 + g * 32 + (static_cast<uint32_t>(q * 32) << 16), v);
        tmem_ld_wait();
        if (co < p.Cout) {
          float4* dst = reinterpret_cast<float4*>(tile + (static_cast<size_t>(q * 32 + lane) * T + j) * 64 + g * 32);
#pragma unroll
          for (int e = 0; e < 8; ++e)
            dst[e] = make_float4(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]), __uint_as_float(v[4 * e + 2]),
                                 __uint_as_float(v[4 * e + 3]));
        }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<kCols>(tmem);
}

// ---------------------------------------------------------------------------------------------------------------
// 3x3 stride-1 variant with a HALO patch: per 16x8-pixel output tile the CTA loads dY once and ONE 18x10-pixel box of X;
// the B operand of tap (r, s) and pixel row k is the 16 consecutive patch rows starting at ((k + r) * 18 + s) -- a
// descriptor start shifted by whole 128-byte rows inside the TMA-written patch (the swizzle is a function of absolute
// smem address bits, same trick as conv3_ws.cu).  Five taps per CTA (groups {0..4}, {5..8}: 5 x 64 fp32 TMEM columns), so X
// is read twice per ci tile instead of nine times and dY twice instead of three times: the tap-by-tap form above is
// L2->SM bandwidth bound on these layers, this one is bound by the MMA smem-operand rate.
constexpr int kHT = 5;
constexpr int kHStages = 3;
constexpr int kHABytes = 2 * 16384;                 // dY: two 64-co blocks of [128 px][128 B]
constexpr int kHPatchBytes = 18 * 10 * 128;         // 23040
constexpr int kHBBytes = 23552;                     // padded to a 1 KB multiple
constexpr int kHStageBytes = kHABytes + kHBBytes;   // 56320

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_halo_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_base = base + kHStages * kHStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kHStages + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * kHStages);
  const uint32_t tmem_slot = done_bar + 8u;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen + kHStages * kHStageBytes + 8 * (2 * kHStages + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ci_tile = blockIdx.z >> 1, grp = blockIdx.z & 1;
  const int co0 = blockIdx.y * 128, ci0 = ci_tile * 64;
  const int tap0 = grp * kHT;
  const int ntap = min(kHT, 9 - tap0);
  const int per_img = p.tiles_w * p.tiles_h;
  const int m_tiles = p.N * per_img;
  const bool second_a = co0 + 64 < p.Cout;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kHStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_gen;

  int my_tiles = 0;
  for (int t = blockIdx.x; t < m_tiles; t += gridDim.x) ++my_tiles;

  if (warp == 0) {
    if (elect_one()) {
      int it = 0;
      for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++it) {
        const int n = tile / per_img;
        const int rem = tile - n * per_img;
        const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
        const int w0 = tw * 16, h0 = th * 8;
        const int st = it % kHStages;
        const uint32_t ph = (it / kHStages) & 1;
        mbar_wait(empty_bar(st), ph ^ 1);
        mbar_arrive_expect_tx(full_bar(st), (second_a ? 2 : 1) * 16384 + kHPatchBytes);
        const uint32_t sb = base + st * kHStageBytes;
        tma_load_4d(sb, &p.tmY, full_bar(st), co0, w0, h0, n);
        if (second_a) tma_load_4d(sb + 16384, &p.tmY, full_bar(st), co0 + 64, w0, h0, n);
        tma_load_4d(sb + kHABytes, &p.tmX[0], full_bar(st), ci0, w0 - 1, h0 - 1, n);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, 64) | (1u << 15) | (1u << 16);
    // A: 64-co blocks 16384 B apart (LBO), 8-pixel atoms 1024 B apart (SBO); B: one 64-ci block
    constexpr uint64_t a_hi = (static_cast<uint64_t>(16384 >> 4) << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) |
                              (1ull << 46) | (2ull << 61);
    constexpr uint64_t b_hi = (static_cast<uint64_t>(8192 >> 4) << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) |
                              (1ull << 46) | (2ull << 61);
    for (int it = 0; it < my_tiles; ++it) {
      const int st = it % kHStages;
      const uint32_t ph = (it / kHStages) & 1;
      mbar_wait(full_bar(st), ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sb = base + st * kHStageBytes;
#pragma unroll
        for (int j = 0; j < kHT; ++j) {
          if (j < ntap) {
            const int tap = tap0 + j, r = tap / 3, s = tap - 3 * r;
#pragma unroll
            for (int k = 0; k < 8; ++k) {   // k = pixel row of the tile = 16 pixels = one K16 step
              const uint64_t a_desc = a_hi | static_cast<uint64_t>(((sb + k * 2048) & 0x3FFFF) >> 4);
              const uint64_t b_desc = b_hi | static_cast<uint64_t>(((sb + kHABytes + ((k + r) * 18 + s) * 128) & 0x3FFFF) >> 4);
              umma_bf16(tmem + j * 64, a_desc, b_desc, idesc, (it | k) != 0 ? 1u : 0u);
            }
          }
        }
        umma_commit(empty_bar(st));
        if (it == my_tiles - 1) umma_commit(done_bar);
      }
      __syncwarp();
    }
  } else if (my_tiles > 0) {
    const int q = warp & 3;
    const int co = co0 + q * 32 + lane;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    // the split-K reduction may be scheduled from here on (it waits for this grid's completion itself): its launch latency
    // hides behind the accumulator read-out.  Not earlier -- waiting reduction blocks would hold the thread slots the
    // BatchNorm kernels of the main stream need.
    pdl_launch_dependents();
    float* tile = p.ws + ((static_cast<size_t>(blockIdx.x) * gridDim.y + blockIdx.y) * gridDim.z + blockIdx.z) *
                             (static_cast<size_t>(kHT) * kWgradTileFloats);
    const int nci = min(64, p.Cin - ci0);
    for (int j = 0; j < ntap; ++j) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if (g * 32 >= nci) break;   // warp-uniform
        uint32_t v[32];
        tmem_ld32(tmem + j * 64 + g * 32 + (static_cast<uint32_t>(q * 32) << 16), v);
        tmem_ld_wait();
        if (co < p.Cout) {
          float4* dst = reinterpret_cast<float4*>(tile + (static_cast<size_t>(q * 32 + lane) * kHT + j) * 64 + g * 32);
#pragma unroll
          for (int e = 0; e < 8; ++e)
            dst[e] = make_float4(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]), __uint_as_float(v[4 * e + 2]),
                                 __uint_as_float(v[4 * e + 3]));
        }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}
constexpr size_t kHSmem = static_cast<size_t>(kHStages) * kHStageBytes + 256 + 1024;

// ---------------------------------------------------------------------------------------------------------------
// 3x3 stride-1, Cout <= 64 per tile ("stacked taps"): with A = dY the M = 128 rows of the MMA are half empty for these layers
// (most of PIDNet-S).  Here the roles are swapped: A = X, and TWO filter taps share one MMA -- the M = 128 rows are the 64
// input channels of tap t0 followed by the 64 input channels of tap t1, which in the MN-major descriptor is just a leading
// byte offset (LBO) equal to the distance between the two taps' windows inside the halo patch ((dr * 18 + ds) * 128 B);
// B = dY (N = 64 output channels).  Five accumulators ([tap pair][ci] x co) cover all nine taps, so ONE CTA does what two
// CTAs of wgrad_halo_kernel do with the same 40 MMAs per tile: half the tensor work and half the dY / patch traffic.
constexpr int kSStages = 4;
constexpr int kSABytes = 16384;                     // dY: one 64-co block of [128 px][128 B]
constexpr int kSStageBytes = kSABytes + kHBBytes;   // 39936
constexpr int kSPairs = 5;
__global__ void __launch_bounds__(kWgThreads, 1) wgrad_stack_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_base = base + kSStages * kSStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kSStages + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * kSStages);
  const uint32_t tmem_slot = done_bar + 8u;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen + kSStages * kSStageBytes + 8 * (2 * kSStages + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int co0 = blockIdx.y * 64, ci0 = blockIdx.z * 64;
  const int per_img = p.tiles_w * p.tiles_h;
  const int m_tiles = p.N * per_img;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kSStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_gen;

  int my_tiles = 0;
  for (int t = blockIdx.x; t < m_tiles; t += gridDim.x) ++my_tiles;

  if (warp == 0) {
    if (elect_one()) {
      int it = 0;
      for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++it) {
        const int n = tile / per_img;
        const int rem = tile - n * per_img;
        const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
        const int w0 = tw * 16, h0 = th * 8;
        const int st = it % kSStages;
        const uint32_t ph = (it / kSStages) & 1;
        mbar_wait(empty_bar(st), ph ^ 1);
        mbar_arrive_expect_tx(full_bar(st), kSABytes + kHPatchBytes);
        const uint32_t sb = base + st * kSStageBytes;
        tma_load_4d(sb, &p.tmY, full_bar(st), co0, w0, h0, n);
        tma_load_4d(sb + kSABytes, &p.tmX[0], full_bar(st), ci0, w0 - 1, h0 - 1, n);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, 64) | (1u << 15) | (1u << 16);
    // B (dY): one 64-co block, 8-pixel atoms 1024 B apart
    constexpr uint64_t b_hi = (static_cast<uint64_t>(8192 >> 4) << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) |
                              (1ull << 46) | (2ull << 61);
    constexpr uint64_t a_hi0 = (static_cast<uint64_t>(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);   // + LBO per tap pair
    for (int it = 0; it < my_tiles; ++it) {
      const int st = it % kSStages;
      const uint32_t ph = (it / kSStages) & 1;
      mbar_wait(full_bar(st), ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sb = base + st * kSStageBytes;
#pragma unroll
        for (int j = 0; j < kSPairs; ++j) {
          const int t0 = 2 * j, t1 = min(2 * j + 1, 8);
          const int r0 = t0 / 3, s0 = t0 - 3 * r0, r1 = t1 / 3, s1 = t1 - 3 * r1;
          const int delta = ((r1 - r0) * 18 + (s1 - s0)) * 128;          // > 0 for a real pair; the lone tap 8 reads itself twice
          const uint64_t a_hi = a_hi0 | (static_cast<uint64_t>((delta > 0 ? delta : 128) >> 4) << 16);
#pragma unroll
          for (int k = 0; k < 8; ++k) {   // k = pixel row of the tile = 16 pixels = one K16 step
            const uint64_t a_desc = a_hi | static_cast<uint64_t>(((sb + kSABytes + ((k + r0) * 18 + s0) * 128) & 0x3FFFF) >> 4);
            const uint64_t b_desc = b_hi | static_cast<uint64_t>(((sb + k * 2048) & 0x3FFFF) >> 4);
            umma_bf16(tmem + j * 64, a_desc, b_desc, idesc, (it | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(empty_bar(st));
        if (it == my_tiles - 1) umma_commit(done_bar);
      }
      __syncwarp();
    }
  } else if (my_tiles > 0) {
    // epilogue: lane quarter q of TMEM = rows 32q..32q+31 = (tap of the pair, ci); columns = co
    const int q = warp & 3;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    // the split-K reduction may be scheduled from here on (it waits for this grid's completion itself): its launch latency
    // hides behind the accumulator read-out.  Not earlier -- waiting reduction blocks would hold the thread slots the
    // BatchNorm kernels of the main stream need.
    pdl_launch_dependents();
    float* tile = p.ws + ((static_cast<size_t>(blockIdx.x) * gridDim.y + blockIdx.y) * gridDim.z + blockIdx.z) *
                             (static_cast<size_t>(kSPairs) * kWgradTileFloats);
    const int nco = min(64, p.Cout - co0);
    for (int j = 0; j < kSPairs; ++j) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if (g * 32 >= nco) break;   // warp-uniform
        uint32_t v[32];
        tmem_ld32(tmem + j * 64 + g * 32 + (static_cast<uint32_t>(q * 32) << 16), v);
        tmem_ld_wait();
        float4* dst = reinterpret_cast<float4*>(tile + (static_cast<size_t>(j) * 128 + q * 32 + lane) * 64 + g * 32);
#pragma unroll
        for (int e = 0; e < 8; ++e)
          dst[e] = make_float4(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]), __uint_as_float(v[4 * e + 2]),
                               __uint_as_float(v[4 * e + 3]));
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}
constexpr size_t kSSmem = static_cast<size_t>(kSStages) * kSStageBytes + 256 + 1024;

// One warp's share of a split-K sum: splits g, g + NW, g + 2 NW, ... with EIGHT loads in flight (the ~19 loads of a warp are
// L2 round trips; two at a time made the reduction as long as the GEMM it follows)
template <int NW>
__device__ __forceinline__ float sum_splits(const float* src, size_t split_stride, int g, int splits, int wide = 1) {
  float acc = 0.f;
  if (!wide) {
    float s0 = 0.f, s1 = 0.f;
    int sp = g;
    for (; sp + NW < splits; sp += 2 * NW) { s0 += src[sp * split_stride]; s1 += src[(sp + NW) * split_stride]; }
    if (sp < splits) s0 += src[sp * split_stride];
    return s0 + s1;
  }
  for (int base = g; base < splits; base += 8 * NW) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int sp = base + u * NW;
      v[u] = sp < splits ? __ldcg(src + static_cast<size_t>(sp) * split_stride) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += v[u];
  }
  return acc;
}

// ---------------------------------------------------------------------------------------------------------------
// Split-K reductions.  A block owns 32 consecutive gradient elements; its eight warps each sum every eighth split (coalesced
// 128-byte rows of the workspace) and the partial sums meet in shared memory -- one thread per element walking all ~148 splits
// serially was latency-bound at ~12 us per layer, as long as the GEMM itself.  The same two bodies serve the per-layer launches
// (programmatic dependent launch behind their GEMM) and the batched launch that sums every deferred layer of a backward range
// at once (wgrad_reduce_all_kernel: ~46 latency-bound launches of 7-15 us become one bandwidth-bound launch).
constexpr int kRedWarps = 8;
// partial tiles of wgrad_stack_kernel ([pair][tap-in-pair * 64 + ci][co]): co fastest (coalesced workspace reads)
__device__ __forceinline__ void reduce_stack_body(const WgReduceJob& p, unsigned blk, float (*red)[32], int wide) {
  const long total = static_cast<long>(9) * p.Cin * p.Cout;
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const long idx = static_cast<long>(blk) * 32 + lane;
  const bool live = idx < total;
  const long id = live ? idx : total - 1;
  const int co = static_cast<int>(id % p.Cout);
  const int ci = static_cast<int>((id / p.Cout) % p.Cin);
  const int tap = static_cast<int>(id / (static_cast<long>(p.Cout) * p.Cin));
  const int j = tap >> 1, tp = tap & 1;
  const size_t tile_floats = static_cast<size_t>(kSPairs) * kWgradTileFloats;
  const int ci_tiles = p.nz;
  const float* src = p.ws + (static_cast<size_t>(co >> 6) * ci_tiles + (ci >> 6)) * tile_floats +
                     (static_cast<size_t>(j) * 128 + tp * 64 + (ci & 63)) * 64 + (co & 63);
  const size_t split_stride = static_cast<size_t>(p.co_tiles) * ci_tiles * tile_floats;
  red[g][lane] = sum_splits<kRedWarps>(src, split_stride, g, p.splits, wide);
  __syncthreads();
  if (g == 0 && live) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < kRedWarps; ++q) t += red[q][lane];
    const int r = tap / 3, q3 = tap - 3 * r;
    p.dW[(static_cast<size_t>(co) * p.Cin_total + p.ci_off + ci) * 9 + r * 3 + q3] += t;
  }
}
// partial tiles of wgrad_tc_kernel<T> / wgrad_halo_kernel ([co row][tap j][64 ci]):
// dW[co][ci_off + ci][r][s] += sum over splits of the partial tiles
__device__ __forceinline__ void reduce_tc_body(const WgReduceJob& p, unsigned blk, float (*red)[32], int wide) {
  const long total = static_cast<long>(p.Cout) * p.ntaps * p.Cin;
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const long idx = static_cast<long>(blk) * 32 + lane;
  const bool live = idx < total;
  const long id = live ? idx : total - 1;
  const int ci = static_cast<int>(id % p.Cin);
  const int tapi = static_cast<int>((id / p.Cin) % p.ntaps);
  const int co = static_cast<int>(id / (static_cast<long>(p.Cin) * p.ntaps));
  const int T = p.T;
  const int ngroups = (p.ntaps + T - 1) / T;
  const int z = (ci >> 6) * ngroups + tapi / T, j = tapi % T;
  const size_t tile_floats = static_cast<size_t>(T) * kWgradTileFloats;
  const float* src = p.ws + (static_cast<size_t>(co >> 7) * p.nz + z) * tile_floats +
                     (static_cast<size_t>(co & 127) * T + j) * 64 + (ci & 63);
  const size_t split_stride = static_cast<size_t>(p.co_tiles) * p.nz * tile_floats;
  red[g][lane] = sum_splits<kRedWarps>(src, split_stride, g, p.splits, wide);
  __syncthreads();
  if (g == 0 && live) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < kRedWarps; ++q) t += red[q][lane];
    const int r = p.tap_rs[tapi] >> 4, s = p.tap_rs[tapi] & 0xF;
    p.dW[(static_cast<size_t>(co) * p.Cin_total + p.ci_off + ci) * (p.k * p.k) + r * p.k + s] += t;
  }
}
__global__ void __launch_bounds__(32 * kRedWarps) wgrad_reduce_kernel(const __grid_constant__ WgReduceJob job, int wide) {
  __shared__ float red[kRedWarps][32];
  pdl_wait();   // launched while the GEMM drains (programmatic dependent launch); its partial tiles are complete after this
  if (job.stack) reduce_stack_body(job, blockIdx.x, red, wide);
  else reduce_tc_body(job, blockIdx.x, red, wide);
}
// every deferred layer of a backward range in one launch: block_start[i] - base = first block of job i
__global__ void __launch_bounds__(32 * kRedWarps) wgrad_reduce_all_kernel(const WgReduceJob* __restrict__ jobs,
                                                                          const unsigned* __restrict__ block_start, int njobs,
                                                                          unsigned base, int wide) {
  __shared__ float red[kRedWarps][32];
  __shared__ WgReduceJob job;
  const unsigned blk = blockIdx.x + base;
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {   // last job whose first block is <= blk
    const int mid = (lo + hi + 1) >> 1;
    if (block_start[mid] <= blk) lo = mid; else hi = mid - 1;
  }
  if (threadIdx.x < sizeof(WgReduceJob) / 4) reinterpret_cast<uint32_t*>(&job)[threadIdx.x] = reinterpret_cast<const uint32_t*>(jobs + lo)[threadIdx.x];
  __syncthreads();
  if (job.stack) reduce_stack_body(job, blk - block_start[lo], red, wide);
  else reduce_tc_body(job, blk - block_start[lo], red, wide);
}

template <int T>
size_t wg_smem() { return static_cast<size_t>(WgStages<T>::value) * (2 * 8192 + T * 8192) + 256 + 1024; }

}  // namespace

cudaError_t wgrad_tc_init() {
  cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(wg_smem<1>()));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(wgrad_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(wg_smem<3>()));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kHSmem));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(wgrad_stack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSSmem));
}

bool wgrad_reduce_job(const WgradLaunch& L, WgReduceJob* job) {
  if (L.p.direct) return false;
  WgReduceJob j;
  std::memset(&j, 0, sizeof(j));
  j.ws = L.p.ws; j.dW = L.p.dW;
  j.Cout = L.p.Cout; j.Cin = L.p.Cin; j.Cin_total = L.p.Cin_total; j.ci_off = L.p.ci_off; j.k = L.p.k; j.ntaps = L.p.ntaps;
  for (int i = 0; i < L.p.ntaps && i < kConvMaxTaps; ++i)
    j.tap_rs[i] = static_cast<unsigned char>((((L.p.taps[i] >> 24) & 0xF) << 4) | ((L.p.taps[i] >> 28) & 0xF));
  j.stack = L.p.halo == 2 ? 1 : 0;
  j.T = L.p.halo == 2 ? kSPairs : (L.p.halo ? kHT : L.taps_per_group);
  j.splits = static_cast<int>(L.grid.x); j.co_tiles = static_cast<int>(L.grid.y); j.nz = static_cast<int>(L.grid.z);
  *job = j;
  return true;
}
unsigned wgrad_reduce_job_blocks(const WgReduceJob& j) {
  const long total = j.stack ? static_cast<long>(9) * j.Cin * j.Cout : static_cast<long>(j.Cout) * j.ntaps * j.Cin;
  return static_cast<unsigned>((total + 31) / 32);
}
static int reduce_wide() {
  static const int red_wide = [] { const char* v = std::getenv("PIDNET_WG_WIDE"); return v && v[0] == '0' ? 0 : 1; }();
  return red_wide;
}
cudaError_t wgrad_reduce_all_launch(const WgReduceJob* dev_jobs, const unsigned* dev_block_start, int njobs, unsigned base,
                                    unsigned total_blocks, cudaStream_t st) {
  if (njobs <= 0 || total_blocks == 0) return cudaSuccess;
  wgrad_reduce_all_kernel<<<total_blocks, 32 * kRedWarps, 0, st>>>(dev_jobs, dev_block_start, njobs, base, reduce_wide());
  return cudaGetLastError();
}

cudaError_t wgrad_tc_launch(const WgradLaunch& L, cudaStream_t st, bool defer_reduce) {
  static const bool red_pdl = [] { const char* v = std::getenv("PIDNET_WG_PDL"); return !(v && v[0] == '0'); }();
  if (L.p.halo == 2) {   // stacked taps: grid (splits, co tiles of 64, ci tiles of 64), all nine taps per CTA
    if (L.p.ntaps != 9 || L.taps_per_group != 2 * kSPairs) return cudaErrorInvalidValue;
    wgrad_stack_kernel<<<L.grid, kWgThreads, kSSmem, st>>>(L.p);
  } else if (L.p.halo) {
    if (L.taps_per_group != kHT || L.p.ntaps != 9 || (L.grid.z & 1)) return cudaErrorInvalidValue;
    wgrad_halo_kernel<<<L.grid, kWgThreads, kHSmem, st>>>(L.p);
  } else if (L.taps_per_group == 1) wgrad_tc_kernel<1><<<L.grid, kWgThreads, wg_smem<1>(), st>>>(L.p);
  else if (L.taps_per_group == 3) wgrad_tc_kernel<3><<<L.grid, kWgThreads, wg_smem<3>(), st>>>(L.p);
  else return cudaErrorInvalidValue;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  WgReduceJob job;
  if (!wgrad_reduce_job(L, &job)) return cudaSuccess;   // direct mode: the GEMM reduced its split-K partials into dW itself
  if (defer_reduce) return cudaSuccess;                 // the caller sums this layer with wgrad_reduce_all_launch
  const unsigned blocks = wgrad_reduce_job_blocks(job);
  if (!red_pdl) {
    wgrad_reduce_kernel<<<blocks, 32 * kRedWarps, 0, st>>>(job, reduce_wide());
    return cudaGetLastError();
  }
  return launch_pdl(wgrad_reduce_kernel, dim3(blocks), dim3(32 * kRedWarps), 0, st, job, reduce_wide());
}

}  // namespace pidnet
