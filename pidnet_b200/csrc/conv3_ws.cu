// conv3x3 (stride 1, pad 1) on tcgen05 -- weight-stationary, halo-patch, persistent.
//
// The generic kernel (conv_tc.cu) re-fetches the input once per filter tap (9x) and the whole weight
// matrix once per 128-pixel tile; measured on B200 that makes every PIDNet 3x3 layer L2->SM-bandwidth
// bound (~9-10 TB/s) with the tensor pipe idle.  This kernel removes both re-reads:
//   * weights: each persistent CTA owns ONE Cout tile and keeps all 9 x Cin x BN weights resident in
//     shared memory (loaded once by TMA);
//   * activations: a 16x8-pixel output tile needs an 18x10-pixel halo patch; ONE TMA box load per
//     64/32-channel chunk brings it in (zero-filled outside the image == the conv padding), and the
//     nine taps are nine tcgen05 smem descriptors whose start address is shifted by whole pixel rows
//     inside the patch (row stride 10 px -> SBO = 10 * row bytes).  The hardware swizzle is a function
//     of absolute smem address bits, so TMA's SWIZZLE_128B/64B write pattern and the shifted reads
//     agree (verified by csrc/probe.cu).
//   * accumulators are double-buffered in TMEM: the 4 epilogue warps drain tile i (bias, residual,
//     ReLU, bf16, swizzled smem staging, TMA store) while the MMA warp already runs tile i+1.
//   * a dedicated store warp issues the TMA stores and signals when a staging buffer has drained, so
//     the epilogue warps never wait on a store and the residual prefetch of tile i+2 starts early.
//   * even and odd tiles run on two independent lanes: each lane has its own MMA-issuing warp, TMEM
//     accumulator, 4-warp epilogue group and staging buffer, so one lane's per-tile barrier/commit
//     overhead and its ~2000-cycle epilogue hide behind the other lane's MMAs (ncu: with one lane both
//     the issuing warp and the epilogue group sat on the critical path).
// Warp roles: 0 = TMA producer, 1 and 10 = MMA issuers (lane 0 / 1; warp 1 also owns the TMEM allocation),
// 2..5 and 6..9 = epilogue groups (lane 0 / 1), 11 = TMA store.
//
// PAIR = true runs the same pipeline on a CTA PAIR (2-CTA cluster = the two SMs of a TPC) with tcgen05 cta_group::2: one
// M = 256 MMA covers the 128-pixel tiles of BOTH CTAs; each CTA keeps only HALF of the Cout tile's weights (B is split
// across the pair), loads its own halo patches and runs its own epilogue / stores.  Only the leader (cluster rank 0) issues
// MMAs; its patch_full / w_full / tmem_empty barriers collect arrivals from both CTAs, and tcgen05.commit multicasts the
// patch_empty / tmem_full arrivals to both.  Measured (tools/probe_pair.py): M256 x N64 x K16 costs 43 cycles against 48 for
// M128 x N64 on each SM alone (the B operand is fetched once per pair), and -- what matters more for the Cin = 128 layers
// -- halving the resident weights (147 -> 74 KB) makes room for a patch ring that can actually prefetch.
//
// Staging buffers rotate over `nstage` (2 or 3) tiles independent of the two lanes: with 3, the residual tile of tile i+3 is
// requested as soon as tile i's store has been read out, a full tile period earlier than with one buffer per lane (the
// residual load latency used to sit on the critical path of every `+res` layer: 95 us against 72 us without residual).
//
// MODE 1 is the same persistent, weight-stationary pipeline for 1x1 stride-1 convs over the flattened
// pixel dimension (tile = 128 consecutive pixels, one tap, up to two K-concatenated sources): these
// layers have K of only 64..512, so they are pure streaming and live or die by the per-tile overhead.
#include "conv_tc.cuh"
#include "ptx.cuh"
#include <cstdlib>
#include <cstring>

namespace pidnet {

namespace {

constexpr int kWsThreads = 384;
constexpr int kTH = 16, kTW = 8, kPH = 18, kPW = 10;
constexpr int kMaxPatch = 16;

__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int CK, int MODE>
struct WsGeom {
  static constexpr int kRowBytes = CK * 2;                          // one pixel of one chunk
  static constexpr int kPatchRows = MODE == 0 ? kPH * kPW : 128;    // halo patch | plain 128-pixel tile
  static constexpr int kPatchBytes = kPatchRows * kRowBytes;
  static constexpr int kPatchStride = (kPatchBytes + 1023) / 1024 * 1024;
  static constexpr int kTaps = MODE == 0 ? 9 : 1;
  static constexpr int kSboBytes = MODE == 0 ? kPW * kRowBytes : 8 * kRowBytes;
};

template <int BN, int CK, int MODE, bool PAIR>
__global__ void __launch_bounds__(kWsThreads, 1) conv3_ws_kernel(const __grid_constant__ Conv3Params p) {
  using G = WsGeom<CK, MODE>;
  constexpr int kWRows = PAIR ? BN / 2 : BN;   // weight rows (output channels) resident in THIS CTA
  constexpr int kSlabC = BN < 64 ? BN : 64;
  constexpr int kSlabRowBytes = kSlabC * 2;
  constexpr int kSlabBytes = 128 * kSlabRowBytes;
  constexpr int kNumSlabs = BN / kSlabC;
  constexpr int kStageBytes = 128 * BN * 2;
  constexpr int kWTileBytes = kWRows * CK * 2;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int chunks = p.chunks;
  const int npatch = p.npatch;
  const int nstage = p.nstage;   // staging buffers (2: one per lane, 3: rotating)
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const uint32_t w_base = smem_base;
  const uint32_t w_bytes = static_cast<uint32_t>(G::kTaps) * chunks * kWTileBytes;
  const uint32_t patch_base = w_base + w_bytes;
  const uint32_t stage_base = patch_base + npatch * G::kPatchStride;
  uint8_t* stage_gen = smem_gen + w_bytes + npatch * G::kPatchStride;
  const uint32_t bar_base = stage_base + nstage * kStageBytes;
  // barriers (8 B each): w_full | patch_full[kMaxPatch] | patch_empty[kMaxPatch] | tmem_full[4] | tmem_empty[4] |
  //                      res_full[3][2] | stage_free[3][2] | stage_ready[3] | tmem slot  -- 8 * (25 + 2 * kMaxPatch) <= 512 bytes
  // res_full / stage_free are indexed by (staging buffer, lane): each is waited on by ONE epilogue group, which therefore
  // observes every phase (a barrier shared by both groups would alias parities when the lanes drift apart).
  const uint32_t w_full = bar_base;
  auto patch_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto patch_empty = [&](int s) { return bar_base + 8u * (1 + kMaxPatch + s); };
  auto tmem_full = [&](int b) { return bar_base + 8u * (1 + 2 * kMaxPatch + b); };
  auto tmem_empty = [&](int b) { return bar_base + 8u * (5 + 2 * kMaxPatch + b); };
  auto res_full = [&](int sb, int g) { return bar_base + 8u * (9 + 2 * kMaxPatch + 2 * sb + g); };
  auto stage_free = [&](int sb, int g) { return bar_base + 8u * (15 + 2 * kMaxPatch + 2 * sb + g); };
  auto stage_ready = [&](int sb) { return bar_base + 8u * (21 + 2 * kMaxPatch + sb); };
  const uint32_t tmem_slot = bar_base + 8u * (24 + 2 * kMaxPatch);
  static_assert(8 * (25 + 2 * kMaxPatch) <= 512, "barrier block overlaps the bias table");
  float* bias_s = reinterpret_cast<float*>(stage_gen + nstage * kStageBytes + 512);   // BN floats (<= 512 B) after the barriers
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(
      stage_gen + nstage * kStageBytes + 8 * (24 + 2 * kMaxPatch));
  // staging-buffer rotation: tile iteration i uses buffer i % nstage; the (buffer, lane) pair recurs every 2 or 6 iterations
  auto stage_of = [&](int i) { return nstage == 3 ? i % 3 : (i & 1); };
  auto stage_use = [&](int i) { return static_cast<uint32_t>(nstage == 3 ? i / 6 : (i >> 1)); };
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tile = blockIdx.y;
  const int c_out0 = n_tile * BN;
  const int per_img = p.tiles_w * p.tiles_h;
  const int m_tiles = MODE == 0 ? p.N * per_img : p.tiles_w;   // MODE 1: tiles_w = ceil(rows / 128)
  // iteration i of this CTA works on tile blockIdx.x + i * gridDim.x; a pair iterates while its FIRST tile exists (the
  // second CTA of the last pair may run past the end: its loads are zero-filled and its stores clipped by TMA)
  auto tile_of = [&](int i) { return static_cast<long>(blockIdx.x) + static_cast<long>(i) * gridDim.x; };
  auto in_range = [&](int i) { return tile_of(i) - static_cast<long>(rank) < static_cast<long>(m_tiles); };
  // tile -> TMA coordinates (w, h, n) of its first output pixel
  auto tile_coord = [&](long tile_l, int& w0, int& h0, int& n) {
    const int tile = static_cast<int>(tile_l);
    if (MODE == 0) {
      n = tile / per_img;
      const int rem = tile - n * per_img;
      const int th = rem / p.tiles_w;
      w0 = (rem - th * p.tiles_w) * kTW;
      h0 = th * kTH;
    } else {
      w0 = tile * 128; h0 = 0; n = 0;
    }
  };

  if (threadIdx.x == 0) pdl_launch_dependents();
  if (threadIdx.x == 0) {
    mbar_init(w_full, PAIR ? 2 : 1);
    for (int s = 0; s < kMaxPatch; ++s) {
      mbar_init(patch_full(s), PAIR ? 2 : 1);
      mbar_init(patch_empty(s), 1);
    }
    for (int b = 0; b < 4; ++b) {
      mbar_init(tmem_full(b), 1);
      mbar_init(tmem_empty(b), PAIR ? 8 : 4);
    }
    for (int sb = 0; sb < 3; ++sb) {
      for (int g = 0; g < 2; ++g) {
        mbar_init(res_full(sb, g), 1);
        mbar_init(stage_free(sb, g), 1);
      }
      mbar_init(stage_ready(sb), 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {   // four accumulators: two per tile-parity lane
    if (PAIR) tmem_alloc_pair<4 * BN>(tmem_slot);
    else tmem_alloc<4 * BN>(tmem_slot);
  }
  pdl_wait();   // everything above overlaps the previous kernel's tail; global memory is touched only from here on
  if (threadIdx.x >= 64 && threadIdx.x < 64 + BN) bias_s[threadIdx.x - 64] = p.bias[c_out0 + threadIdx.x - 64];
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync();   // both CTAs' barriers are initialised before any remote arrive / TMA completion
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot_gen;
  // barriers that collect arrivals from both CTAs live in the leader (cluster rank 0)
  auto leader = [&](uint32_t bar) { return PAIR ? mapa_u32(bar, 0) : bar; };

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (elect_one()) {
      tma_prefetch_desc(&p.tmA);
      tma_prefetch_desc(&p.tmW);
      if (PAIR) {
        const uint32_t wf = leader(w_full);
        mbar_arrive_expect_tx_cluster(wf, w_bytes);
        for (int t = 0; t < G::kTaps * chunks; ++t)
          tma_load_2d_pair(w_base + t * kWTileBytes, &p.tmW, wf, t * CK, c_out0 + static_cast<int>(rank) * kWRows);
      } else {
        mbar_arrive_expect_tx(w_full, w_bytes);
        for (int t = 0; t < G::kTaps * chunks; ++t)
          tma_load_2d(w_base + t * kWTileBytes, &p.tmW, w_full, t * CK, c_out0);
      }
      // The patch ring is partitioned per lane (even slots: even tiles, odd slots: odd tiles) so that every
      // waiter observes EVERY phase of the barriers it uses (a shared ring would alias mbarrier parities).
      const int np_lane = npatch >> 1;
      for (int i = 0; in_range(i); ++i) {
        int w0, h0, n;
        tile_coord(tile_of(i), w0, h0, n);
        const int ln = i & 1;
        int j = (i >> 1) * chunks;   // lane-local item index
        for (int c = 0; c < chunks; ++c, ++j) {
          const int slot = ln + 2 * (j % np_lane);
          const uint32_t ph = (j / np_lane) & 1;
          mbar_wait(patch_empty(slot), ph ^ 1);
          const uint32_t dst = patch_base + slot * G::kPatchStride;
          const CUtensorMap* map = &p.tmA;
          int c0 = c * CK, c1 = w0 - 1, c2 = h0 - 1, c3 = n;
          if (MODE == 1) {
            const bool first = c < p.chunks0;
            map = first ? &p.tmA : &p.tmA2;
            c0 = (first ? c : c - p.chunks0) * CK; c1 = w0; c2 = 0; c3 = 0;
          }
          if (PAIR) {
            const uint32_t pf = leader(patch_full(slot));
            mbar_arrive_expect_tx_cluster(pf, G::kPatchBytes);
            tma_load_4d_pair(dst, map, pf, c0, c1, c2, c3);
          } else {
            mbar_arrive_expect_tx(patch_full(slot), G::kPatchBytes);
            tma_load_4d(dst, map, patch_full(slot), c0, c1, c2, c3);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp == 10) {
    // ============================== MMA issuers (one per tile parity; leader CTA only in PAIR mode) ==============================
    // The whole warp runs the loop converged (waits included); one ELECTED lane issues tcgen05.mma / commit
    // (elect.sync lets ptxas keep descriptors in uniform registers without per-MMA uniformisation loops).
    if (!PAIR || rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(PAIR ? 256 : 128, BN);
      constexpr uint64_t kLayout = (CK == 64) ? 2ull : 4ull;  // SWIZZLE_128B / SWIZZLE_64B
      // descriptor high words are constant: SBO (A: tile rows are kPW patch pixels apart), version, layout
      constexpr uint64_t a_hi = (static_cast<uint64_t>(G::kSboBytes >> 4) << 32) | (1ull << 46) | (kLayout << 61) |
                                (1ull << 16);
      constexpr uint64_t b_hi = (static_cast<uint64_t>((8 * G::kRowBytes) >> 4) << 32) | (1ull << 46) |
                                (kLayout << 61) | (1ull << 16);
      if (PAIR) mbar_wait_cluster(w_full, 0); else mbar_wait(w_full, 0);
      tc_fence_after();
      const int ab = warp == 1 ? 0 : 1;   // this warp's tile parity (lane); each lane alternates two accumulators
      for (int i = ab; in_range(i); i += 2) {
        const uint32_t u = static_cast<uint32_t>(i >> 1);      // lane-local tile counter
        const int ai = ab + 2 * static_cast<int>(u & 1);       // accumulator index 0..3
        const uint32_t acc = tmem_acc + ai * BN;
        if (PAIR) mbar_wait_cluster(tmem_empty(ai), ((u >> 1) & 1) ^ 1); else mbar_wait(tmem_empty(ai), ((u >> 1) & 1) ^ 1);
        tc_fence_after();
        const int np_lane = npatch >> 1;
        int j = (i >> 1) * chunks;   // lane-local item index (see the producer)
        for (int c = 0; c < chunks; ++c, ++j) {
          const int slot = ab + 2 * (j % np_lane);
          const uint32_t ph = (j / np_lane) & 1;
          if (PAIR) mbar_wait_cluster(patch_full(slot), ph); else mbar_wait(patch_full(slot), ph);
          tc_fence_after();
          const uint32_t pbase = patch_base + slot * G::kPatchStride;
          const uint32_t wbase = w_base + c * kWTileBytes;
          if (elect_one()) {
#pragma unroll
            for (int tap = 0; tap < G::kTaps; ++tap) {
              const int r = tap / 3, s = tap % 3;
              const uint64_t a_desc =
                  a_hi | static_cast<uint64_t>(((pbase + (r * kPW + s) * G::kRowBytes) & 0x3FFFF) >> 4);
              const uint64_t b_desc =
                  b_hi | static_cast<uint64_t>(((wbase + tap * chunks * kWTileBytes) & 0x3FFFF) >> 4);
#pragma unroll
              for (int k = 0; k < CK / 16; ++k) {
                if (PAIR) umma_bf16_pair(acc, a_desc + 2 * k, b_desc + 2 * k, idesc, (c | tap | k) != 0 ? 1u : 0u);
                else umma_bf16(acc, a_desc + 2 * k, b_desc + 2 * k, idesc, (c | tap | k) != 0 ? 1u : 0u);
              }
            }
            if (PAIR) {
              umma_commit_pair(patch_empty(slot), 3);
              if (c == chunks - 1) umma_commit_pair(tmem_full(ai), 3);
            } else {
              umma_commit(patch_empty(slot));
              if (c == chunks - 1) umma_commit(tmem_full(ai));
            }
          }
          __syncwarp();
        }
      }
    }
    __syncwarp();
  } else if (warp == 11 && p.stats != nullptr) {
    // ============================== TMA store warp + BatchNorm statistics (training forward) ==============================
    // Same protocol as the plain store warp below (lane 0 issues the stores and owns the bulk-async groups), but while the TMA
    // engine reads the staged tile out, the WHOLE warp reads it too and accumulates, per output channel, sum x and sum x^2 of
    // the bf16 values exactly as they are stored: lane l owns the 32-bit word l of a 128-byte row (two channels; 64-byte rows:
    // two rows per trip), which is bank-conflict free under any swizzle.  Rows outside the image (ragged tiles, the extra tile
    // of a last pair) are clipped by the TMA store and skipped here.  One fp64 atomic per channel and CTA at the end: the
    // BatchNorm that follows needs no statistics pass and no grid barrier.
    if (p.out_mode == kOutNHWCbf16) {
      // 16-byte loads: lane = (row within a group of kRowsPerLd rows, logical 8-channel chunk j); the physical chunk is
      // j ^ swizzle(row), so a quarter-warp always reads 128 contiguous bytes (no bank conflicts) and a lane always sees the
      // same eight channels.  (32-bit loads, one row per instruction, cost ~28 cycles each next to the saturated MMA operand
      // fetch of the C = 64 layers: +19 us on a 24 us conv; this form issues a quarter of the requests.)
      constexpr int kChunks = kSlabRowBytes / 16;        // 8 (128-byte rows) or 4 (64-byte rows)
      constexpr int kRowsPerLd = 32 / kChunks;           // 4 or 8 rows per warp-wide load
      constexpr int kBatch = kNumSlabs == 1 ? 8 : 4;     // loads in flight per slab
      const int j = lane % kChunks, rq = lane / kChunks;
      float s1[kNumSlabs][8], s2[kNumSlabs][8];
#pragma unroll
      for (int sl = 0; sl < kNumSlabs; ++sl)
#pragma unroll
        for (int e = 0; e < 8; ++e) s1[sl][e] = s2[sl][e] = 0.f;
      const long total_rows = static_cast<long>(p.N) * p.Ho * p.Wo;
      for (int i = 0; in_range(i); ++i) {
        int w0, h0, n;
        tile_coord(tile_of(i), w0, h0, n);
        const int sb = stage_of(i);
        mbar_wait(stage_ready(sb), static_cast<uint32_t>(i / nstage) & 1);
        if (lane == 0) {
          for (int sl = 0; sl < kNumSlabs; ++sl)
            if (c_out0 + sl * kSlabC < p.Cout)
              tma_store_4d(&p.tmD, stage_base + sb * kStageBytes + sl * kSlabBytes, c_out0 + sl * kSlabC, w0, h0, n);
          tma_store_commit();
        }
        __syncwarp();
        if (tile_of(i) < static_cast<long>(m_tiles)) {
          // valid rows of the tile: MODE 0 rows are (h0 + r / 8, w0 + r % 8); MODE 1 rows are consecutive pixels
          int vrows = 128, vcols = kTW;
          if (MODE == 0) {
            vrows = min(kTH, p.Ho - h0) * kTW;
            vcols = min(kTW, p.Wo - w0);
          } else {
            vrows = static_cast<int>(min(128L, total_rows - static_cast<long>(w0)));
          }
          const uint8_t* stage = stage_gen + sb * kStageBytes;
          constexpr uint32_t kOnes = 0x3F803F80u;   // bf16 (1.0, 1.0)
          for (int r0 = 0; r0 < vrows; r0 += kBatch * kRowsPerLd) {
            uint4 v[kBatch][kNumSlabs];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
              const int r = r0 + u * kRowsPerLd + rq;   // <= 127: inside the staging buffer whatever vrows is
              const uint32_t swz_r = (kSlabRowBytes == 128) ? (r & 7) : ((r >> 1) & 3);
              const uint32_t off = static_cast<uint32_t>(r) * kSlabRowBytes + ((static_cast<uint32_t>(j) ^ swz_r) << 4);
#pragma unroll
              for (int sl = 0; sl < kNumSlabs; ++sl) v[u][sl] = *reinterpret_cast<const uint4*>(stage + sl * kSlabBytes + off);
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
              const int r = r0 + u * kRowsPerLd + rq;
              const bool ok = r < vrows && (MODE != 0 || (r & (kTW - 1)) < vcols);
#pragma unroll
              for (int sl = 0; sl < kNumSlabs; ++sl) {
                const uint32_t x[4] = {ok ? v[u][sl].x : 0u, ok ? v[u][sl].y : 0u, ok ? v[u][sl].z : 0u, ok ? v[u][sl].w : 0u};
#pragma unroll
                for (int q2 = 0; q2 < 4; ++q2) {
                  s1[sl][2 * q2] = fma_bf16_ll(x[q2], kOnes, s1[sl][2 * q2]);
                  s2[sl][2 * q2] = fma_bf16_ll(x[q2], x[q2], s2[sl][2 * q2]);
                  s1[sl][2 * q2 + 1] = fma_bf16_hh(x[q2], kOnes, s1[sl][2 * q2 + 1]);
                  s2[sl][2 * q2 + 1] = fma_bf16_hh(x[q2], x[q2], s2[sl][2 * q2 + 1]);
                }
              }
            }
          }
        }
        __syncwarp();   // every lane is done with the buffer before lane 0 hands it on
        if (lane == 0) {
          tma_store_wait_read();
          mbar_arrive(stage_free(sb, (i + nstage) & 1));
        }
      }
      if (lane == 0) tma_store_wait_all();
      __syncwarp();
      double* tab = p.stats + static_cast<size_t>(blockIdx.x % kBnStatReplicas) * 2 * p.stats_C;
#pragma unroll
      for (int sl = 0; sl < kNumSlabs; ++sl) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float a = s1[sl][e], b2 = s2[sl][e];
#pragma unroll
          for (int o = kChunks; o < 32; o <<= 1) {   // the lanes that share chunk j hold different rows of the same channels
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b2 += __shfl_xor_sync(0xffffffffu, b2, o);
          }
          const int c = c_out0 + sl * kSlabC + j * 8 + e;
          if (rq == 0 && c < p.Cout) {
            atomicAdd(tab + c, static_cast<double>(a));
            atomicAdd(tab + p.stats_C + c, static_cast<double>(b2));
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 11) {
    // ============================== TMA store warp ==============================
    if (p.out_mode == kOutNHWCbf16 && elect_one()) {
      // This thread also prefetches the residual tiles: staging buffer sb is free exactly when this thread has seen tile i's
      // store read it out, so the residual of the buffer's NEXT tile (i + nstage) is requested right there (the TMA
      // producer never blocks on a staging buffer, which used to serialise its patch loads behind the epilogue).
      auto load_residual = [&](int i) {
        int w0, h0, n;
        tile_coord(tile_of(i), w0, h0, n);
        const int sb = stage_of(i);
        const uint32_t bar = res_full(sb, i & 1);
        mbar_arrive_expect_tx(bar, kStageBytes);
        for (int sl = 0; sl < kNumSlabs; ++sl)
          tma_load_4d(stage_base + sb * kStageBytes + sl * kSlabBytes, &p.tmR, bar, c_out0 + sl * kSlabC, w0, h0, n);
      };
      if (p.has_res) {
        tma_prefetch_desc(&p.tmR);
        for (int i = 0; i < nstage; ++i)
          if (in_range(i)) load_residual(i);
      }
      for (int i = 0; in_range(i); ++i) {
        int w0, h0, n;
        tile_coord(tile_of(i), w0, h0, n);
        const int sb = stage_of(i);
        mbar_wait(stage_ready(sb), static_cast<uint32_t>(i / nstage) & 1);
        for (int sl = 0; sl < kNumSlabs; ++sl)
          if (c_out0 + sl * kSlabC < p.Cout)
            tma_store_4d(&p.tmD, stage_base + sb * kStageBytes + sl * kSlabBytes, c_out0 + sl * kSlabC, w0, h0, n);
        tma_store_commit();
        tma_store_wait_read();       // smem of this buffer has been read out
        const int nxt = i + nstage;  // next user of the buffer (lane nxt & 1)
        if (p.has_res) { if (in_range(nxt)) load_residual(nxt); }
        else mbar_arrive(stage_free(sb, nxt & 1));
      }
      tma_store_wait_all();
    }
    __syncwarp();
  } else {
    // ============================== epilogue groups (warps 2..5: even tiles, 6..9: odd tiles) ==============================
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;        // tile row == TMEM lane
    const uint32_t swz = (kSlabRowBytes == 128) ? (row & 7) : (kSlabRowBytes == 64 ? ((row >> 1) & 3) : ((row >> 2) & 1));
    const int b = warp >= 6 ? 1 : 0;   // lane == tile parity
    for (int i = b; in_range(i); i += 2) {
      int w0, h0, n;
      tile_coord(tile_of(i), w0, h0, n);
      const uint32_t u = static_cast<uint32_t>(i >> 1);      // lane-local tile counter
      const int ai = b + 2 * static_cast<int>(u & 1);        // accumulator index 0..3
      mbar_wait(tmem_full(ai), (u >> 1) & 1);
      tc_fence_after();
      const uint32_t t_row = tmem_acc + ai * BN + (static_cast<uint32_t>(q * 32) << 16);
      if (p.out_mode == kOutNHWCbf16) {
        const int sb = stage_of(i);
        if (p.has_res) mbar_wait(res_full(sb, b), stage_use(i) & 1);
        else if (i >= nstage) mbar_wait(stage_free(sb, b), stage_use(i - nstage) & 1);
        uint8_t* stage = stage_gen + sb * kStageBytes;
#pragma unroll
        for (int g2 = 0; g2 < BN / 32; g2 += 2) {
          // two 32-column TMEM loads in flight, one wait
          uint32_t v[2][32];
          tmem_ld32(t_row + g2 * 32, v[0]);
          if (g2 + 1 < BN / 32) tmem_ld32(t_row + (g2 + 1) * 32, v[1]);
          tmem_ld_wait();
#pragma unroll
          for (int gg = 0; gg < 2; ++gg) {
            const int g = g2 + gg;
            if (g >= BN / 32) break;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int c = g * 32 + j * 8;
              const int slab = c / kSlabC;
              const int chunk = (c % kSlabC) / 8;
              uint4* ptr =
                  reinterpret_cast<uint4*>(stage + slab * kSlabBytes + row * kSlabRowBytes + ((chunk ^ swz) << 4));
              const float4 b0 = *reinterpret_cast<const float4*>(bias_s + c);
              const float4 b1 = *reinterpret_cast<const float4*>(bias_s + c + 4);
              float f[8];
              f[0] = __uint_as_float(v[gg][j * 8 + 0]) + b0.x; f[1] = __uint_as_float(v[gg][j * 8 + 1]) + b0.y;
              f[2] = __uint_as_float(v[gg][j * 8 + 2]) + b0.z; f[3] = __uint_as_float(v[gg][j * 8 + 3]) + b0.w;
              f[4] = __uint_as_float(v[gg][j * 8 + 4]) + b1.x; f[5] = __uint_as_float(v[gg][j * 8 + 5]) + b1.y;
              f[6] = __uint_as_float(v[gg][j * 8 + 6]) + b1.z; f[7] = __uint_as_float(v[gg][j * 8 + 7]) + b1.w;
              if (p.has_res) {
                const uint4 rr = *ptr;
                f[0] += bf16lo(rr.x); f[1] += bf16hi(rr.x); f[2] += bf16lo(rr.y); f[3] += bf16hi(rr.y);
                f[4] += bf16lo(rr.z); f[5] += bf16hi(rr.z); f[6] += bf16lo(rr.w); f[7] += bf16hi(rr.w);
              }
              if (p.relu) {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
              }
              uint4 o;
              o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]);
              o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
              *ptr = o;
            }
          }
        }
        // accumulator drained -> MMA warp may reuse it; staged tile complete -> store warp may ship it
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_cluster(leader(tmem_empty(ai))); else mbar_arrive(tmem_empty(ai));
          mbar_arrive(stage_ready(sb));
        }
      } else {
        int w, h;
        bool ok;
        const size_t plane = static_cast<size_t>(p.Ho) * p.Wo;
        if (MODE == 0) {
          w = w0 + row % kTW; h = h0 + row / kTW;
          ok = (w < p.Wo) && (h < p.Ho) && (n < p.N);
        } else {
          const long pix = static_cast<long>(w0) + row;          // flat pixel index over (n, h, w)
          ok = pix < static_cast<long>(p.N) * static_cast<long>(plane);
          n = static_cast<int>(pix / static_cast<long>(plane));
          const int rem = static_cast<int>(pix - static_cast<long>(n) * static_cast<long>(plane));
          h = rem / p.Wo; w = rem - h * p.Wo;
        }
        float* dst = p.out_f32 + (static_cast<size_t>(n) * p.Cout) * plane + static_cast<size_t>(h) * p.Wo + w;
#pragma unroll
        for (int g = 0; g < BN / 32; ++g) {
          uint32_t v[32];
          tmem_ld32(t_row + g * 32, v);
          tmem_ld_wait();
          if (ok) {
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              const int c = c_out0 + g * 32 + e;
              if (c < p.Cout) {
                float f = __uint_as_float(v[e]) + bias_s[g * 32 + e];
                if (p.relu) f = fmaxf(f, 0.f);
                dst[static_cast<size_t>(c) * plane] = f;
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (PAIR) mbar_arrive_cluster(leader(tmem_empty(ai))); else mbar_arrive(tmem_empty(ai)); }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync();   // no CTA leaves (or frees TMEM) while its peer may still signal it or read its smem
  if (warp == 1) {
    if (PAIR) tmem_dealloc_pair<4 * BN>(tmem_acc);
    else tmem_dealloc<4 * BN>(tmem_acc);
  }
}

// All instances are launched with programmatic stream serialization (see ptx.cuh); PDL can be disabled with PIDNET_PDL=0.
template <class Kernel>
cudaError_t ws_launch_ex(Kernel kernel, const Conv3Launch& L, int cluster, cudaStream_t stream) {
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = L.grid; cfg.blockDim = dim3(kWsThreads, 1, 1); cfg.dynamicSmemBytes = L.smem_bytes; cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (cluster > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = cluster; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kernel, L.p);
}
template <int BN, int CK>
cudaError_t ws_launch_inst(const Conv3Launch& L, cudaStream_t stream) {
  if (L.mode == 0) return ws_launch_ex(conv3_ws_kernel<BN, CK, 0, false>, L, 1, stream);
  return ws_launch_ex(conv3_ws_kernel<BN, CK, 1, false>, L, 1, stream);
}
// CTA-pair instance: launched as 2-CTA clusters along x
template <int BN, int CK>
cudaError_t ws_launch_pair(const Conv3Launch& L, cudaStream_t stream) {
  return ws_launch_ex(conv3_ws_kernel<BN, CK, 0, true>, L, 2, stream);
}
template <int BN, int CK>
cudaError_t ws_init_inst() {
  cudaError_t e = cudaFuncSetAttribute(conv3_ws_kernel<BN, CK, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kConv3MaxSmem);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(conv3_ws_kernel<BN, CK, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kConv3MaxSmem);
}

}  // namespace

cudaError_t conv3_ws_launch(const Conv3Launch& L, cudaStream_t stream) {
  const int BN = L.BN, CK = L.CK;
  if (L.pair) {
    if (L.mode == 0 && BN == 64 && CK == 64 && L.grid.x % 2 == 0) return ws_launch_pair<64, 64>(L, stream);
    return cudaErrorInvalidValue;
  }
  if (BN == 32 && CK == 32) return ws_launch_inst<32, 32>(L, stream);
  if (BN == 32 && CK == 64) return ws_launch_inst<32, 64>(L, stream);
  if (BN == 64 && CK == 32) return ws_launch_inst<64, 32>(L, stream);
  if (BN == 64 && CK == 64) return ws_launch_inst<64, 64>(L, stream);
  if (BN == 128 && CK == 32) return ws_launch_inst<128, 32>(L, stream);
  if (BN == 128 && CK == 64) return ws_launch_inst<128, 64>(L, stream);
  return cudaErrorInvalidValue;
}

cudaError_t conv3_ws_init() {
  cudaError_t e;
  if ((e = ws_init_inst<32, 32>()) != cudaSuccess) return e;
  if ((e = ws_init_inst<32, 64>()) != cudaSuccess) return e;
  if ((e = ws_init_inst<64, 32>()) != cudaSuccess) return e;
  if ((e = ws_init_inst<64, 64>()) != cudaSuccess) return e;
  if ((e = ws_init_inst<128, 32>()) != cudaSuccess) return e;
  if ((e = ws_init_inst<128, 64>()) != cudaSuccess) return e;
  return cudaFuncSetAttribute(conv3_ws_kernel<64, 64, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kConv3MaxSmem);
}

// The CTA-pair instance exists for (mode 0, BN 64, CK 64): the 3x3 C >= 64 layers, which are the MMA-bound ones.
bool conv3_ws_pair_available(int mode, int BN, int CK) { return mode == 0 && BN == 64 && CK == 64; }

// Shared-memory plan for (BN, CK, chunks): returns the number of patch buffers (0: does not fit), the staging-buffer
// count (3 when that costs no ring depth worth having, else 2) and the bytes.  pair: each CTA holds half of the weights.
int conv3_ws_plan(int mode, int BN, int CK, int chunks, int pair, int want_stages, int* nstage, size_t* smem_bytes) {
  const size_t w = static_cast<size_t>(mode == 0 ? 9 : 1) * chunks * (pair ? BN / 2 : BN) * CK * 2;
  const size_t patch = (static_cast<size_t>(mode == 0 ? kPH * kPW : 128) * CK * 2 + 1023) / 1024 * 1024;
  auto ring = [&](int stages, size_t* bytes) -> int {
    const size_t fixed = w + static_cast<size_t>(stages) * 128 * BN * 2 + 1024 /*barriers + bias*/ + 1024 /*alignment slack*/;
    if (fixed + 2 * patch > static_cast<size_t>(kConv3MaxSmem)) return 0;
    size_t np = (kConv3MaxSmem - fixed) / patch;
    if (np > static_cast<size_t>(kMaxPatch)) np = kMaxPatch;
    np &= ~static_cast<size_t>(1);   // even: the ring is split between the two tile-parity lanes
    *bytes = fixed + np * patch;
    return static_cast<int>(np);
  };
  size_t b2 = 0, b3 = 0;
  const int np2 = ring(2, &b2);
  const int np3 = want_stages >= 3 ? ring(3, &b3) : 0;
  const int need = 2 * (chunks < 2 ? 2 : chunks);   // per lane: a tile's chunks, at least two patches
  const bool use3 = np3 > 0 && (np3 >= np2 || np3 >= need);
  if (nstage) *nstage = use3 ? 3 : 2;
  if (smem_bytes) *smem_bytes = use3 ? b3 : b2;
  return use3 ? np3 : np2;
}

}  // namespace pidnet
