// Fused stem: conv1.0 (3x3 s2, 3 -> C) + BN + ReLU  ->  conv1.3 (3x3 s2, C -> C) + BN + ReLU in ONE kernel
// (models/pidnet.py:24-31).  Unfused, the conv1.0 output is the largest tensor of the network (N x H/2 x W/2 x C bf16:
// 1.07 GB at 32 x 1024 x 2048, C = 32) and is written once and read once; here it never leaves shared memory.
//
// One CTA produces a 16 x 8 tile of the conv1.3 output:
//   P0  the 67 x 36 x 3 input patch of the tile -> bf16 in smem: fp32 NCHW images with 16-byte loads issued ONE TILE AHEAD
//       into registers (or uint8 HWC BGR frames through the input_transform table);
//   P1  software im2col of the 33 x 17 conv1.0 positions the tile needs (561 rows x K = 27 -> 32, bf16, SWIZZLE_64B);
//   P2  five M128 x N=C x K32 tcgen05 MMAs against the resident conv1.0 weights -> five TMEM accumulators;
//   P3  epilogue A: ReLU, bf16 (the bias rides in the MMA: K columns 27/28 of the im2col rows are 1.0 and the weight tile
//       holds bias_hi / bias_lo there; rows outside the conv1.0 output image are all-zero == conv1.3's padding), written as FOUR
//       PARITY PLANES (row parity x column parity of the intermediate position) in swizzled K-major layout;
//   P4  conv1.3 as 9 taps x C/16 MMAs whose A operand is a shifted 16 x 8 window of one parity plane: tap (r, s) of a
//       stride-2 conv reads intermediate (2 oh + r, 2 ow + s) = plane (r & 1, s & 1) at (oh + (r == 2), ow + (s == 2)),
//       i.e. a dense window -- the same shifted-descriptor trick as conv3_ws.cu (8-row groups are one plane row apart:
//       SBO = plane pitch x row bytes; the hardware swizzle is a function of absolute smem address bits);
//   P5  epilogue B: + bias, ReLU, bf16, swizzled staging tile, one TMA store.
// Rows of the im2col tile are ordered plane by plane so that TMEM lane -> plane address is a few compares.
#include "kernels.cuh"
#include "ptx.cuh"

namespace pidnet {
namespace {

constexpr int kS2Threads = 288;                 // 9 warps: 561 im2col rows in two rounds (384 threads measured slower: 0.785 vs 0.72-0.75 ms)
constexpr int kPosH = 33, kPosW = 17, kPos = kPosH * kPosW;   // conv1.0 positions per tile
// plane (pj, pi): rows a = 0..16 (pj = 0) / 0..15 (pj = 1), pitch 9 (pi = 0) / 8 (pi = 1) pixels
constexpr int kPl0 = 0, kPl1 = 17 * 9, kPl2 = kPl1 + 17 * 8, kPl3 = kPl2 + 16 * 9;
static_assert(kPl3 + 16 * 8 == kPos, "plane sizes");
constexpr int kPatH = 67;                                   // input rows per tile (4 * 16 + 3)
constexpr int kLdItems = 3 * kPatH * 9;                     // float4 loads per tile: 3 ch x 67 rows x 9 groups of 4 columns
constexpr int kLdRounds = (kLdItems + kS2Threads - 1) / kS2Threads;

__device__ __forceinline__ uint32_t pk2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// {lo, hi} -> bf16x2 with ReLU folded into the conversion (one instruction for two channels)
__device__ __forceinline__ uint32_t pk2_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

template <int C>
struct S2Geom {
  static constexpr int kRowB = C * 2;
  static constexpr int kA1Bytes = 640 * 64;
  static constexpr int pl_bytes(int px) { return (px * kRowB + 1023) / 1024 * 1024; }
  static constexpr int kPlOff0 = 0;
  static constexpr int kPlOff1 = kPlOff0 + pl_bytes(17 * 9);
  static constexpr int kPlOff2 = kPlOff1 + pl_bytes(17 * 8);
  static constexpr int kPlOff3 = kPlOff2 + pl_bytes(16 * 9);
  static constexpr int kPlanesBytes = kPlOff3 + pl_bytes(16 * 8);
  static constexpr int kW2Bytes = 9 * C * kRowB;
  static constexpr int kW1Bytes = C * 64;
  static constexpr int kStageBytes = 128 * kRowB;
  static constexpr int kOffPlanes = kA1Bytes;
  static constexpr int kOffW2 = kOffPlanes + kPlanesBytes;
  static constexpr int kOffW1 = kOffW2 + kW2Bytes;
  static constexpr int kOffStage = kOffW1 + kW1Bytes;
  static constexpr int kOffMisc = kOffStage + kStageBytes;   // bias1 | bias2 | barriers | tmem slot
  static constexpr int kSmem = kOffMisc + 1024 + 1024;       // + alignment slack
  static constexpr int kTmemCols = 6 * C <= 256 ? 256 : 512;  // five conv1.0 accumulators + one conv1.3 accumulator
};

template <int C, bool U8>
__global__ void __launch_bounds__(kS2Threads, C == 32 ? 2 : 1) stem2_tc_kernel(const __grid_constant__ Stem2Params p) {
  using G = S2Geom<C>;
  constexpr int kRowB = G::kRowB;
  constexpr uint32_t kSwzMask = C == 32 ? 3u : 7u;   // SWIZZLE_64B / SWIZZLE_128B: chunk16 ^= (addr >> 7) & mask
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* a1_gen = gen;
  uint8_t* pl_gen = gen + G::kOffPlanes;
  uint8_t* st_gen = gen + G::kOffStage;
  float* bias1_s = reinterpret_cast<float*>(gen + G::kOffMisc);
  float* bias2_s = bias1_s + C;
  const uint32_t bar1 = base + G::kOffMisc + 512, bar2 = bar1 + 8, slot = bar1 + 16;
  volatile uint32_t* slot_gen = reinterpret_cast<volatile uint32_t*>(gen + G::kOffMisc + 512 + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) pdl_launch_dependents();
  pdl_wait();
  for (int i = threadIdx.x; i < G::kW2Bytes / 16; i += kS2Threads)
    reinterpret_cast<uint4*>(gen + G::kOffW2)[i] = reinterpret_cast<const uint4*>(p.w2_swz)[i];
  for (int i = threadIdx.x; i < G::kW1Bytes / 16; i += kS2Threads)
    reinterpret_cast<uint4*>(gen + G::kOffW1)[i] = reinterpret_cast<const uint4*>(p.w1_swz)[i];
  if (threadIdx.x < C) { bias1_s[threadIdx.x] = p.bias1[threadIdx.x]; bias2_s[threadIdx.x] = p.bias2[threadIdx.x]; }
  if (threadIdx.x == 0) {
    mbar_init(bar1, 1);
    mbar_init(bar2, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<G::kTmemCols>(slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_gen;
  const long plane_in = static_cast<long>(p.H) * p.W;
  const int per_img = p.tiles_w * p.tiles_h;
  const long tiles = static_cast<long>(p.N) * per_img;
  uint32_t phase = 0;
  // input patch of a tile: image rows 4 oh0 - 3 .. + 66, columns 4 ow0 - 4 .. + 35 (16-byte aligned groups), 3 channels.
  // It lives in the parity-plane region: dead once P1 has built the im2col tile, before epilogue A writes the planes.
  __nv_bfloat16* patch_gen = reinterpret_cast<__nv_bfloat16*>(pl_gen);
  float4 pre[kLdRounds];
  // (ci, row, column group) of this thread's k-th load never changes: packed once as ci << 11 | row << 4 | g (0xFFFF: none)
  uint32_t ldk[(kLdRounds + 1) / 2];
#pragma unroll
  for (int k = 0; k < kLdRounds; ++k) {
    const int idx = threadIdx.x + k * kS2Threads;
    const int cr = idx / 9, g = idx - cr * 9;        // cr = ci * 67 + row
    const int ci = cr / kPatH, row = cr - ci * kPatH;
    const uint32_t code = idx < kLdItems ? static_cast<uint32_t>(ci << 11 | row << 4 | g) : 0xFFFFu;
    if (k & 1) ldk[k >> 1] |= code << 16; else ldk[k >> 1] = code;
  }
  auto load_patch = [&](long t, float4 (&dst)[kLdRounds]) {
    const int tn = static_cast<int>(t / per_img);
    const int trem = static_cast<int>(t - static_cast<long>(tn) * per_img);
    const int tth = trem / p.tiles_w;
    const int ih_base = 4 * (tth * 16) - 3, iw_base = 4 * ((trem - tth * p.tiles_w) * 8) - 4;
    const float* xt = p.x + static_cast<long>(tn) * 3 * plane_in + static_cast<long>(ih_base) * p.W + iw_base;
    const int plane_i = static_cast<int>(plane_in);
    if (ih_base >= 0 && ih_base + kPatH <= p.H && iw_base >= 0 && iw_base + 36 <= p.W) {
      // interior tile (all but the image border): no per-element bounds checks
#pragma unroll
      for (int k = 0; k < kLdRounds; ++k) {
        const uint32_t code = (ldk[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu;
        const int off = static_cast<int>(code >> 11) * plane_i + static_cast<int>((code >> 4) & 127) * p.W + static_cast<int>(code & 15) * 4;
        dst[k] = (k < kLdRounds - 1 || code != 0xFFFFu) ? __ldg(reinterpret_cast<const float4*>(xt + off))
                                                        : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
#pragma unroll
      for (int k = 0; k < kLdRounds; ++k) {
        const uint32_t code = (ldk[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu;
        const int ci = code >> 11, row = (code >> 4) & 127, g = code & 15;
        const int ih = ih_base + row, iw = iw_base + 4 * g;
        dst[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (code != 0xFFFFu && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W)
          dst[k] = __ldg(reinterpret_cast<const float4*>(xt + ci * plane_i + row * p.W + 4 * g));
      }
    }
  };
  if (!U8 && static_cast<long>(blockIdx.x) < tiles) load_patch(blockIdx.x, pre);

  for (long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int n = static_cast<int>(tile / per_img);
    const int rem_t = static_cast<int>(tile - static_cast<long>(n) * per_img);
    const int th = rem_t / p.tiles_w;
    const int oh0 = th * 16, ow0 = (rem_t - th * p.tiles_w) * 8;
    const int y0 = 2 * oh0 - 1, x0 = 2 * ow0 - 1;   // conv1.0 position of (j, i) = (0, 0)

    // ---- P0: the tile's input patch (3 ch x 67 rows x 36 cols, bf16) -> smem.  fp32 images: the 16-byte loads were issued
    // one tile ahead (registers), so their latency hides behind the previous tile's MMAs and epilogues.
    if (U8) {
      const uint8_t* xn = p.x_u8 + static_cast<long>(n) * 3 * plane_in;
      for (int idx = threadIdx.x; idx < kPatH * 36; idx += kS2Threads) {
        const int row = idx / 36, col = idx - row * 36;
        const int ih = 4 * oh0 - 3 + row, iw = 4 * ow0 - 4 + col;
        const bool ok = ih >= 0 && ih < p.H && iw >= 0 && iw < p.W;
        const uint8_t* px = xn + (static_cast<long>(ok ? ih : 0) * p.W + (ok ? iw : 0)) * 3;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {   // model channel ci (RGB) = byte 2 - ci of the BGR pixel
          const float f = ok ? __ldg(p.lut + ci * 256 + __ldg(px + (2 - ci))) : 0.f;
          patch_gen[(ci * kPatH + row) * 36 + col] = __float2bfloat16_rn(f);
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < kLdRounds; ++k) {
        const int idx = threadIdx.x + k * kS2Threads;
        if (k < kLdRounds - 1 || idx < kLdItems) {   // (only the last round is partial)
          uint2 o;
          o.x = pk2(pre[k].x, pre[k].y); o.y = pk2(pre[k].z, pre[k].w);
          *reinterpret_cast<uint2*>(patch_gen + idx * 4) = o;   // idx == (ci * 67 + row) * 9 + g  ->  element (..) * 36 + 4 g
        }
      }
    }
    __syncthreads();
    if (!U8) {   // next tile's loads
      const long nt = tile + gridDim.x;
      if (nt < tiles) load_patch(nt, pre);
    }

    // ---- P1: im2col rows of the 33 x 17 conv1.0 positions, from the smem patch
    for (int idx = threadIdx.x; idx < kPos; idx += kS2Threads) {
      const int j = idx / kPosW, i = idx - j * kPosW;
      const int pi = i & 1, pj = j & 1;
      const int prow = (pj ? (pi ? kPl3 : kPl2) : (pi ? kPl1 : kPl0)) + (j >> 1) * (pi ? 8 : 9) + (i >> 1);
      const int swz = (prow >> 1) & 3;
      const int y = y0 + j, x = x0 + i;
      if (y < 0 || y >= p.H1 || x < 0 || x >= p.W1) {
        // outside the conv1.0 output == conv1.3's zero padding: an all-zero row (bias columns included) gives relu(0) = 0
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) *reinterpret_cast<uint4*>(a1_gen + prow * 64 + (ch << 4)) = make_uint4(0u, 0u, 0u, 0u);
        continue;
      }
      // The taps of (ci, r) are patch columns 2i+1 .. 2i+3 = the high half of 32-bit word i and both halves of word i+1, so
      // the K order is chosen to need no per-element shuffling: k = 2m, 2m+1: taps s = 1, 2 of m = ci*3 + r (word i+1 as
      // is); k = 18 + m: tap s = 0; k = 27, 28: 1.0 (the weight tile carries bias_hi / bias_lo there).
      const uint32_t* pw = reinterpret_cast<const uint32_t*>(patch_gen) + (2 * j) * 18 + i;
      uint32_t wa[9], w[16];
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          wa[ci * 3 + r] = pw[(ci * kPatH + r) * 18];
          w[ci * 3 + r] = pw[(ci * kPatH + r) * 18 + 1];
        }
#pragma unroll
      for (int m = 0; m < 4; ++m) w[9 + m] = __byte_perm(wa[2 * m], wa[2 * m + 1], 0x7632);
      w[13] = (wa[8] >> 16) | 0x3F800000u;
      w[14] = 0x00003F80u;
      w[15] = 0u;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch)
        *reinterpret_cast<uint4*>(a1_gen + prow * 64 + ((ch ^ swz) << 4)) = make_uint4(w[ch * 4], w[ch * 4 + 1], w[ch * 4 + 2], w[ch * 4 + 3]);
    }
    if (threadIdx.x == 0) tma_store_wait_read();   // the previous tile's store has read the staging buffer out
    fence_proxy_async_smem();
    __syncthreads();

    // ---- P2: conv1.0 = five M128 x C x K32 MMAs
    if (warp == 0) {
      if (elect_one()) {
        tc_fence_after();
        constexpr uint32_t idesc = make_idesc_bf16(128, C);
        const uint64_t bd = make_kmajor_desc(base + G::kOffW1, 64);
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          const uint64_t ad = make_kmajor_desc(base + q * 8192, 64);
          umma_bf16(tmem + q * C, ad, bd, idesc, 0u);
          umma_bf16(tmem + q * C, ad + 2, bd + 2, idesc, 1u);
        }
        umma_commit(bar1);
      }
      __syncwarp();
    }
    if (warp == 0) mbar_wait(bar1, phase);   // one warp polls; the others sleep in the barrier
    __syncthreads();
    tc_fence_after();

    // ---- P3: epilogue A -> parity planes (warp w reads TMEM lane quarter w & 3)
    {
      const int quarter = warp & 3;
      const int stepq = (kS2Threads / 32 - quarter + 3) / 4;   // warps sharing this TMEM lane quarter split its five accumulators
      for (int q = warp >> 2; q < 5; q += stepq) {
        const int prow = q * 128 + quarter * 32 + lane;
        const uint32_t t_addr = tmem + q * C + (static_cast<uint32_t>(quarter * 32) << 16);
        // plane row address: plane offset + (row - first row of the plane) * row bytes
        const int adj = prow >= kPl3 ? G::kPlOff3 - kPl3 * kRowB
                                     : (prow >= kPl2 ? G::kPlOff2 - kPl2 * kRowB : (prow >= kPl1 ? G::kPlOff1 - kPl1 * kRowB : G::kPlOff0));
        const uint32_t roff = static_cast<uint32_t>(G::kOffPlanes + adj + prow * kRowB);   // 1024-aligned base: offset bits == address bits
        const uint32_t swz = (roff >> 7) & kSwzMask;
#pragma unroll
        for (int g = 0; g < C / 32; ++g) {
          uint32_t acc[32];
          tmem_ld32(t_addr + g * 32, acc);
          tmem_ld_wait();
          if (prow < kPos) {
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const int c = g * 32 + jj * 8;
              const uint32_t* av = acc + jj * 8;   // bias came through the MMA; ReLU rides in the conversion
              uint4 o;
              o.x = pk2_relu(__uint_as_float(av[0]), __uint_as_float(av[1]));
              o.y = pk2_relu(__uint_as_float(av[2]), __uint_as_float(av[3]));
              o.z = pk2_relu(__uint_as_float(av[4]), __uint_as_float(av[5]));
              o.w = pk2_relu(__uint_as_float(av[6]), __uint_as_float(av[7]));
              *reinterpret_cast<uint4*>(gen + roff + (((c >> 3) ^ swz) << 4)) = o;
            }
          }
        }
      }
    }
    tc_fence_before();
    fence_proxy_async_smem();
    __syncthreads();

    // ---- P4: conv1.3 = 9 taps x C/16 MMAs on shifted windows of the parity planes
    if (warp == 0) {
      if (elect_one()) {
        tc_fence_after();
        constexpr uint32_t idesc = make_idesc_bf16(128, C);
        constexpr uint64_t kLayout = C == 32 ? 4ull : 2ull;   // SWIZZLE_64B / SWIZZLE_128B
        const uint32_t acc2 = tmem + 5 * C;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int r = tap / 3, s = tap % 3;
          const int pl = (r & 1) * 2 + (s & 1);
          const int pitch = (s & 1) ? 8 : 9;
          const int ploff = pl == 0 ? G::kPlOff0 : (pl == 1 ? G::kPlOff1 : (pl == 2 ? G::kPlOff2 : G::kPlOff3));
          const uint32_t a_start = base + G::kOffPlanes + ploff + ((r == 2 ? pitch : 0) + (s == 2 ? 1 : 0)) * kRowB;
          const uint64_t ad = static_cast<uint64_t>((a_start & 0x3FFFF) >> 4) | (1ull << 16) |
                              (static_cast<uint64_t>((pitch * kRowB) >> 4) << 32) | (1ull << 46) | (kLayout << 61);
          const uint64_t bd = make_kmajor_desc(base + G::kOffW2 + tap * C * kRowB, kRowB);
#pragma unroll
          for (int k = 0; k < C / 16; ++k) umma_bf16(acc2, ad + 2 * k, bd + 2 * k, idesc, (tap | k) != 0 ? 1u : 0u);
        }
        umma_commit(bar2);
      }
      __syncwarp();
    }
    if (warp == 0) mbar_wait(bar2, phase);
    __syncthreads();
    phase ^= 1;
    tc_fence_after();

    // ---- P5: epilogue B (warps 0..3) -> staging tile -> TMA store
    if (warp < 4) {
      const int row = warp * 32 + lane;
      const uint32_t t_addr = tmem + 5 * C + (static_cast<uint32_t>(warp * 32) << 16);
      const uint32_t swz = kRowB == 128 ? (row & 7) : ((row >> 1) & 3);
#pragma unroll
      for (int g = 0; g < C / 32; ++g) {
        uint32_t acc[32];
        tmem_ld32(t_addr + g * 32, acc);
        tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int c = g * 32 + jj * 8;
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(acc[jj * 8 + e]) + bias2_s[c + e];
          uint4 o;
          o.x = pk2_relu(f[0], f[1]); o.y = pk2_relu(f[2], f[3]); o.z = pk2_relu(f[4], f[5]); o.w = pk2_relu(f[6], f[7]);
          *reinterpret_cast<uint4*>(st_gen + row * kRowB + (((c >> 3) ^ swz) << 4)) = o;
        }
      }
    }
    tc_fence_before();
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
      tma_store_4d(&p.tmD, base + G::kOffStage, 0, ow0, oh0, n);
      tma_store_commit();
    }
  }
  if (threadIdx.x == 0) tma_store_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<G::kTmemCols>(tmem);
}

// ------------------------------------------------------------------------------------------------------------------------
// Pipelined form for C = 32 (one CTA per SM, three tiles in flight).  The lock-step kernel above keeps every warp of a CTA in
// the same phase, so the SM idles through each MMA round trip and barrier (ncu: 40 % issue activity, 29 % of the samples in
// barriers).  Here the phases are ROLES of different warps connected by mbarriers, and consecutive tiles occupy consecutive
// pipeline stages at the same time:
//   producer warps (12): tile i+2 -- image patch -> smem (16-byte loads issued a tile ahead), im2col -> A1
//   MMA warp       (1): conv1.0 of tile i+1 (A1 -> acc1[slot]) and conv1.3 of tile i (parity planes -> acc2[slot])
//   epilogue A     (8): tile i+1 -- acc1[slot] -> ReLU -> parity planes
//   epilogue B     (4): tile i   -- acc2[slot] -> bias, ReLU -> staging tile -> TMA store
// A1 and the planes are single buffers (the MMAs that read them are short); the TMEM accumulators are double-buffered
// (2 x 160 + 2 x 32 = 384 columns).
constexpr int kS3ProdWarps = 12, kS3Threads = (8 + 4 + 1 + kS3ProdWarps) * 32;   // 608
constexpr int kS3ProdThreads = kS3ProdWarps * 32;
constexpr int kS3LdRounds = (kLdItems + kS3ProdThreads - 1) / kS3ProdThreads;
constexpr int kS3Ring = 3;                                     // fp32 image patches in flight (TMA)
constexpr int kS3RingTx = 36 * kPatH * 3 * 4;                  // bytes of one box {36, 67, 3} fp32
constexpr int kS3RingBytes = (kS3RingTx + 1023) / 1024 * 1024;

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

template <bool U8>
__global__ void __launch_bounds__(kS3Threads, 1) stem3_tc_kernel(const __grid_constant__ Stem2Params p) {
  constexpr int C = 32, kRowB = 64;
  using G = S2Geom<32>;
  constexpr int kOffPatch = G::kOffPlanes + G::kPlanesBytes;                 // own buffer (the planes are live across tiles here)
  constexpr int kPatchBytes = (3 * kPatH * 36 * 2 + 1023) / 1024 * 1024;
  constexpr int kOffStage = kOffPatch + kPatchBytes;                         // two staging tiles
  constexpr int kOffW2 = kOffStage + 2 * G::kStageBytes;
  constexpr int kOffW1 = kOffW2 + G::kW2Bytes;
  constexpr int kOffMisc = kOffW1 + G::kW1Bytes;
  constexpr int kOffRing = kOffMisc + 1024;                                  // fp32 image patches as TMA delivers them
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* a1_gen = gen;
  float* bias2_s = reinterpret_cast<float*>(gen + kOffMisc);
  const uint32_t bars = base + kOffMisc + 512;
  auto ring_full = [&](int s) { return bars + 104u + 8u * s; };
  auto ring_empty = [&](int s) { return bars + 128u + 8u * s; };
  const uint32_t a1_full = bars, a1_free = bars + 8, pl_full = bars + 16, pl_free = bars + 24;
  auto acc1_full = [&](int s) { return bars + 32u + 8u * s; };
  auto acc1_free = [&](int s) { return bars + 48u + 8u * s; };
  auto acc2_full = [&](int s) { return bars + 64u + 8u * s; };
  auto acc2_free = [&](int s) { return bars + 80u + 8u * s; };
  const uint32_t slot = bars + 96;
  volatile uint32_t* slot_gen = reinterpret_cast<volatile uint32_t*>(gen + kOffMisc + 512 + 96);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) pdl_launch_dependents();
  pdl_wait();
  for (int i = threadIdx.x; i < G::kW2Bytes / 16; i += kS3Threads)
    reinterpret_cast<uint4*>(gen + kOffW2)[i] = reinterpret_cast<const uint4*>(p.w2_swz)[i];
  for (int i = threadIdx.x; i < G::kW1Bytes / 16; i += kS3Threads)
    reinterpret_cast<uint4*>(gen + kOffW1)[i] = reinterpret_cast<const uint4*>(p.w1_swz)[i];
  if (threadIdx.x < C) bias2_s[threadIdx.x] = p.bias2[threadIdx.x];
  if (threadIdx.x == 0) {
    mbar_init(a1_full, kS3ProdWarps);
    mbar_init(a1_free, 1);
    mbar_init(pl_full, 8);
    mbar_init(pl_free, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc1_full(s), 1);
      mbar_init(acc1_free(s), 8);
      mbar_init(acc2_full(s), 1);
      mbar_init(acc2_free(s), 4);
    }
    for (int s = 0; s < kS3Ring; ++s) {
      mbar_init(ring_full(s), 1);
      mbar_init(ring_empty(s), kS3ProdWarps);
    }
    fence_barrier_init();
  }
  if (warp == 12) tmem_alloc<512>(slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_gen;
  const long plane_in = static_cast<long>(p.H) * p.W;
  const int per_img = p.tiles_w * p.tiles_h;
  const long tiles = static_cast<long>(p.N) * per_img;
  const int n_it = static_cast<long>(blockIdx.x) < tiles ? static_cast<int>((tiles - 1 - blockIdx.x) / gridDim.x) + 1 : 0;
  auto tile_of = [&](int i) { return static_cast<long>(blockIdx.x) + static_cast<long>(i) * gridDim.x; };
  auto tile_origin = [&](long t, int& n, int& oh0, int& ow0) {
    n = static_cast<int>(t / per_img);
    const int rem_t = static_cast<int>(t - static_cast<long>(n) * per_img);
    const int th = rem_t / p.tiles_w;
    oh0 = th * 16; ow0 = (rem_t - th * p.tiles_w) * 8;
  };

  if (warp >= 13) {
    // ============================== producers: patch -> smem, im2col -> A1 ==============================
    const int ptid = threadIdx.x - 13 * 32;
    __nv_bfloat16* patch_gen = reinterpret_cast<__nv_bfloat16*>(gen + kOffPatch);
    // fp32 images: ONE TMA box load per tile ({36 columns, 67 rows, 3 channels} of the NCHW image, zero-filled outside
    // == the conv padding) into a ring of kS3Ring fp32 patches, issued kS3Ring - 1 tiles ahead: the lock-step kernel and
    // a register prefetch keep one 29 KB tile of loads in flight per SM, which caps it near 1.5 TB/s.
    auto issue_tma = [&](int i) {
      int tn, oh0, ow0;
      tile_origin(tile_of(i), tn, oh0, ow0);
      const int sl = i % kS3Ring;
      mbar_arrive_expect_tx(ring_full(sl), kS3RingTx);
      tma_load_4d(base + kOffRing + sl * kS3RingBytes, &p.tmX, ring_full(sl), 4 * ow0 - 4, 4 * oh0 - 3, 0, tn);
    };
    if (!U8 && ptid == 0) {
      tma_prefetch_desc(&p.tmX);
      for (int i = 0; i < kS3Ring - 1 && i < n_it; ++i) issue_tma(i);
    }
    for (int i = 0; i < n_it; ++i) {
      int n, oh0, ow0;
      tile_origin(tile_of(i), n, oh0, ow0);
      const int y0 = 2 * oh0 - 1, x0 = 2 * ow0 - 1;
      // P0: patch -> smem (bf16)
      if (U8) {
        const uint8_t* xn = p.x_u8 + static_cast<long>(n) * 3 * plane_in;
        for (int idx = ptid; idx < kPatH * 36; idx += kS3ProdThreads) {
          const int row = idx / 36, col = idx - row * 36;
          const int ih = 4 * oh0 - 3 + row, iw = 4 * ow0 - 4 + col;
          const bool ok = ih >= 0 && ih < p.H && iw >= 0 && iw < p.W;
          const uint8_t* px = xn + (static_cast<long>(ok ? ih : 0) * p.W + (ok ? iw : 0)) * 3;
#pragma unroll
          for (int ci = 0; ci < 3; ++ci) {   // model channel ci (RGB) = byte 2 - ci of the BGR pixel
            const float f = ok ? __ldg(p.lut + ci * 256 + __ldg(px + (2 - ci))) : 0.f;
            patch_gen[(ci * kPatH + row) * 36 + col] = __float2bfloat16_rn(f);
          }
        }
      } else {
        const int nx = i + kS3Ring - 1;   // refill the slot tile i-1 has just released
        if (ptid == 0 && nx < n_it) {
          if (nx >= kS3Ring) mbar_wait(ring_empty(nx % kS3Ring), static_cast<uint32_t>(nx / kS3Ring - 1) & 1);
          issue_tma(nx);
        }
        const int sl = i % kS3Ring;
        mbar_wait(ring_full(sl), static_cast<uint32_t>(i / kS3Ring) & 1);
        const float4* src = reinterpret_cast<const float4*>(gen + kOffRing + sl * kS3RingBytes);
#pragma unroll
        for (int k = 0; k < kS3LdRounds; ++k) {
          const int idx = ptid + k * kS3ProdThreads;
          if (k < kS3LdRounds - 1 || idx < kLdItems) {
            const float4 v = src[idx];
            uint2 o;
            o.x = pk2(v.x, v.y); o.y = pk2(v.z, v.w);
            *reinterpret_cast<uint2*>(patch_gen + idx * 4) = o;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(ring_empty(sl));
      }
      named_bar_sync(2, kS3ProdThreads);
      if (i >= 1) mbar_wait(a1_free, static_cast<uint32_t>(i - 1) & 1);   // conv1.0 of the previous tile has read A1
      // P1: im2col (same row order / K order as the lock-step kernel)
      for (int idx = ptid; idx < kPos; idx += kS3ProdThreads) {
        const int j = idx / kPosW, ii = idx - j * kPosW;
        const int pi = ii & 1, pj = j & 1;
        const int prow = (pj ? (pi ? kPl3 : kPl2) : (pi ? kPl1 : kPl0)) + (j >> 1) * (pi ? 8 : 9) + (ii >> 1);
        const int swz = (prow >> 1) & 3;
        const int y = y0 + j, x = x0 + ii;
        if (y < 0 || y >= p.H1 || x < 0 || x >= p.W1) {
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) *reinterpret_cast<uint4*>(a1_gen + prow * 64 + (ch << 4)) = make_uint4(0u, 0u, 0u, 0u);
          continue;
        }
        const uint32_t* pw = reinterpret_cast<const uint32_t*>(patch_gen) + (2 * j) * 18 + ii;
        uint32_t wa[9], w[16];
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            wa[ci * 3 + r] = pw[(ci * kPatH + r) * 18];
            w[ci * 3 + r] = pw[(ci * kPatH + r) * 18 + 1];
          }
#pragma unroll
        for (int m = 0; m < 4; ++m) w[9 + m] = __byte_perm(wa[2 * m], wa[2 * m + 1], 0x7632);
        w[13] = (wa[8] >> 16) | 0x3F800000u;
        w[14] = 0x00003F80u;
        w[15] = 0u;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
          *reinterpret_cast<uint4*>(a1_gen + prow * 64 + ((ch ^ swz) << 4)) = make_uint4(w[ch * 4], w[ch * 4 + 1], w[ch * 4 + 2], w[ch * 4 + 3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(a1_full);
      named_bar_sync(2, kS3ProdThreads);   // every producer is done reading the patch before the next tile overwrites it
    }
  } else if (warp == 12) {
    // ============================== MMA issuer ==============================
    constexpr uint32_t idesc = make_idesc_bf16(128, C);
    constexpr uint64_t kLayout = 4ull;   // SWIZZLE_64B
    // The warp serves two independent queues -- conv1.0 of tile n1 (needs A1 and a free acc1 slot) and conv1.3 of tile n2
    // (needs the planes of an already-converted tile and a free acc2 slot) -- and issues whichever is ready, the older
    // tile first: a fixed order would hold conv1.3 of tile i-1 back until the producers have finished tile i.
    int n1 = 0, n2 = 0;
    uint32_t spins = 0;
    while (n2 < n_it) {
      bool did = false;
      if (n2 < n1) {
        const int j = n2, sl = j & 1;
        bool ok = mbar_try_wait(pl_full, static_cast<uint32_t>(j) & 1);
        if (ok && j >= 2) ok = mbar_try_wait(acc2_free(sl), static_cast<uint32_t>((j >> 1) - 1) & 1);
        if (__all_sync(0xffffffffu, ok)) {
          tc_fence_after();
          if (elect_one()) {
            const uint32_t acc2 = tmem + 320 + sl * C;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const int r = tap / 3, sx = tap % 3;
              const int pl = (r & 1) * 2 + (sx & 1);
              const int pitch = (sx & 1) ? 8 : 9;
              const int ploff = pl == 0 ? G::kPlOff0 : (pl == 1 ? G::kPlOff1 : (pl == 2 ? G::kPlOff2 : G::kPlOff3));
              const uint32_t a_start = base + G::kOffPlanes + ploff + ((r == 2 ? pitch : 0) + (sx == 2 ? 1 : 0)) * kRowB;
              const uint64_t ad = static_cast<uint64_t>((a_start & 0x3FFFF) >> 4) | (1ull << 16) |
                                  (static_cast<uint64_t>((pitch * kRowB) >> 4) << 32) | (1ull << 46) | (kLayout << 61);
              const uint64_t bd = make_kmajor_desc(base + kOffW2 + tap * C * kRowB, kRowB);
#pragma unroll
              for (int k = 0; k < C / 16; ++k) umma_bf16(acc2, ad + 2 * k, bd + 2 * k, idesc, (tap | k) != 0 ? 1u : 0u);
            }
            umma_commit(pl_free);
            umma_commit(acc2_full(sl));
          }
          __syncwarp();
          ++n2;
          did = true;
        }
      }
      if (!did && n1 < n_it) {
        const int i = n1, sl = i & 1;
        bool ok = mbar_try_wait(a1_full, static_cast<uint32_t>(i) & 1);
        if (ok && i >= 2) ok = mbar_try_wait(acc1_free(sl), static_cast<uint32_t>((i >> 1) - 1) & 1);
        if (__all_sync(0xffffffffu, ok)) {
          tc_fence_after();
          if (elect_one()) {
            const uint64_t bd = make_kmajor_desc(base + kOffW1, 64);
#pragma unroll
            for (int q = 0; q < 5; ++q) {
              const uint64_t ad = make_kmajor_desc(base + q * 8192, 64);
              umma_bf16(tmem + sl * 160 + q * C, ad, bd, idesc, 0u);
              umma_bf16(tmem + sl * 160 + q * C, ad + 2, bd + 2, idesc, 1u);
            }
            umma_commit(a1_free);
            umma_commit(acc1_full(sl));
          }
          __syncwarp();
          ++n1;
          did = true;
        }
      }
      if (did) {
        spins = 0;
      } else if (++spins > (1u << 26)) {
        if (lane == 0) printf("pidnet_b200: stem3 MMA scheduler timed out (block %d n1 %d n2 %d of %d)\n", blockIdx.x, n1, n2, n_it);
        __trap();
      }
    }
  } else if (warp < 8) {
    // ============================== epilogue A: acc1 -> ReLU -> parity planes ==============================
    const int quarter = warp & 3;
    for (int i = 0; i < n_it; ++i) {
      const int s = i & 1;
      mbar_wait(acc1_full(s), static_cast<uint32_t>(i >> 1) & 1);
      if (i >= 1) mbar_wait(pl_free, static_cast<uint32_t>(i - 1) & 1);   // conv1.3 of the previous tile has read the planes
      tc_fence_after();
      for (int q = warp >> 2; q < 5; q += 2) {
        const int prow = q * 128 + quarter * 32 + lane;
        const uint32_t t_addr = tmem + s * 160 + q * C + (static_cast<uint32_t>(quarter * 32) << 16);
        const int adj = prow >= kPl3 ? G::kPlOff3 - kPl3 * kRowB
                                     : (prow >= kPl2 ? G::kPlOff2 - kPl2 * kRowB : (prow >= kPl1 ? G::kPlOff1 - kPl1 * kRowB : G::kPlOff0));
        const uint32_t roff = static_cast<uint32_t>(G::kOffPlanes + adj + prow * kRowB);
        const uint32_t swz = (roff >> 7) & 3u;
        uint32_t acc[32];
        tmem_ld32(t_addr, acc);
        tmem_ld_wait();
        if (prow < kPos) {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const uint32_t* av = acc + jj * 8;
            uint4 o;
            o.x = pk2_relu(__uint_as_float(av[0]), __uint_as_float(av[1]));
            o.y = pk2_relu(__uint_as_float(av[2]), __uint_as_float(av[3]));
            o.z = pk2_relu(__uint_as_float(av[4]), __uint_as_float(av[5]));
            o.w = pk2_relu(__uint_as_float(av[6]), __uint_as_float(av[7]));
            *reinterpret_cast<uint4*>(gen + roff + ((static_cast<uint32_t>(jj) ^ swz) << 4)) = o;
          }
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(acc1_free(s));
        mbar_arrive(pl_full);
      }
    }
  } else {
    // ============================== epilogue B (warps 8..11): acc2 -> bias, ReLU -> staging -> TMA store ==============================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t swz = (row >> 1) & 3;
    for (int i = 0; i < n_it; ++i) {
      const int s = i & 1;
      int n, oh0, ow0;
      tile_origin(tile_of(i), n, oh0, ow0);
      mbar_wait(acc2_full(s), static_cast<uint32_t>(i >> 1) & 1);
      tc_fence_after();
      uint32_t acc[32];
      tmem_ld32(tmem + 320 + s * C + (static_cast<uint32_t>(q * 32) << 16), acc);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc2_free(s));
      if (threadIdx.x == 8 * 32) tma_store_wait_read1();   // the store that used this staging buffer (tile i-2) has read it out
      named_bar_sync(1, 128);
      uint8_t* st = gen + kOffStage + s * G::kStageBytes;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(acc[jj * 8 + e]) + bias2_s[jj * 8 + e];
        uint4 o;
        o.x = pk2_relu(f[0], f[1]); o.y = pk2_relu(f[2], f[3]); o.z = pk2_relu(f[4], f[5]); o.w = pk2_relu(f[6], f[7]);
        *reinterpret_cast<uint4*>(st + row * kRowB + ((static_cast<uint32_t>(jj) ^ swz) << 4)) = o;
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (threadIdx.x == 8 * 32) {
        tma_store_4d(&p.tmD, base + kOffStage + s * G::kStageBytes, 0, ow0, oh0, n);
        tma_store_commit();
      }
    }
    if (threadIdx.x == 8 * 32) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) tmem_dealloc<512>(tmem);
}

template <bool U8>
cudaError_t stem3_launch_inst(const Stem2Params& p, int num_sms, cudaStream_t st) {
  using G = S2Geom<32>;
  constexpr int kSmem = G::kOffPlanes + G::kPlanesBytes + (3 * kPatH * 36 * 2 + 1023) / 1024 * 1024 + 2 * G::kStageBytes +
                        G::kW2Bytes + G::kW1Bytes + 1024 + kS3Ring * kS3RingBytes + 1024;
  static bool init = false;
  if (!init) {
    cudaError_t e = cudaFuncSetAttribute(stem3_tc_kernel<U8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return e;
    init = true;
  }
  long blocks = static_cast<long>(p.N) * p.tiles_w * p.tiles_h;
  if (blocks > num_sms) blocks = num_sms;
  return launch_pdl(stem3_tc_kernel<U8>, dim3(static_cast<unsigned>(blocks), 1, 1), dim3(kS3Threads, 1, 1), kSmem, st, p);
}

template <int C, bool U8>
cudaError_t stem2_launch_inst(const Stem2Params& p, int num_sms, cudaStream_t st) {
  using G = S2Geom<C>;
  static bool init = false;
  if (!init) {
    cudaError_t e = cudaFuncSetAttribute(stem2_tc_kernel<C, U8>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::kSmem);
    if (e != cudaSuccess) return e;
    init = true;
  }
  long blocks = static_cast<long>(p.N) * p.tiles_w * p.tiles_h;
  const long cap = static_cast<long>(num_sms) * (C == 32 ? 2 : 1);
  if (blocks > cap) blocks = cap;
  return launch_pdl(stem2_tc_kernel<C, U8>, dim3(static_cast<unsigned>(blocks), 1, 1), dim3(kS2Threads, 1, 1), G::kSmem, st, p);
}

}  // namespace

// pipelined: 1 = the warp-specialised three-tiles-in-flight kernel (C = 32 only), 0 = the lock-step kernel
cudaError_t stem2_tc_launch(const Stem2Params& p, int C, int num_sms, int pipelined, cudaStream_t st) {
  // the fp32 path reads the image with 16-byte loads (W is a multiple of 8 by the engine's contract)
  if (!p.x_u8 && ((reinterpret_cast<uintptr_t>(p.x) & 15) != 0 || (p.W & 3) != 0)) return cudaErrorMisalignedAddress;
  if (C == 32 && pipelined) return p.x_u8 ? stem3_launch_inst<true>(p, num_sms, st) : stem3_launch_inst<false>(p, num_sms, st);
  if (C == 32) return p.x_u8 ? stem2_launch_inst<32, true>(p, num_sms, st) : stem2_launch_inst<32, false>(p, num_sms, st);
  if (C == 64) return p.x_u8 ? stem2_launch_inst<64, true>(p, num_sms, st) : stem2_launch_inst<64, false>(p, num_sms, st);
  return cudaErrorInvalidValue;
}

}  // namespace pidnet
