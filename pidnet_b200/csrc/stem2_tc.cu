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

constexpr int kS2Threads = 288;                 // 9 warps: 561 im2col rows in two rounds
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
  stem2_tc_kernel<C, U8><<<static_cast<unsigned>(blocks), kS2Threads, G::kSmem, st>>>(p);
  return cudaGetLastError();
}

}  // namespace

cudaError_t stem2_tc_launch(const Stem2Params& p, int C, int num_sms, cudaStream_t st) {
  // the fp32 path reads the image with 16-byte loads (W is a multiple of 8 by the engine's contract)
  if (!p.x_u8 && ((reinterpret_cast<uintptr_t>(p.x) & 15) != 0 || (p.W & 3) != 0)) return cudaErrorMisalignedAddress;
  if (C == 32) return p.x_u8 ? stem2_launch_inst<32, true>(p, num_sms, st) : stem2_launch_inst<32, false>(p, num_sms, st);
  if (C == 64) return p.x_u8 ? stem2_launch_inst<64, true>(p, num_sms, st) : stem2_launch_inst<64, false>(p, num_sms, st);
  return cudaErrorInvalidValue;
}

}  // namespace pidnet
