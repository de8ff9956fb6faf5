// Stem conv1.0 (3x3, stride 2, pad 1, Cin = 3, +bias, BN folded, ReLU) on tcgen05.
//
// K = 27 is far too small for a TMA-fed implicit GEMM (and the input is the caller's fp32 NCHW image),
// so the A operand is produced in software: thread t of the CTA gathers the 27 input values of output
// pixel t, rounds them to bf16 and writes row t of a 128 x 32 K-major SWIZZLE_64B tile (zero-padded to
// K = 32); two tcgen05.mma (K = 16 each) against the resident weight tile give the 128 x Cout fp32
// accumulator in TMEM; the epilogue adds the bias, applies ReLU, stages the bf16 NHWC tile in swizzled
// smem and ships it with one TMA store.  CUDA-core version: 1050 instructions per pixel (FFMA issue
// bound, 1.09 ms at batch 32); this one: ~150, i.e. HBM bound.
#include "kernels.cuh"
#include "ptx.cuh"

namespace pidnet {
namespace {

__device__ __forceinline__ uint32_t pk2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int BN, bool U8>  // Cout (32 or 64); U8: uint8 HWC BGR input normalised on load
__global__ void __launch_bounds__(128) stem_tc_kernel(const __grid_constant__ StemParams p) {
  constexpr int kRowB = BN * 2;              // output row bytes == swizzle span of the store map
  __shared__ __align__(1024) uint8_t a_s[128 * 64];       // A tile: 128 rows x 64 B, SWIZZLE_64B
  __shared__ __align__(1024) uint8_t b_s[BN * 64];        // weights: BN rows x 64 B, SWIZZLE_64B (pre-swizzled on host)
  __shared__ __align__(1024) uint8_t o_s[128 * kRowB];    // staged output tile
  __shared__ float bias_s[BN];
  __shared__ __align__(8) uint64_t bar_s;
  __shared__ uint32_t tmem_s;

  const int warp = threadIdx.x >> 5;
  const int row = threadIdx.x;
  for (int i = threadIdx.x; i < BN * 64 / 16; i += 128)
    reinterpret_cast<uint4*>(b_s)[i] = reinterpret_cast<const uint4*>(p.w_swz)[i];
  if (threadIdx.x < BN) bias_s[threadIdx.x] = p.bias[threadIdx.x];
  const uint32_t bar = smem_u32(&bar_s);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<BN>(smem_u32(&tmem_s));
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&tmem_s);
  const uint32_t a_addr = smem_u32(a_s), b_addr = smem_u32(b_s), o_addr = smem_u32(o_s);
  const long plane = static_cast<long>(p.H) * p.W;
  const int a_swz = (row >> 1) & 3;
  const uint32_t o_swz = (kRowB == 128) ? (row & 7) : ((row >> 1) & 3);
  uint32_t phase = 0;
  // batch statistics of the stored tile (training): warp w reads rows 32w..32w+31 of the staged tile, lane = 32-bit word of a row
  constexpr int kWords = kRowB / 4, kRowsPerTrip = 32 / kWords;
  const int lane = threadIdx.x & 31, wl = lane % kWords, rsub = lane / kWords;
  float st1[2] = {0.f, 0.f}, st2[2] = {0.f, 0.f};

  for (long tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
    // ---- A producer: im2col row of output pixel `pix`
    const long pix = tile * 128 + row;
    float v[28];
#pragma unroll
    for (int k = 0; k < 28; ++k) v[k] = 0.f;
    if (pix < p.rows) {
      const int ow = static_cast<int>(pix % p.Wo);
      const long t1 = pix / p.Wo;
      const int oh = static_cast<int>(t1 % p.Ho);
      const int n = static_cast<int>(t1 / p.Ho);
      const int ih0 = oh * 2 - 1, iw0 = ow * 2 - 1;
      if (U8) {
        const uint8_t* xn = p.x_u8 + static_cast<long>(n) * 3 * plane;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int ih = ih0 + r;
          const bool rok = ih >= 0 && ih < p.H;
          const uint8_t* rp = xn + static_cast<long>(rok ? ih : 0) * p.W * 3;
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            const int iw = iw0 + s;
            if (rok && iw >= 0 && iw < p.W) {
#pragma unroll
              for (int ci = 0; ci < 3; ++ci) {   // model channel ci (RGB) = byte 2 - ci of the BGR pixel
                // 256-entry table per channel, built on the host with numpy's exact arithmetic (engine.cu: u8_lut)
                v[(ci * 3 + r) * 3 + s] = __ldg(p.lut + ci * 256 + __ldg(rp + iw * 3 + (2 - ci)));
              }
            }
          }
        }
      }
      const float* xn = p.x + static_cast<long>(n) * 3 * plane;
#pragma unroll
      for (int ci = 0; ci < 3 && !U8; ++ci)
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int ih = ih0 + r;
          const bool rok = ih >= 0 && ih < p.H;
          const float* rp = xn + ci * plane + static_cast<long>(rok ? ih : 0) * p.W;
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            const int iw = iw0 + s;
            if (rok && iw >= 0 && iw < p.W) v[(ci * 3 + r) * 3 + s] = __ldg(rp + iw);
          }
        }
    }
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      uint4 q;
      q.x = pk2(v[ch * 8 + 0], v[ch * 8 + 1]);
      q.y = pk2(v[ch * 8 + 2], v[ch * 8 + 3]);
      q.z = ch < 3 ? pk2(v[ch * 8 + 4], v[ch * 8 + 5]) : pk2(0.f, 0.f);
      q.w = ch < 3 ? pk2(v[ch * 8 + 6], v[ch * 8 + 7]) : pk2(0.f, 0.f);
      if (ch == 3) { q.x = pk2(v[24], v[25]); q.y = pk2(v[26], 0.f); }
      *reinterpret_cast<uint4*>(a_s + row * 64 + ((ch ^ a_swz) << 4)) = q;
    }
    fence_proxy_async_smem();
    __syncthreads();
    // ---- two MMAs (K = 32)
    if (warp == 0) {
      if (elect_one()) {
        tc_fence_after();
        constexpr uint32_t idesc = make_idesc_bf16(128, BN);
        const uint64_t ad = make_kmajor_desc(a_addr, 64), bd = make_kmajor_desc(b_addr, 64);
        umma_bf16(tmem, ad, bd, idesc, 0u);
        umma_bf16(tmem, ad + 2, bd + 2, idesc, 1u);
        umma_commit(bar);
      }
      __syncwarp();
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- epilogue
    const uint32_t t_row = tmem + (static_cast<uint32_t>(warp * 32) << 16);
#pragma unroll
    for (int g = 0; g < BN / 32; ++g) {
      uint32_t acc[32];
      tmem_ld32(t_row + g * 32, acc);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = g * 32 + j * 8;
        uint4 o;
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          f[e] = __uint_as_float(acc[j * 8 + e]) + bias_s[c + e];
          if (p.relu) f[e] = fmaxf(f[e], 0.f);
        }
        o.x = pk2(f[0], f[1]); o.y = pk2(f[2], f[3]); o.z = pk2(f[4], f[5]); o.w = pk2(f[6], f[7]);
        *reinterpret_cast<uint4*>(o_s + row * kRowB + (((c >> 3) ^ o_swz) << 4)) = o;
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
      tma_store_4d(&p.tmD, o_addr, 0, static_cast<int>(tile * 128), 0, 0);
      tma_store_commit();
    }
    if (p.stats) {   // while the TMA engine reads the tile out (rows past the end of the tensor are clipped by the store, skipped here)
      const int vrows = static_cast<int>(min(128L, p.rows - tile * 128));
      constexpr uint32_t kOnes = 0x3F803F80u;   // bf16 (1.0, 1.0)
      for (int t0 = 0; t0 < 32; t0 += 8 * kRowsPerTrip) {   // eight loads in flight per trip
        uint32_t v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int r = warp * 32 + t0 + u * kRowsPerTrip + rsub;
          const uint32_t swz_r = (kRowB == 128) ? (r & 7) : ((r >> 1) & 3);
          v[u] = *reinterpret_cast<const uint32_t*>(o_s + r * kRowB + (((static_cast<uint32_t>(wl) >> 2) ^ swz_r) << 4) + (wl & 3) * 4);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int r = warp * 32 + t0 + u * kRowsPerTrip + rsub;
          const uint32_t x = r < vrows ? v[u] : 0u;
          st1[0] = fma_bf16_ll(x, kOnes, st1[0]); st2[0] = fma_bf16_ll(x, x, st2[0]);
          st1[1] = fma_bf16_hh(x, kOnes, st1[1]); st2[1] = fma_bf16_hh(x, x, st2[1]);
        }
      }
    }
    if (threadIdx.x == 0) tma_store_wait_read();   // o_s (and, by program order, a_s) may be overwritten by the next tile
    __syncthreads();
  }
  if (p.stats) {   // combine the four warps in shared memory (the staging tile is free), then one fp64 atomic per channel and CTA
    float* red = reinterpret_cast<float*>(o_s);   // [4 warps][2 (sum, sum sq)][BN]
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      float a = st1[e], b2 = st2[e];
      if (kRowsPerTrip == 2) {
        a += __shfl_down_sync(0xffffffffu, a, 16);
        b2 += __shfl_down_sync(0xffffffffu, b2, 16);
      }
      if (rsub == 0) {
        red[(warp * 2 + 0) * BN + 2 * wl + e] = a;
        red[(warp * 2 + 1) * BN + 2 * wl + e] = b2;
      }
    }
    __syncthreads();
    if (threadIdx.x < 2 * BN) {
      const int which = threadIdx.x / BN, c = threadIdx.x % BN;
      const float t = red[(0 * 2 + which) * BN + c] + red[(1 * 2 + which) * BN + c] + red[(2 * 2 + which) * BN + c] +
                      red[(3 * 2 + which) * BN + c];
      atomicAdd(p.stats + static_cast<size_t>(blockIdx.x % kBnStatReplicas) * 2 * BN + which * BN + c, static_cast<double>(t));
    }
  }
  if (threadIdx.x == 0) tma_store_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<BN>(tmem);
}

__global__ void stem_pack_kernel(const float* __restrict__ w, uint16_t* __restrict__ wsw, int Cout) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cout * 32) return;
  const int co = i >> 5, k = i & 31;
  const int chunk = k >> 3, within = k & 7;
  const __nv_bfloat16 v = __float2bfloat16_rn(k < 27 ? w[co * 27 + k] : 0.f);
  wsw[co * 32 + ((chunk ^ ((co >> 1) & 3)) * 8) + within] = *reinterpret_cast<const uint16_t*>(&v);
}

}  // namespace

cudaError_t stem_pack_launch(const float* w, uint8_t* w_swz, int Cout, cudaStream_t st) {
  stem_pack_kernel<<<(Cout * 32 + 255) / 256, 256, 0, st>>>(w, reinterpret_cast<uint16_t*>(w_swz), Cout);
  return cudaGetLastError();
}

cudaError_t stem_tc_launch(const StemParams& p, int Cout, int num_sms, cudaStream_t st) {
  long blocks = p.tiles;
  const long cap = static_cast<long>(num_sms) * 8;
  if (blocks > cap) blocks = cap;
  if (p.x_u8) {
    if (Cout == 32) stem_tc_kernel<32, true><<<static_cast<unsigned>(blocks), 128, 0, st>>>(p);
    else if (Cout == 64) stem_tc_kernel<64, true><<<static_cast<unsigned>(blocks), 128, 0, st>>>(p);
    else return cudaErrorInvalidValue;
  } else if (Cout == 32) stem_tc_kernel<32, false><<<static_cast<unsigned>(blocks), 128, 0, st>>>(p);
  else if (Cout == 64) stem_tc_kernel<64, false><<<static_cast<unsigned>(blocks), 128, 0, st>>>(p);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

}  // namespace pidnet
