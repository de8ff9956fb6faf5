// Training-mode kernels of the PIDNet path: BatchNorm with batch statistics (forward statistics, finalize +
// running-stat update, backward reduce/apply), device-side weight packing (fp32 master weights -> the bf16
// K-major layouts of the conv kernels, forward and dgrad orientation), layout converters and the transposes
// of the bilinear-upsample / average-pool operators.  Reference semantics: nn.BatchNorm2d(momentum=0.1,
// eps=1e-5) in train mode (models/model_utils.py:8-9), SURVEY.md Appendix H.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include "train_kernels.cuh"
#include "ptx.cuh"

namespace pidnet {
namespace {

struct F8 { float v[8]; };
__device__ __forceinline__ F8 ld8(const bf16* p) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  F8 r;
  r.v[0] = __uint_as_float(u.x << 16); r.v[1] = __uint_as_float(u.x & 0xFFFF0000u);
  r.v[2] = __uint_as_float(u.y << 16); r.v[3] = __uint_as_float(u.y & 0xFFFF0000u);
  r.v[4] = __uint_as_float(u.z << 16); r.v[5] = __uint_as_float(u.z & 0xFFFF0000u);
  r.v[6] = __uint_as_float(u.w << 16); r.v[7] = __uint_as_float(u.w & 0xFFFF0000u);
  return r;
}
__device__ __forceinline__ F8 unpack8(const uint4& u) {
  F8 r;
  r.v[0] = __uint_as_float(u.x << 16); r.v[1] = __uint_as_float(u.x & 0xFFFF0000u);
  r.v[2] = __uint_as_float(u.y << 16); r.v[3] = __uint_as_float(u.y & 0xFFFF0000u);
  r.v[4] = __uint_as_float(u.z << 16); r.v[5] = __uint_as_float(u.z & 0xFFFF0000u);
  r.v[6] = __uint_as_float(u.w << 16); r.v[7] = __uint_as_float(u.w & 0xFFFF0000u);
  return r;
}
__device__ __forceinline__ uint32_t pk(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void st8(bf16* p, const F8& f) {
  uint4 o;
  o.x = pk(f.v[0], f.v[1]); o.y = pk(f.v[2], f.v[3]); o.z = pk(f.v[4], f.v[5]); o.w = pk(f.v[6], f.v[7]);
  *reinterpret_cast<uint4*>(p) = o;
}
__device__ __forceinline__ uint4 pack8(const F8& f) {
  uint4 o;
  o.x = pk(f.v[0], f.v[1]); o.y = pk(f.v[2], f.v[3]); o.z = pk(f.v[4], f.v[5]); o.w = pk(f.v[6], f.v[7]);
  return o;
}
inline unsigned blocks_for(long total, int threads) { return static_cast<unsigned>((total + threads - 1) / threads); }

// ----------------------------------------------------------------------------- per-channel reductions
// grid.x = pixel chunks, block = 256 threads = (C/8 channel groups) x (pixel lanes); each thread strides over
// the pixels of its chunk four at a time (independent 16-byte loads in flight); fp32 partials per block, fp64
// atomics across blocks.  The chunk size adapts to the tensor (reduce_geometry) so that small maps still fill
// the machine.  `out` must be zero on entry; its consumers (finalize / coef / add_sums) clear it again.
// MODE 0: sum x, sum x^2 (forward statistics)
// MODE 1: sum dz', sum dz' * xhat with dz' = dz * [z > 0 if relu]   (BN backward)
template <int MODE>
__global__ void __launch_bounds__(256) chan_reduce_kernel(View x, View dz, View z, const float* __restrict__ mean,
                                                          const float* __restrict__ invstd, int relu, long pix_per_block,
                                                          double* __restrict__ out /*[2][C]*/) {
  extern __shared__ float red[];  // [lanes][groups*16]
  const int groups = x.C >> 3;
  const int lanes = blockDim.x / groups;
  const int cg = threadIdx.x % groups, ln = threadIdx.x / groups;
  const long npix = static_cast<long>(x.N) * x.H * x.W;
  const long p0 = static_cast<long>(blockIdx.x) * pix_per_block;
  const long p1 = min(p0 + pix_per_block, npix);
  float a[8], b[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) a[e] = b[e] = 0.f;
  float mu[8], is[8];
  if (MODE == 1) {
#pragma unroll
    for (int e = 0; e < 8; ++e) { mu[e] = mean[cg * 8 + e]; is[e] = invstd[cg * 8 + e]; }
  }
  constexpr int U = 4;
  for (long p = p0 + ln; p < p1; p += static_cast<long>(lanes) * U) {
    F8 xv[U], gv[U], zv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long q = p + static_cast<long>(u) * lanes;
      const bool ok = q < p1;
      const long qq = ok ? q : p;
      xv[u] = ld8(x.ptr + qq * x.ps + cg * 8);
      if (MODE == 1) {
        gv[u] = ld8(dz.ptr + qq * dz.ps + cg * 8);
        if (relu) zv[u] = ld8(z.ptr + qq * z.ps + cg * 8);
      }
      if (!ok) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { if (MODE == 0) xv[u].v[e] = 0.f; else gv[u].v[e] = 0.f; }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (MODE == 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { a[e] += xv[u].v[e]; b[e] += xv[u].v[e] * xv[u].v[e]; }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float g = gv[u].v[e];
          if (relu && !(zv[u].v[e] > 0.f)) g = 0.f;
          a[e] += g;
          b[e] += g * (xv[u].v[e] - mu[e]) * is[e];
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) { red[(ln * groups + cg) * 16 + e] = a[e]; red[(ln * groups + cg) * 16 + 8 + e] = b[e]; }
  __syncthreads();
  for (int i = threadIdx.x; i < groups * 16; i += blockDim.x) {
    float s = 0.f;
    for (int q = 0; q < lanes; ++q) s += red[q * groups * 16 + i];
    const int g = i / 16, e = i % 16;
    atomicAdd(out + (e >= 8 ? x.C : 0) + g * 8 + (e & 7), static_cast<double>(s));
  }
}

// ----------------------------------------------------------------------------- BN finalize (forward)
__global__ void bn_finalize_kernel(double* __restrict__ sums, int C, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, const float* __restrict__ conv_bias, float eps,
                                   float momentum, float* __restrict__ mean_out, float* __restrict__ invstd_out,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ run_mean,
                                   float* __restrict__ run_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = sums[c] / count;
  double var = sums[C + c] / count - m * m;
  sums[c] = 0.0; sums[C + c] = 0.0;   // leave the accumulator clear for the backward reduction
  if (var < 0) var = 0;
  const float is = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  const float g = gamma[c];
  mean_out[c] = static_cast<float>(m);
  invstd_out[c] = is;
  scale[c] = g * is;
  shift[c] = beta[c] - static_cast<float>(m) * g * is;
  if (run_mean) {
    // the conv bias (stem convs) cancels in the normalisation but is part of the batch mean the reference tracks
    const float bm = static_cast<float>(m) + (conv_bias ? conv_bias[c] : 0.f);
    const double unbiased = count > 1 ? var * count / (count - 1.0) : var;
    run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * bm;
    run_var[c] = (1.f - momentum) * run_var[c] + momentum * static_cast<float>(unbiased);
  }
}

// ----------------------------------------------------------------------------- BN backward
//   dz' = dz * [z > 0]        (relu)
//   dx  = gamma*invstd * (dz' - sum(dz')/M - xhat * sum(dz' xhat)/M)  =  A*dz' + B*x + D  per channel    (+= if accumulate)
//   dres (+)= dz'             (residual added before the ReLU)
// coef kernel: per-channel A, B, D (written over the forward's scale/shift/mean scratch), dgamma/dbeta accumulation,
// and clears the fp64 accumulator for the next step.
__global__ void bn_bwd_coef_kernel(double* __restrict__ sums, int C, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ mean, const float* __restrict__ invstd,
                                   float* __restrict__ coef /*[3][C]*/, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double sb = sums[c], sg = sums[C + c];
  sums[c] = 0.0; sums[C + c] = 0.0;
  if (dgamma) { dbeta[c] += static_cast<float>(sb); dgamma[c] += static_cast<float>(sg); }
  const double is = invstd[c], A = static_cast<double>(gamma[c]) * is;
  const double B = -A * is * (sg / count);
  coef[c] = static_cast<float>(A);
  coef[C + c] = static_cast<float>(B);
  coef[2 * C + c] = static_cast<float>(-A * (sb / count) - B * static_cast<double>(mean[c]));
}

__device__ __forceinline__ void ldc8(const float* p, float* o) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(View x, View dz, View z, View dx, View dres,
                                                           const float* __restrict__ coef, int relu, int acc_dx,
                                                           int acc_dres) {
  const int groups = x.C >> 3;
  const long total = static_cast<long>(x.N) * x.H * x.W * groups;
  const long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = static_cast<int>(idx % groups);
  const long p = idx / groups;
  F8 g = ld8(dz.ptr + p * dz.ps + cg * 8);
  if (relu) {
    const F8 zv = ld8(z.ptr + p * z.ps + cg * 8);
#pragma unroll
    for (int e = 0; e < 8; ++e) if (!(zv.v[e] > 0.f)) g.v[e] = 0.f;
  }
  if (dres.ptr) {
    F8 r = g;
    if (acc_dres) {
      const F8 o = ld8(dres.ptr + p * dres.ps + cg * 8);
#pragma unroll
      for (int e = 0; e < 8; ++e) r.v[e] += o.v[e];
    }
    st8(dres.ptr + p * dres.ps + cg * 8, r);
  }
  if (dx.ptr) {
    const F8 xv = ld8(x.ptr + p * x.ps + cg * 8);
    float A[8], B[8], D[8];
    ldc8(coef + cg * 8, A); ldc8(coef + x.C + cg * 8, B); ldc8(coef + 2 * x.C + cg * 8, D);
    F8 o;
#pragma unroll
    for (int e = 0; e < 8; ++e) o.v[e] = fmaf(A[e], g.v[e], fmaf(B[e], xv.v[e], D[e]));
    if (acc_dx) {
      const F8 old = ld8(dx.ptr + p * dx.ps + cg * 8);
#pragma unroll
      for (int e = 0; e < 8; ++e) o.v[e] += old.v[e];
    }
    st8(dx.ptr + p * dx.ps + cg * 8, o);
  }
}

// ----------------------------------------------------------------------------- fused single-launch BatchNorm
// One persistent launch per BatchNorm (forward or backward): phase 1 streams the block's pixel range once,
// accumulating the per-channel statistics and parking the loaded vectors in shared memory; a grid barrier; phase 2
// derives the per-channel coefficients and applies them to the parked data (re-reading from L2/HBM only what did not
// fit).  Versus the three-kernel form this removes one full read pass (forward) / three (backward) and two launches.
// The grid is sized to be co-resident (2 blocks of 512 threads per SM); blocks of other streams that temporarily
// hold an SM always terminate on their own, so the barrier cannot deadlock; the spin is bounded and traps.
constexpr int kBnThreads = 384;   // 2 blocks x 384 threads x 64 regs leave register room for a co-resident wgrad CTA
constexpr int kBnBwdThreadsMin = 256;   // backward: 2 blocks x 256 threads (<= 96 registers) co-reside with a 192-thread wgrad CTA
constexpr int kBnReplicas = kBnStatReplicas, kBnArriveSlots = 16, kBnArriveStride = 64;   // (stride in unsigned: 256 B; the replica count is shared with the conv kernels that accumulate the forward statistics in their epilogue)
constexpr int kBnRedMin = kBnThreads * 4;                    // block reduction scratch: four rounds of 512 x 4 values
__host__ __device__ inline int bn_red_floats(int C) { return 3 * C > kBnRedMin ? (3 * C + 3) / 4 * 4 : kBnRedMin; }   // also the coefficient table (3 x C)
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Same-address L2 atomics serialise at ~27 cycles each (B300_MICROARCH: "LTS atomic-ALU serializes per-address"): with ~300
// blocks a single arrival counter costs ~4 us, and so does every per-channel accumulator.  Both are therefore replicated:
// block b arrives on counter b % kBnArriveSlots (256 B apart) and adds its partial sums to replica b % kBnReplicas; the
// waiters poll all counters (one thread each), the coefficient phase adds the replicas.  The accumulators and counters of a
// launch are never cleared by the kernel: every BatchNorm owns its slots per direction and the step zeroes the whole region
// with one memset (bn_fused_acc_bytes, TrainNet::bn_acc_*), which removed the second (departure) counter round trip.
__device__ __forceinline__ void grid_barrier(unsigned* arrive, unsigned nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(arrive + (blockIdx.x % kBnArriveSlots) * kBnArriveStride, 1u);
  }
  if (threadIdx.x < kBnArriveSlots) {
    const unsigned k = threadIdx.x;
    const unsigned expect = (nblocks + kBnArriveSlots - 1 - k) / kBnArriveSlots;   // blocks b < nblocks with b % slots == k
    const unsigned* slot = arrive + k * kBnArriveStride;
    unsigned spins = 0;
    while (ld_acquire_u32(slot) < expect) {
      __nanosleep(20);
      if (++spins > (1u << 25)) { printf("pidnet_b200: BatchNorm grid barrier timed out (block %d)\n", blockIdx.x); __trap(); }
    }
    __threadfence();
  }
  __syncthreads();
}
// block-level reduction of 16 per-thread partials over the pixel lanes, then fp64 atomics: a[8] -> sums[0..C),
// b[8] -> sums[C..2C)
__device__ __forceinline__ void block_channel_reduce(float* red, const float (&a)[8], const float (&b)[8], int groups, int lanes,
                                                     int cg, int ln, bool active, int C, double* sums) {
#pragma unroll
  for (int round = 0; round < 4; ++round) {   // a[0..3], a[4..7], b[0..3], b[4..7]
    __syncthreads();
    if (active) {
#pragma unroll
      for (int e = 0; e < 4; ++e) red[(ln * groups + cg) * 4 + e] = round < 2 ? a[(round & 1) * 4 + e] : b[(round & 1) * 4 + e];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < groups * 4; i += blockDim.x) {
      float s = 0.f;
      for (int q = 0; q < lanes; ++q) s += red[q * groups * 4 + i];
      const int c = (i >> 2) * 8 + (round & 1) * 4 + (i & 3);
      if (s != 0.f) atomicAdd(sums + (round >> 1) * C + c, static_cast<double>(s));
    }
  }
}

struct BnFwdParams {
  View x, res, z;
  const float *gamma, *beta, *conv_bias;
  float *run_mean, *run_var, *mean, *invstd;
  float *scale, *shift;   // the folded per-channel affine this launch applied (the backward derives the ReLU mask from it)
  double* sums;      // [kBnReplicas][2C], zero on entry (the step's memset)
  unsigned* sync;    // [kBnArriveSlots x kBnArriveStride], zero on entry
  double count;
  long pix_per_block;
  int relu, stage_iters;
  int pre;           // 1: `sums` already holds the statistics (accumulated by the producing conv's store warp, conv3_ws.cu):
                     // no statistics pass, no grid barrier, plain (non-cooperative) launch, stage_iters == 0
};
// (loops: each thread walks its own pixels -- q = p0 + ln + m * lanes -- with running pointers, full trips of 4 without
// per-element bounds / parking tests and a scalar tail; ncu on the first form: 142 instructions per 16-byte vector, half of them
// 64-bit index arithmetic and predicates, at 51 % issue-active)
__global__ void __launch_bounds__(kBnThreads, 2) bn_fwd_fused_kernel(BnFwdParams p) {
  extern __shared__ uint4 smem_v[];
  float* red = reinterpret_cast<float*>(smem_v);
  uint4* park = smem_v + bn_red_floats(p.x.C) / 4;   // [stage_iters][kBnThreads]
  const int C = p.x.C, groups = C >> 3, lanes = kBnThreads / groups;
  const int cg = threadIdx.x % groups, ln = threadIdx.x / groups;
  const bool active = ln < lanes;
  const long npix = static_cast<long>(p.x.N) * p.x.H * p.x.W;
  const long p0 = static_cast<long>(blockIdx.x) * p.pix_per_block, p1 = min(p0 + p.pix_per_block, npix);
  const int nmine = (active && p1 > p0 + ln) ? static_cast<int>((p1 - p0 - ln + lanes - 1) / lanes) : 0;   // my pixels
  const int nfull = nmine & ~3;
  const int stage = p.stage_iters;                   // multiple of 4
  const bf16* xp = p.x.ptr + (p0 + ln) * p.x.ps + cg * 8;
  const long xs = static_cast<long>(lanes) * p.x.ps;
  uint4* mypark = park + threadIdx.x;
  float a[8], b[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) a[e] = b[e] = 0.f;
  if (!p.pre) {
    const bf16* xq = xp;
    int m = 0;
    for (; m < nfull; m += 4, xq += 4 * xs) {
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = __ldg(reinterpret_cast<const uint4*>(xq + k * xs));
      if (m < stage) {
#pragma unroll
        for (int k = 0; k < 4; ++k) mypark[(m + k) * kBnThreads] = u[k];
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const F8 f = unpack8(u[k]);
#pragma unroll
        for (int e = 0; e < 8; ++e) { a[e] += f.v[e]; b[e] = fmaf(f.v[e], f.v[e], b[e]); }
      }
    }
    for (; m < nmine; ++m, xq += xs) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(xq));
      if (m < stage) mypark[m * kBnThreads] = u;
      const F8 f = unpack8(u);
#pragma unroll
      for (int e = 0; e < 8; ++e) { a[e] += f.v[e]; b[e] = fmaf(f.v[e], f.v[e], b[e]); }
    }
  }
  if (!p.pre) {
    block_channel_reduce(red, a, b, groups, lanes, cg, ln, active, C, p.sums + (blockIdx.x % kBnReplicas) * 2 * C);
    grid_barrier(p.sync, gridDim.x);
  }
  // per-channel coefficients once per block (fp64 only for mean / variance), shared through the scratch table
  {
    const double inv = 1.0 / p.count;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      double s1 = 0.0, s2 = 0.0;
#pragma unroll
      for (int r = 0; r < kBnReplicas; ++r) { s1 += __ldcg(p.sums + r * 2 * C + c); s2 += __ldcg(p.sums + r * 2 * C + C + c); }
      const double m = s1 * inv;
      double var = s2 * inv - m * m;
      if (var < 0) var = 0;
      const float is = 1.f / sqrtf(static_cast<float>(var) + 1e-5f);
      const float g = p.gamma[c];
      red[c] = g * is;
      red[C + c] = p.beta[c] - static_cast<float>(m) * g * is;
      if (blockIdx.x == 0) {
        p.mean[c] = static_cast<float>(m);
        p.invstd[c] = is;
        p.scale[c] = red[c];
        p.shift[c] = red[C + c];
        if (p.run_mean) {
          // the conv bias (stem convs) cancels in the normalisation but is part of the batch mean the reference tracks
          const float bm = static_cast<float>(m) + (p.conv_bias ? p.conv_bias[c] : 0.f);
          const double unbiased = p.count > 1 ? var * p.count / (p.count - 1.0) : var;
          p.run_mean[c] = 0.9f * p.run_mean[c] + 0.1f * bm;
          p.run_var[c] = 0.9f * p.run_var[c] + 0.1f * static_cast<float>(unbiased);
        }
      }
    }
  }
  __syncthreads();   // publishes the table
  if (nmine > 0) {
    float sc[8], sh[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { sc[e] = red[cg * 8 + e]; sh[e] = red[C + cg * 8 + e]; }
    const bool has_res = p.res.ptr != nullptr;
    const bf16* rp = has_res ? p.res.ptr + (p0 + ln) * p.res.ps + cg * 8 : nullptr;
    const long rs = static_cast<long>(lanes) * p.res.ps;
    bf16* zp = p.z.ptr + (p0 + ln) * p.z.ps + cg * 8;
    const long zs = static_cast<long>(lanes) * p.z.ps;
    const bool relu = p.relu != 0;
    auto apply = [&](const uint4& u, const uint4& r, bf16* dst) {
      F8 f = unpack8(u);
      const F8 rv = unpack8(r);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        f.v[e] = fmaf(f.v[e], sc[e], sh[e]) + rv.v[e];
        if (relu) f.v[e] = fmaxf(f.v[e], 0.f);
      }
      st8(dst, f);
    };
    const uint4 zero = make_uint4(0, 0, 0, 0);
    // four pixels per trip: all loads of the batch are issued before the first use (the parked vectors cover only the
    // first `stage` of a thread's pixels; the rest comes back from L2)
    const bf16* xq = xp;
    int m = 0;
    for (; m < nfull; m += 4, xq += 4 * xs, zp += 4 * zs) {
      uint4 u[4], r[4];
      if (m < stage) {
#pragma unroll
        for (int k = 0; k < 4; ++k) u[k] = mypark[(m + k) * kBnThreads];
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) u[k] = __ldg(reinterpret_cast<const uint4*>(xq + k * xs));
      }
      if (has_res) {
#pragma unroll
        for (int k = 0; k < 4; ++k) r[k] = __ldg(reinterpret_cast<const uint4*>(rp + k * rs));
        rp += 4 * rs;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) r[k] = zero;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) apply(u[k], r[k], zp + k * zs);
    }
    for (; m < nmine; ++m, xq += xs, zp += zs) {
      const uint4 u = m < stage ? mypark[m * kBnThreads] : __ldg(reinterpret_cast<const uint4*>(xq));
      uint4 r = zero;
      if (has_res) { r = __ldg(reinterpret_cast<const uint4*>(rp)); rp += rs; }
      apply(u, r, zp);
    }
  }
}

struct BnBwdParams {
  View x, dz, z, dx, dres;
  const float *mean, *invstd, *gamma;
  const float *scale, *shift;   // forward's folded affine (mask_x)
  float *dgamma, *dbeta;
  double* sums;
  unsigned* sync;
  double count;
  long pix_per_block;
  int relu, acc_dx, acc_dres, stage_iters;
  int mask_x;   // relu without a residual: z > 0  <=>  bf16(relu(scale * x + shift)) > 0, recomputed from x (already loaded)
                // exactly as the forward evaluated it -- the output tensor z is not read at all
};
// z = bf16(max(fma(x, sc, sh), 0)) > 0  <=>  fma(x, sc, sh) > 0 (same fma as bn_fwd_fused_kernel; a positive fp32 value only
// rounds to a zero bf16 below 2^-134, where either sub-gradient of the ReLU kink is valid): one fma + select per element, the
// same instruction count as unpacking and testing z -- these kernels are issue-bound, not bandwidth-bound
__device__ __forceinline__ void mask_from_x(F8& g, const F8& xv, const float (&sc)[8], const float (&sh)[8]) {
#pragma unroll
  for (int e = 0; e < 8; ++e) if (!(fmaf(xv.v[e], sc[e], sh[e]) > 0.f)) g.v[e] = 0.f;
}
template <bool HAS_DR, int NT>   // HAS_DR: a residual input receives the masked gradient too (residual blocks); NT: threads per block
__global__ void __launch_bounds__(NT == 256 ? 320 : NT, 2) bn_bwd_fused_kernel(BnBwdParams p) {
  extern __shared__ uint4 smem_v[];
  float* red = reinterpret_cast<float*>(smem_v);
  uint4* park = smem_v + bn_red_floats(p.x.C) / 4;   // [stage_iters][2][NT]: x, masked dz
  const int C = p.x.C, groups = C >> 3, lanes = NT / groups;
  const int cg = threadIdx.x % groups, ln = threadIdx.x / groups;
  const bool active = ln < lanes;
  const long npix = static_cast<long>(p.x.N) * p.x.H * p.x.W;
  const long p0 = static_cast<long>(blockIdx.x) * p.pix_per_block, p1 = min(p0 + p.pix_per_block, npix);
  const int nmine = (active && p1 > p0 + ln) ? static_cast<int>((p1 - p0 - ln + lanes - 1) / lanes) : 0;   // my pixels
  const int nfull = nmine & ~1;
  const int stage = p.stage_iters;                   // multiple of 2
  const long first = p0 + ln;
  const bf16* xp = p.x.ptr + first * p.x.ps + cg * 8;
  const bf16* gp = p.dz.ptr + first * p.dz.ps + cg * 8;
  const long xs = static_cast<long>(lanes) * p.x.ps, gs = static_cast<long>(lanes) * p.dz.ps;
  const bool relu_z = p.relu && !p.mask_x;   // the mask has to come from the stored output (residual blocks)
  const bool mask_x = p.mask_x != 0;
  const bf16* zp = relu_z ? p.z.ptr + first * p.z.ps + cg * 8 : nullptr;
  const long zs = static_cast<long>(lanes) * p.z.ps;
  uint4* mypark = park + threadIdx.x;
  float a[8], b[8], msc[8], msh[8];   // a = sum dz', b = sum dz' * x (centred and scaled after the grid barrier, in fp64)
#pragma unroll
  for (int e = 0; e < 8; ++e) a[e] = b[e] = 0.f;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  // masked gradient of one vector (ReLU backward), returned unpacked; `ug` is rewritten with the masked values
  auto masked = [&](const F8& xv, uint4& ug, const uint4& uz) {
    F8 g = unpack8(ug);
    if (mask_x) {
      mask_from_x(g, xv, msc, msh);
      ug = pack8(g);
    } else if (relu_z) {
      const F8 zv = unpack8(uz);
#pragma unroll
      for (int e = 0; e < 8; ++e) if (!(zv.v[e] > 0.f)) g.v[e] = 0.f;
      ug = pack8(g);
    }
    return g;
  };
  if (nmine > 0) {
    if (mask_x) {
#pragma unroll
      for (int e = 0; e < 8; ++e) { msc[e] = p.scale[cg * 8 + e]; msh[e] = p.shift[cg * 8 + e]; }
    }
    const bf16 *xq = xp, *gq = gp, *zq = zp;
    int m = 0;
    for (; m < nfull; m += 2, xq += 2 * xs, gq += 2 * gs) {
      uint4 ux[2], ug[2], uz[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        ux[k] = __ldg(reinterpret_cast<const uint4*>(xq + k * xs));
        ug[k] = __ldg(reinterpret_cast<const uint4*>(gq + k * gs));
        uz[k] = relu_z ? __ldg(reinterpret_cast<const uint4*>(zq + k * zs)) : zero;
      }
      if (relu_z) zq += 2 * zs;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const F8 xv = unpack8(ux[k]);
        const F8 g = masked(xv, ug[k], uz[k]);
        if (m < stage) {
          mypark[((m + k) * 2 + 0) * NT] = ux[k];
          mypark[((m + k) * 2 + 1) * NT] = ug[k];
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) { a[e] += g.v[e]; b[e] = fmaf(g.v[e], xv.v[e], b[e]); }
      }
    }
    for (; m < nmine; ++m, xq += xs, gq += gs) {
      const uint4 ux = __ldg(reinterpret_cast<const uint4*>(xq));
      uint4 ug = __ldg(reinterpret_cast<const uint4*>(gq));
      uint4 uz = zero;
      if (relu_z) { uz = __ldg(reinterpret_cast<const uint4*>(zq)); zq += zs; }
      const F8 xv = unpack8(ux);
      const F8 g = masked(xv, ug, uz);
      if (m < stage) {
        mypark[(m * 2 + 0) * NT] = ux;
        mypark[(m * 2 + 1) * NT] = ug;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) { a[e] += g.v[e]; b[e] = fmaf(g.v[e], xv.v[e], b[e]); }
    }
  }
  block_channel_reduce(red, a, b, groups, lanes, cg, ln, active, C, p.sums + (blockIdx.x % kBnReplicas) * 2 * C);
  grid_barrier(p.sync, gridDim.x);
  {
    const double inv = 1.0 / p.count;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      // sum dz' * xhat = invstd * (sum dz' * x - mean * sum dz')
      double sb = 0.0, sx = 0.0;
#pragma unroll
      for (int r = 0; r < kBnReplicas; ++r) { sb += __ldcg(p.sums + r * 2 * C + c); sx += __ldcg(p.sums + r * 2 * C + C + c); }
      const double sg = (sx - static_cast<double>(p.mean[c]) * sb) * static_cast<double>(p.invstd[c]);
      const float isf = p.invstd[c], A = p.gamma[c] * isf;
      const float B = -A * isf * static_cast<float>(sg * inv);
      red[c] = A;
      red[C + c] = B;
      red[2 * C + c] = -A * static_cast<float>(sb * inv) - B * p.mean[c];
      if (blockIdx.x == 0 && p.dgamma) {
        p.dbeta[c] += static_cast<float>(sb);
        p.dgamma[c] += static_cast<float>(sg);
      }
    }
  }
  __syncthreads();   // publishes the table
  if (nmine > 0 && (p.dx.ptr || p.dres.ptr)) {
    // (A = gamma * invstd is also the forward's scale: with mask_x the same registers serve the ReLU mask of unparked vectors)
    float cB[8], cD[8];
    float (&cA)[8] = msc;
#pragma unroll
    for (int e = 0; e < 8; ++e) { cA[e] = red[cg * 8 + e]; cB[e] = red[C + cg * 8 + e]; cD[e] = red[2 * C + cg * 8 + e]; }
    const bool has_dx = p.dx.ptr != nullptr;
    constexpr bool has_dr = HAS_DR;
    const bool acc_dx = has_dx && p.acc_dx, acc_dr = has_dr && p.acc_dres;
    bf16* dxp = has_dx ? p.dx.ptr + first * p.dx.ps + cg * 8 : nullptr;
    bf16* drp = has_dr ? p.dres.ptr + first * p.dres.ps + cg * 8 : nullptr;
    const long dxs = static_cast<long>(lanes) * p.dx.ps, drs = static_cast<long>(lanes) * p.dres.ps;
    auto emit = [&](const uint4& ux, const F8& g, const uint4& uo, const uint4& ur, bf16* dxq, bf16* drq) {
      if (has_dr) {
        const F8 old = unpack8(ur);
        F8 r;
#pragma unroll
        for (int e = 0; e < 8; ++e) r.v[e] = g.v[e] + old.v[e];
        st8(drq, r);
      }
      if (has_dx) {
        const F8 xv = unpack8(ux);
        const F8 old = unpack8(uo);
        F8 o;
#pragma unroll
        for (int e = 0; e < 8; ++e) o.v[e] = fmaf(cA[e], g.v[e], fmaf(cB[e], xv.v[e], cD[e])) + old.v[e];
        st8(dxq, o);
      }
    };
    const bf16 *xq = xp, *gq = gp, *zq = zp;
    int m = 0;
    for (; m < nfull; m += 2, xq += 2 * xs, gq += 2 * gs) {   // two pixels per trip, loads first
      uint4 ux[2], ug[2], uz[2], ur[2], uo[2];
      const bool parked = m < stage;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        if (parked) {
          ux[k] = mypark[((m + k) * 2 + 0) * NT];
          ug[k] = mypark[((m + k) * 2 + 1) * NT];
          uz[k] = zero;
        } else {
          ux[k] = __ldg(reinterpret_cast<const uint4*>(xq + k * xs));
          ug[k] = __ldg(reinterpret_cast<const uint4*>(gq + k * gs));
          uz[k] = relu_z ? __ldg(reinterpret_cast<const uint4*>(zq + k * zs)) : zero;
        }
        ur[k] = acc_dr ? *reinterpret_cast<const uint4*>(drp + k * drs) : zero;
        uo[k] = acc_dx ? *reinterpret_cast<const uint4*>(dxp + k * dxs) : zero;
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        F8 g;
        if (parked) g = unpack8(ug[k]);               // parked gradients are already masked
        else g = masked(unpack8(ux[k]), ug[k], uz[k]);
        emit(ux[k], g, uo[k], ur[k], has_dx ? dxp + k * dxs : nullptr, has_dr ? drp + k * drs : nullptr);
      }
      if (relu_z) zq += 2 * zs;
      if (has_dx) dxp += 2 * dxs;
      if (has_dr) drp += 2 * drs;
    }
    for (; m < nmine; ++m, xq += xs, gq += gs) {
      const bool parked = m < stage;
      uint4 ux, ug, uz = zero;
      if (parked) {
        ux = mypark[(m * 2 + 0) * NT];
        ug = mypark[(m * 2 + 1) * NT];
      } else {
        ux = __ldg(reinterpret_cast<const uint4*>(xq));
        ug = __ldg(reinterpret_cast<const uint4*>(gq));
        if (relu_z) uz = __ldg(reinterpret_cast<const uint4*>(zq));
      }
      const uint4 ur = acc_dr ? *reinterpret_cast<const uint4*>(drp) : zero;
      const uint4 uo = acc_dx ? *reinterpret_cast<const uint4*>(dxp) : zero;
      F8 g;
      if (parked) g = unpack8(ug);
      else g = masked(unpack8(ux), ug, uz);
      emit(ux, g, uo, ur, dxp, drp);
      if (relu_z) zq += zs;
      if (has_dx) dxp += dxs;
      if (has_dr) drp += drs;
    }
  }
}

// ----------------------------------------------------------------------------- weight packing
// dst[row][kofs + (tap_i*chunks + cc)*BK + j] (bf16); forward: row = co, channel = ci; dgrad: row = ci, channel = co
// and the tap is mirrored.  One thread per destination element of this source's K range.
__device__ __forceinline__ void pack_one(const PackJob& j, long idx) {
  const long per_row = static_cast<long>(j.ntaps) * j.chunks * j.BK;
  const long total = static_cast<long>(j.rows_pad) * per_row;
  if (idx >= total) return;
  const int row = static_cast<int>(idx / per_row);
  const long k = idx % per_row;
  const int ti = static_cast<int>(k / (static_cast<long>(j.chunks) * j.BK));
  const int ch = static_cast<int>(k % (static_cast<long>(j.chunks) * j.BK));
  float v = 0.f;
  const int nrow = j.dgrad ? j.Cin : j.Cout, nch = j.dgrad ? j.Cout : j.Cin;
  if (row < nrow && ch < nch) {
    int r = (j.taps[ti] >> 4) & 0xF, s = j.taps[ti] & 0xF;
    const int co = j.dgrad ? ch : row, ci = j.dgrad ? row : ch;
    v = j.src[((static_cast<long>(co) * j.Cin_total + j.ci_off + ci) * j.k + r) * j.k + s];
  }
  j.dst[static_cast<long>(row) * j.Ktot + j.kofs + k] = __float2bfloat16_rn(v);
}
__global__ void __launch_bounds__(256) pack_weights_kernel(PackJob j) {
  pack_one(j, static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x);
}
// all jobs of a step in one launch: block_start[i] = first block of job i (block_start[njobs] = grid size)
__global__ void __launch_bounds__(256) pack_all_kernel(const PackJob* __restrict__ jobs, const unsigned* __restrict__ block_start,
                                                       int njobs) {
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {   // last job whose first block is <= blockIdx.x
    const int mid = (lo + hi + 1) >> 1;
    if (block_start[mid] <= blockIdx.x) lo = mid; else hi = mid - 1;
  }
  pack_one(jobs[lo], static_cast<long>(blockIdx.x - block_start[lo]) * blockDim.x + threadIdx.x);
}

// ----------------------------------------------------------------------------- layout converters
// fp32 NCHW image -> bf16 NHWC with C padded to Cp (zeros)
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ x, int N, int C, int H, int W, View out) {
  const long total = static_cast<long>(N) * H * W;
  const long p = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= total) return;
  const long hw = static_cast<long>(H) * W;
  const int n = static_cast<int>(p / hw);
  const long r = p % hw;
  for (int c0 = 0; c0 < out.C; c0 += 8) {
    F8 f;
#pragma unroll
    for (int e = 0; e < 8; ++e) f.v[e] = (c0 + e < C) ? __ldg(x + (static_cast<long>(n) * C + c0 + e) * hw + r) : 0.f;
    st8(out.ptr + p * out.ps + c0, f);
  }
}

// ----------------------------------------------------------------------------- transposed operators
// dlow (+)= U^T dhi for bilinear align_corners=False (gather over the hi-res support of each low-res pixel)
__device__ __forceinline__ void lerp_af(int dst, int in, int out, int& i0, int& i1, float& l) {
  const float scale = static_cast<float>(in) / static_cast<float>(out);
  float src = scale * (static_cast<float>(dst) + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  i0 = static_cast<int>(src);
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  l = src - static_cast<float>(i0);
}
// One output element group (low-res pixel x 8 channels) is gathered by LANES cooperating lanes: lane j takes the hi-res rows
// y0 + j, y0 + j + LANES, ... of the window and the partial sums are combined with shuffles.  (One thread per output walked a
// window of up to (2*scale+3)^2 hi-res pixels serially -- the whole 16x16 map for the global-pool branch of PAPPM -- with a few
// hundred threads in flight: 0.1 ms for a few hundred KB.)
template <int LANES>
__global__ void __launch_bounds__(256) upsample_transpose_kernel(View dhi, View dlow, int accumulate) {
  const int groups = dlow.C >> 3;
  const long total = static_cast<long>(dlow.N) * dlow.H * dlow.W * groups;
  const long gid = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long idx_raw = gid / LANES;
  const int sub = static_cast<int>(gid % LANES);
  const bool live = idx_raw < total;
  const long idx = live ? idx_raw : total - 1;     // surplus lanes of the last warp keep the shuffles convergent
  const int cg = static_cast<int>(idx % groups);
  const long p = idx / groups;
  const int lx = static_cast<int>(p % dlow.W);
  const long t = p / dlow.W;
  const int ly = static_cast<int>(t % dlow.H);
  const int n = static_cast<int>(t / dlow.H);
  // hi-res rows/cols whose 2-tap footprint can include (ly, lx): conservative window from the inverse map
  const float sy = static_cast<float>(dhi.H) / dlow.H, sx = static_cast<float>(dhi.W) / dlow.W;
  const int y0 = max(0, static_cast<int>(floorf((ly - 1) * sy)) - 1), y1 = min(dhi.H - 1, static_cast<int>(ceilf((ly + 2) * sy)) + 1);
  const int x0 = max(0, static_cast<int>(floorf((lx - 1) * sx)) - 1), x1 = min(dhi.W - 1, static_cast<int>(ceilf((lx + 2) * sx)) + 1);
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  for (int y = y0 + sub; y <= y1; y += LANES) {
    int a0, a1; float la;
    lerp_af(y, dlow.H, dhi.H, a0, a1, la);
    float wy = 0.f;
    if (a0 == ly) wy += 1.f - la;
    if (a1 == ly) wy += la;
    if (wy == 0.f) continue;
    for (int x = x0; x <= x1; ++x) {
      int b0, b1; float lb;
      lerp_af(x, dlow.W, dhi.W, b0, b1, lb);
      float wx = 0.f;
      if (b0 == lx) wx += 1.f - lb;
      if (b1 == lx) wx += lb;
      if (wx == 0.f) continue;
      const F8 g = ld8(dhi.ptr + ((static_cast<long>(n) * dhi.H + y) * dhi.W + x) * dhi.ps + cg * 8);
      const float w = wy * wx;
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += w * g.v[e];
    }
  }
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) {
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], o);
  }
  if (!live || sub != 0) return;
  F8 o;
  bf16* dst = dlow.ptr + p * dlow.ps + cg * 8;
  if (accumulate) {
    const F8 old = ld8(dst);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += old.v[e];
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) o.v[e] = acc[e];
  st8(dst, o);
}

// dx (+)= P^T dy for AvgPool(k,s,p) with count_include_pad (k == 0: global average pool)
__global__ void __launch_bounds__(256) pool_transpose_kernel(View dy, View dx, int k, int stride, int pad, int accumulate) {
  const int groups = dx.C >> 3;
  const long total = static_cast<long>(dx.N) * dx.H * dx.W * groups;
  const long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = static_cast<int>(idx % groups);
  const long p = idx / groups;
  const int x = static_cast<int>(p % dx.W);
  const long t = p / dx.W;
  const int y = static_cast<int>(t % dx.H);
  const int n = static_cast<int>(t / dx.H);
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (k == 0) {
    const F8 g = ld8(dy.ptr + static_cast<long>(n) * dy.ps + cg * 8);
    const float inv = 1.f / static_cast<float>(dx.H * dx.W);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = g.v[e] * inv;
  } else {
    const float inv = 1.f / static_cast<float>(k * k);
    for (int oy = 0; oy < dy.H; ++oy) {
      const int h0 = oy * stride - pad;
      if (y < h0 || y >= h0 + k) continue;
      for (int ox = 0; ox < dy.W; ++ox) {
        const int w0 = ox * stride - pad;
        if (x < w0 || x >= w0 + k) continue;
        const F8 g = ld8(dy.ptr + ((static_cast<long>(n) * dy.H + oy) * dy.W + ox) * dy.ps + cg * 8);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += g.v[e] * inv;
      }
    }
  }
  bf16* dst = dx.ptr + p * dx.ps + cg * 8;
  if (accumulate) {
    const F8 old = ld8(dst);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += old.v[e];
  }
  F8 o;
#pragma unroll
  for (int e = 0; e < 8; ++e) o.v[e] = acc[e];
  st8(dst, o);
}

// generic elementwise: out (+)= a * [mask > 0]   (ReLU backward / plain copy / accumulate)
__global__ void __launch_bounds__(256) masked_add_kernel(View a, View mask, View out, int accumulate) {
  const int groups = out.C >> 3;
  const long total = static_cast<long>(out.N) * out.H * out.W * groups;
  const long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = static_cast<int>(idx % groups);
  const long p = idx / groups;
  F8 g = ld8(a.ptr + p * a.ps + cg * 8);
  if (mask.ptr) {
    const F8 m = ld8(mask.ptr + p * mask.ps + cg * 8);
#pragma unroll
    for (int e = 0; e < 8; ++e) if (!(m.v[e] > 0.f)) g.v[e] = 0.f;
  }
  bf16* dst = out.ptr + p * out.ps + cg * 8;
  if (accumulate) {
    const F8 old = ld8(dst);
#pragma unroll
    for (int e = 0; e < 8; ++e) g.v[e] += old.v[e];
  }
  st8(dst, g);
}


// bilinear sample (align_corners=False) of 8 channels of a low-res view at hi-res pixel (y, x)
__device__ __forceinline__ F8 sample8(const View& b, int n, int y, int x, int H, int W, int c) {
  int y0, y1, x0, x1; float ly, lx;
  lerp_af(y, b.H, H, y0, y1, ly);
  lerp_af(x, b.W, W, x0, x1, lx);
  const bf16* base = b.ptr + static_cast<long>(n) * b.H * b.W * b.ps + c;
  const F8 v00 = ld8(base + (static_cast<long>(y0) * b.W + x0) * b.ps), v01 = ld8(base + (static_cast<long>(y0) * b.W + x1) * b.ps);
  const F8 v10 = ld8(base + (static_cast<long>(y1) * b.W + x0) * b.ps), v11 = ld8(base + (static_cast<long>(y1) * b.W + x1) * b.ps);
  const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
  F8 r;
#pragma unroll
  for (int e = 0; e < 8; ++e) r.v[e] = w00 * v00.v[e] + w01 * v01.v[e] + w10 * v10.v[e] + w11 * v11.v[e];
  return r;
}

// ---- PagFM (model_utils.py:292-312), unfused training form.  Forward: g = sigmoid(sum_c xk_c * U(yq)_c) (saved),
// out = relu((1-g) x + g U(y)).  Threads: LPX lanes per pixel over C = 8*LPX channels of x; xk/yq have Cm channels.
template <int LPX>
__global__ void __launch_bounds__(256) pag_train_fwd_kernel(View x, View xk, View yq, View y, View out, float* gate) {
  const long gid = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long npix = static_cast<long>(x.N) * x.H * x.W;
  long pix = gid / LPX;
  const int cg = static_cast<int>(gid % LPX);
  const bool valid = pix < npix;
  if (!valid) pix = npix - 1;
  const int w = static_cast<int>(pix % x.W);
  const long t1 = pix / x.W;
  const int h = static_cast<int>(t1 % x.H);
  const int n = static_cast<int>(t1 / x.H);
  float dot = 0.f;
  if (cg * 8 < xk.C) {
    const F8 a = ld8(xk.ptr + pix * xk.ps + cg * 8);
    const F8 b = sample8(yq, n, h, w, x.H, x.W, cg * 8);
#pragma unroll
    for (int e = 0; e < 8; ++e) dot += a.v[e] * b.v[e];
  }
#pragma unroll
  for (int o = LPX / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  const float g = 1.f / (1.f + __expf(-dot));
  const F8 xv = ld8(x.ptr + pix * x.ps + cg * 8);
  const F8 yv = sample8(y, n, h, w, x.H, x.W, cg * 8);
  F8 o;
#pragma unroll
  for (int e = 0; e < 8; ++e) o.v[e] = fmaxf((1.f - g) * xv.v[e] + g * yv.v[e], 0.f);
  if (valid) {
    st8(out.ptr + pix * out.ps + cg * 8, o);
    if (cg == 0) gate[pix] = g;
  }
}
// Backward: dz' = dout*[out>0]; dg = sum_c dz'_c (U(y)_c - x_c); ds = dg g (1-g);
//   dx (+)= (1-g) dz';  dxk = ds U(yq);  t1 = ds xk  (-> U^T -> dyq);  t2 = g dz'  (-> U^T -> dy)
template <int LPX>
__global__ void __launch_bounds__(256) pag_train_bwd_kernel(View x, View xk, View yq, View y, View out, View dout,
                                                            const float* __restrict__ gate, View dx, int acc_dx, View dxk,
                                                            View t1, View t2) {
  const long gid = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long npix = static_cast<long>(x.N) * x.H * x.W;
  long pix = gid / LPX;
  const int cg = static_cast<int>(gid % LPX);
  const bool valid = pix < npix;
  if (!valid) pix = npix - 1;
  const int w = static_cast<int>(pix % x.W);
  const long tq = pix / x.W;
  const int h = static_cast<int>(tq % x.H);
  const int n = static_cast<int>(tq / x.H);
  const float g = gate[pix];
  F8 dz = ld8(dout.ptr + pix * dout.ps + cg * 8);
  const F8 ov = ld8(out.ptr + pix * out.ps + cg * 8);
  const F8 xv = ld8(x.ptr + pix * x.ps + cg * 8);
  const F8 yv = sample8(y, n, h, w, x.H, x.W, cg * 8);
  float dg = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if (!(ov.v[e] > 0.f)) dz.v[e] = 0.f;
    dg += dz.v[e] * (yv.v[e] - xv.v[e]);
  }
#pragma unroll
  for (int o = LPX / 2; o > 0; o >>= 1) dg += __shfl_xor_sync(0xffffffffu, dg, o);
  const float ds = dg * g * (1.f - g);
  if (!valid) return;
  F8 a, b;
#pragma unroll
  for (int e = 0; e < 8; ++e) { a.v[e] = (1.f - g) * dz.v[e]; b.v[e] = g * dz.v[e]; }
  if (acc_dx) {
    const F8 old = ld8(dx.ptr + pix * dx.ps + cg * 8);
#pragma unroll
    for (int e = 0; e < 8; ++e) a.v[e] += old.v[e];
  }
  st8(dx.ptr + pix * dx.ps + cg * 8, a);
  st8(t2.ptr + pix * t2.ps + cg * 8, b);
  if (cg * 8 < xk.C) {
    const F8 kv = ld8(xk.ptr + pix * xk.ps + cg * 8);
    const F8 qv = sample8(yq, n, h, w, x.H, x.W, cg * 8);
    F8 c, d;
#pragma unroll
    for (int e = 0; e < 8; ++e) { c.v[e] = ds * qv.v[e]; d.v[e] = ds * kv.v[e]; }
    st8(dxk.ptr + pix * dxk.ps + cg * 8, c);
    st8(t1.ptr + pix * t1.ps + cg * 8, d);
  }
}

// ---- Light_Bag backward: e = sigmoid(d); u = (1-e) i + p; v = i + e p
//   dp (+)= du + e dv;  dd (+)= (dv p - du i) e (1-e);  ti = (1-e) du + dv  (-> U^T -> di_low)
__global__ void __launch_bounds__(256) lightbag_bwd_kernel(View p, View il, View d, View duv, View dp, int acc_dp, View dd,
                                                           int acc_dd, View ti) {
  const int groups = p.C >> 3;
  const long total = static_cast<long>(p.N) * p.H * p.W * groups;
  const long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = static_cast<int>(idx % groups);
  const long pix = idx / groups;
  const int w = static_cast<int>(pix % p.W);
  const long t1 = pix / p.W;
  const int h = static_cast<int>(t1 % p.H);
  const int n = static_cast<int>(t1 / p.H);
  const F8 iv = sample8(il, n, h, w, p.H, p.W, cg * 8);
  const F8 pv = ld8(p.ptr + pix * p.ps + cg * 8);
  const F8 dv_ = ld8(d.ptr + pix * d.ps + cg * 8);
  const F8 du = ld8(duv.ptr + pix * duv.ps + cg * 8);
  const F8 dv = ld8(duv.ptr + pix * duv.ps + p.C + cg * 8);
  F8 op, od, ot;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float g = 1.f / (1.f + __expf(-dv_.v[e]));
    op.v[e] = du.v[e] + g * dv.v[e];
    od.v[e] = (dv.v[e] * pv.v[e] - du.v[e] * iv.v[e]) * g * (1.f - g);
    ot.v[e] = (1.f - g) * du.v[e] + dv.v[e];
  }
  if (acc_dp) {
    const F8 o = ld8(dp.ptr + pix * dp.ps + cg * 8);
#pragma unroll
    for (int e = 0; e < 8; ++e) op.v[e] += o.v[e];
  }
  if (acc_dd) {
    const F8 o = ld8(dd.ptr + pix * dd.ps + cg * 8);
#pragma unroll
    for (int e = 0; e < 8; ++e) od.v[e] += o.v[e];
  }
  st8(dp.ptr + pix * dp.ps + cg * 8, op);
  st8(dd.ptr + pix * dd.ps + cg * 8, od);
  st8(ti.ptr + pix * ti.ps + cg * 8, ot);
}

// ---- Bag (model_utils.py:375-377), train mode: e = sigmoid(d); out = e p + (1-e) U(i)   (BN + ReLU + conv follow)
__global__ void __launch_bounds__(256) bag_train_fwd_kernel(View p, View il, View d, View out) {
  const int groups = p.C >> 3;
  const long total = static_cast<long>(p.N) * p.H * p.W * groups;
  const long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = static_cast<int>(idx % groups);
  const long pix = idx / groups;
  const int w = static_cast<int>(pix % p.W);
  const long t1 = pix / p.W;
  const int h = static_cast<int>(t1 % p.H);
  const int n = static_cast<int>(t1 / p.H);
  const F8 iv = sample8(il, n, h, w, p.H, p.W, cg * 8);
  const F8 pv = ld8(p.ptr + pix * p.ps + cg * 8);
  const F8 dv = ld8(d.ptr + pix * d.ps + cg * 8);
  F8 o;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float g = 1.f / (1.f + __expf(-dv.v[e]));
    o.v[e] = g * pv.v[e] + (1.f - g) * iv.v[e];
  }
  st8(out.ptr + pix * out.ps + cg * 8, o);
}
//   dp (+)= e g;  dd (+)= g (p - U(i)) e (1-e);  ti = (1-e) g  (-> U^T -> di_low)
__global__ void __launch_bounds__(256) bag_train_bwd_kernel(View p, View il, View d, View dout, View dp, int acc_dp, View dd,
                                                            int acc_dd, View ti) {
  const int groups = p.C >> 3;
  const long total = static_cast<long>(p.N) * p.H * p.W * groups;
  const long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = static_cast<int>(idx % groups);
  const long pix = idx / groups;
  const int w = static_cast<int>(pix % p.W);
  const long t1 = pix / p.W;
  const int h = static_cast<int>(t1 % p.H);
  const int n = static_cast<int>(t1 / p.H);
  const F8 iv = sample8(il, n, h, w, p.H, p.W, cg * 8);
  const F8 pv = ld8(p.ptr + pix * p.ps + cg * 8);
  const F8 dv = ld8(d.ptr + pix * d.ps + cg * 8);
  const F8 go = ld8(dout.ptr + pix * dout.ps + cg * 8);
  F8 op, od, ot;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float g = 1.f / (1.f + __expf(-dv.v[e]));
    op.v[e] = g * go.v[e];
    od.v[e] = go.v[e] * (pv.v[e] - iv.v[e]) * g * (1.f - g);
    ot.v[e] = (1.f - g) * go.v[e];
  }
  if (acc_dp) {
    const F8 o = ld8(dp.ptr + pix * dp.ps + cg * 8);
#pragma unroll
    for (int e = 0; e < 8; ++e) op.v[e] += o.v[e];
  }
  if (acc_dd) {
    const F8 o = ld8(dd.ptr + pix * dd.ps + cg * 8);
#pragma unroll
    for (int e = 0; e < 8; ++e) od.v[e] += o.v[e];
  }
  st8(dp.ptr + pix * dp.ps + cg * 8, op);
  st8(dd.ptr + pix * dd.ps + cg * 8, od);
  st8(ti.ptr + pix * ti.ps + cg * 8, ot);
}

// ---- fused SGD on the flat parameter / gradient / momentum buffers (torch.optim.SGD semantics, tools/train.py:139-148):
//   d = g + wd * p;  buf = first ? d : momentum * buf + (1 - dampening) * d;  d = nesterov ? d + momentum * buf : buf;  p -= lr * d
__global__ void __launch_bounds__(256) sgd_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf,
                                                       long n4, float lr, float momentum, float dampening, float wd, int nesterov,
                                                       int first, float grad_scale) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 pv = reinterpret_cast<float4*>(p)[i];
  const float4 gv = reinterpret_cast<const float4*>(g)[i];
  float4 bv = first ? make_float4(0.f, 0.f, 0.f, 0.f) : reinterpret_cast<float4*>(buf)[i];
  float* pp = reinterpret_cast<float*>(&pv);
  const float* gp = reinterpret_cast<const float*>(&gv);
  float* bp = reinterpret_cast<float*>(&bv);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float d = gp[e] * grad_scale;
    if (wd != 0.f) d = d + wd * pp[e];
    if (momentum != 0.f) {
      bp[e] = first ? d : momentum * bp[e] + (1.f - dampening) * d;
      d = nesterov ? d + momentum * bp[e] : bp[e];
    }
    pp[e] = pp[e] - lr * d;
  }
  reinterpret_cast<float4*>(p)[i] = pv;
  if (momentum != 0.f) reinterpret_cast<float4*>(buf)[i] = bv;
}

__global__ void add_sums_kernel(double* __restrict__ sums, float* __restrict__ dst, int C, int Cacc) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) dst[c] += static_cast<float>(sums[c]);
  if (c < 2 * Cacc) sums[c] = 0.0;   // both halves of the [2][Cacc] accumulator
}

}  // namespace

static int reduce_geometry(const View& x, int num_sms, long& pix_per_block, int& threads, size_t& smem, unsigned& blocks) {
  const int groups = x.C / 8;
  if (groups < 1 || groups > 256 || x.C % 8) return -1;
  const int lanes = 256 / groups;
  threads = lanes * groups;
  const long npix = static_cast<long>(x.N) * x.H * x.W;
  // ~4 blocks per SM; at least 8 pixels per lane so the block-level reduction and atomics stay amortised
  const long target = static_cast<long>(num_sms) * 4;
  long ppb = (npix + target - 1) / target;
  const long unit = static_cast<long>(lanes) * 4;
  if (ppb < 2 * unit) ppb = 2 * unit;
  ppb = (ppb + unit - 1) / unit * unit;
  pix_per_block = ppb;
  blocks = static_cast<unsigned>((npix + ppb - 1) / ppb);
  smem = static_cast<size_t>(threads) * 16 * sizeof(float);
  return 0;
}

cudaError_t bn_stats_launch(View x, double* sums, int num_sms, cudaStream_t st) {
  long ppb; int threads; size_t smem; unsigned blocks;
  if (reduce_geometry(x, num_sms, ppb, threads, smem, blocks)) return cudaErrorInvalidValue;
  chan_reduce_kernel<0><<<blocks, threads, smem, st>>>(x, View{}, View{}, nullptr, nullptr, 0, ppb, sums);
  return cudaGetLastError();
}

cudaError_t bn_finalize_launch(double* sums, int C, double count, const float* gamma, const float* beta,
                               const float* conv_bias, float* mean, float* invstd, float* scale, float* shift,
                               float* run_mean, float* run_var, cudaStream_t st) {
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, C, count, gamma, beta, conv_bias, 1e-5f, 0.1f, mean, invstd,
                                                      scale, shift, run_mean, run_var);
  return cudaGetLastError();
}

cudaError_t bn_backward_launch(View x, View dz, View z, View dx, View dres, const float* mean, const float* invstd,
                               const float* gamma, double* sums, float* coef, int relu, int acc_dx, int acc_dres,
                               float* dgamma, float* dbeta, int num_sms, cudaStream_t st) {
  long ppb; int threads; size_t smem; unsigned blocks;
  if (reduce_geometry(x, num_sms, ppb, threads, smem, blocks)) return cudaErrorInvalidValue;
  chan_reduce_kernel<1><<<blocks, threads, smem, st>>>(x, dz, z, mean, invstd, relu, ppb, sums);
  const long total = static_cast<long>(x.N) * x.H * x.W * (x.C / 8);
  const double count = static_cast<double>(x.N) * x.H * x.W;
  bn_bwd_coef_kernel<<<(x.C + 127) / 128, 128, 0, st>>>(sums, x.C, count, gamma, mean, invstd, coef, dgamma, dbeta);
  if (dx.ptr || dres.ptr)
    bn_bwd_apply_kernel<<<blocks_for(total, 256), 256, 0, st>>>(x, dz, z, dx, dres, coef, relu, acc_dx, acc_dres);
  return cudaGetLastError();
}

// ---- fused single-launch BatchNorm
// shared memory per block (2 blocks per SM): small enough to co-reside with a wgrad CTA of the side stream
static constexpr size_t kBnSmemBudget = 24 * 1024;
// The fused BatchNorm kernels contain a grid barrier: they are launched COOPERATIVELY (cudaLaunchAttributeCooperative), so the
// runtime guarantees that all 2 x num_sms blocks are co-resident (or rejects the launch) instead of the kernel assuming it.
// PIDNET_BN_COOP=0 falls back to a plain launch (A/B measurements).
template <class P>
static cudaError_t launch_bn(void (*kernel)(P), unsigned blocks, int threads, size_t smem, cudaStream_t st, const P& p) {
  static const bool coop = [] { const char* v = std::getenv("PIDNET_BN_COOP"); return !(v && v[0] == '0'); }();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(blocks, 1, 1);
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = coop ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, p);
}
static cudaError_t bn_fused_geometry(const View& x, int num_sms, int vec_per_iter, int nt, unsigned& blocks, long& ppb,
                                     int& stage_iters, size_t& smem) {
  const int groups = x.C / 8;
  if (groups < 1 || groups > nt || x.C % 8) return cudaErrorInvalidValue;
  const int lanes = nt / groups;
  const long npix = static_cast<long>(x.N) * x.H * x.W;
  // full machine (2 blocks per SM) for the large tensors; small maps get fewer blocks -- every thread still has >= ~8 vectors and
  // the grid barrier collects fewer arrivals (its latency is most of a small launch)
  const long vectors = npix * groups;
  long want = (vectors + static_cast<long>(nt) * 8 - 1) / (static_cast<long>(nt) * 8);
  want = std::max<long>(std::min<long>(want, 2L * num_sms), std::max(1, num_sms / 4));
  blocks = static_cast<unsigned>(want);
  const long unit = static_cast<long>(lanes) * 4;
  ppb = (npix + blocks - 1) / blocks;
  ppb = (ppb + unit - 1) / unit * unit;
  const size_t red = static_cast<size_t>(bn_red_floats(x.C)) * sizeof(float);
  const size_t per_iter = static_cast<size_t>(nt) * 16 * vec_per_iter;
  stage_iters = red < kBnSmemBudget ? static_cast<int>((kBnSmemBudget - red) / per_iter) & ~3 : 0;   // whole trips of the kernels' loops
  smem = red + stage_iters * per_iter;
  return cudaSuccess;
}
// accumulators + arrival counters of ONE fused launch (256-byte multiple); the counters start at bn_fused_sync_offset
size_t bn_fused_sync_offset(int C) { return (static_cast<size_t>(kBnReplicas) * 2 * C * sizeof(double) + 255) / 256 * 256; }
size_t bn_fused_acc_bytes(int C) { return bn_fused_sync_offset(C) + kBnArriveSlots * kBnArriveStride * sizeof(unsigned); }
bool bn_fused_supported(int C) { return C % 8 == 0 && C / 8 >= 1 && C / 8 <= kBnBwdThreadsMin && 3 * C * sizeof(float) <= 40 * 1024; }

cudaError_t bn_forward_fused_launch(View x, View res, View z, const float* gamma, const float* beta, const float* conv_bias,
                                    float* mean, float* invstd, float* scale, float* shift, float* run_mean, float* run_var,
                                    double* sums, unsigned* sync, int relu, int num_sms, cudaStream_t st, int pre) {
  BnFwdParams p;
  unsigned blocks = 0;
  size_t smem = 0;
  cudaError_t e = bn_fused_geometry(x, num_sms, 1, kBnThreads, blocks, p.pix_per_block, p.stage_iters, smem);
  if (e != cudaSuccess) return e;
  p.x = x; p.res = res; p.z = z; p.gamma = gamma; p.beta = beta; p.conv_bias = conv_bias; p.mean = mean; p.invstd = invstd;
  p.run_mean = run_mean; p.run_var = run_var; p.sums = sums; p.sync = sync; p.relu = relu;
  p.scale = scale; p.shift = shift;
  p.count = static_cast<double>(x.N) * x.H * x.W;
  p.pre = pre ? 1 : 0;
  if (pre) {
    // statistics come from the conv epilogue: a plain streaming launch (no barrier => no co-residency requirement, nothing parked)
    p.stage_iters = 0;
    smem = static_cast<size_t>(bn_red_floats(x.C)) * sizeof(float);
    // (programmatic dependent launch was measured here: no gain, 3.455 vs 3.403 ms forward -- a conv CTA leaves no room for
    // staged BatchNorm blocks on its SM)
    bn_fwd_fused_kernel<<<blocks, kBnThreads, smem, st>>>(p);
    return cudaGetLastError();
  }
  return launch_bn(bn_fwd_fused_kernel, blocks, kBnThreads, smem, st, p);
}

cudaError_t bn_backward_fused_launch(View x, View dz, View z, View dx, View dres, const float* mean, const float* invstd,
                                     const float* gamma, const float* scale, const float* shift, int mask_x, double* sums,
                                     unsigned* sync, int relu, int acc_dx, int acc_dres, float* dgamma, float* dbeta, int num_sms,
                                     cudaStream_t st) {
  BnBwdParams p;
  unsigned blocks = 0;
  size_t smem = 0;
  // PIDNET_BN_BWD_THREADS=384: the wider block (80 registers x 768 threads per SM leave no room for a wgrad CTA of the side stream)
  static const int nt = [] { const char* v = std::getenv("PIDNET_BN_BWD_THREADS"); return v && std::atoi(v) == 384 ? 384 : kBnBwdThreadsMin; }();
  cudaError_t e = bn_fused_geometry(x, num_sms, 2, nt, blocks, p.pix_per_block, p.stage_iters, smem);
  if (e != cudaSuccess) return e;
  p.x = x; p.dz = dz; p.z = z; p.dx = dx; p.dres = dres; p.mean = mean; p.invstd = invstd; p.gamma = gamma;
  p.dgamma = dgamma; p.dbeta = dbeta; p.sums = sums; p.sync = sync; p.relu = relu; p.acc_dx = acc_dx; p.acc_dres = acc_dres;
  p.scale = scale; p.shift = shift; p.mask_x = (mask_x && relu) ? 1 : 0;
  p.count = static_cast<double>(x.N) * x.H * x.W;
  if (nt == 384)
    return dres.ptr ? launch_bn(bn_bwd_fused_kernel<true, 384>, blocks, nt, smem, st, p)
                    : launch_bn(bn_bwd_fused_kernel<false, 384>, blocks, nt, smem, st, p);
  return dres.ptr ? launch_bn(bn_bwd_fused_kernel<true, kBnBwdThreadsMin>, blocks, nt, smem, st, p)
                  : launch_bn(bn_bwd_fused_kernel<false, kBnBwdThreadsMin>, blocks, nt, smem, st, p);
}

cudaError_t pack_weights_launch(const PackJob& j, cudaStream_t st) {
  const long total = static_cast<long>(j.rows_pad) * j.ntaps * j.chunks * j.BK;
  pack_weights_kernel<<<blocks_for(total, 256), 256, 0, st>>>(j);
  return cudaGetLastError();
}

unsigned pack_job_blocks(const PackJob& j) {
  return blocks_for(static_cast<long>(j.rows_pad) * j.ntaps * j.chunks * j.BK, 256);
}
cudaError_t pack_all_launch(const PackJob* dev_jobs, const unsigned* dev_block_start, int njobs, unsigned total_blocks,
                            cudaStream_t st) {
  if (njobs <= 0) return cudaSuccess;
  pack_all_kernel<<<total_blocks, 256, 0, st>>>(dev_jobs, dev_block_start, njobs);
  return cudaGetLastError();
}

// --------------------------------------------------------------------------- stem weight gradient (see train_kernels.cuh)
namespace {
constexpr int kSwWarps = 4;   // warps per block; each warp works on its own 32-pixel chunks
__global__ void __launch_bounds__(kSwWarps * 32) stem_wgrad_kernel(const float* __restrict__ x, int N, int H, int W, View dy,
                                                                  float* __restrict__ dW) {
  __shared__ __align__(16) float xs[kSwWarps][32][32];          // im2col rows of the chunk (k = (ci*3+r)*3+s, 27..31 = 0)
  __shared__ __align__(16) __nv_bfloat16 dys[kSwWarps][32][32];  // dY rows of the chunk (this block's 32-channel slice)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Ho = dy.H, Wo = dy.W;
  const long P = static_cast<long>(N) * Ho * Wo;
  const long chunks = (P + 31) / 32;
  const int co0 = blockIdx.y * 32;
  const int cg = lane & 7, kg = lane >> 3;   // thread tile: co 4cg..4cg+3  x  k 8kg..8kg+7
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const long plane = static_cast<long>(H) * W;
  for (long c = static_cast<long>(blockIdx.x) * kSwWarps + warp; c < chunks; c += static_cast<long>(gridDim.x) * kSwWarps) {
    // ---- stage: lane = pixel
    const long pix = c * 32 + lane;
    float v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = 0.f;
    uint4 d4[4] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
    if (pix < P) {
      const int ow = static_cast<int>(pix % Wo);
      const long t1 = pix / Wo;
      const int oh = static_cast<int>(t1 % Ho);
      const int n = static_cast<int>(t1 / Ho);
      const float* xn = x + static_cast<long>(n) * 3 * plane;
      const int ih0 = 2 * oh - 1, iw0 = 2 * ow - 1;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int ih = ih0 + r;
          const bool rok = ih >= 0 && ih < H;
          const float* rp = xn + ci * plane + static_cast<long>(rok ? ih : 0) * W;
#pragma unroll
          for (int s2 = 0; s2 < 3; ++s2) {
            const int iw = iw0 + s2;
            if (rok && iw >= 0 && iw < W) v[(ci * 3 + r) * 3 + s2] = __ldg(rp + iw);
          }
        }
      const uint4* dp = reinterpret_cast<const uint4*>(dy.ptr + pix * dy.ps + co0);
#pragma unroll
      for (int q = 0; q < 4; ++q) d4[q] = __ldg(dp + q);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q)
      *reinterpret_cast<float4*>(&xs[warp][lane][q * 4]) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
#pragma unroll
    for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(&dys[warp][lane][q * 8]) = d4[q];
    __syncwarp();
    // ---- 32 pixels x (4 co x 8 k) FMAs per thread
#pragma unroll 4
    for (int pz = 0; pz < 32; ++pz) {
      const uint2 dq = *reinterpret_cast<const uint2*>(&dys[warp][pz][cg * 4]);
      const float4 xa = *reinterpret_cast<const float4*>(&xs[warp][pz][kg * 8]);
      const float4 xb = *reinterpret_cast<const float4*>(&xs[warp][pz][kg * 8 + 4]);
      const float dv[4] = {__uint_as_float(dq.x << 16), __uint_as_float(dq.x & 0xFFFF0000u), __uint_as_float(dq.y << 16),
                           __uint_as_float(dq.y & 0xFFFF0000u)};
      const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(dv[i], xv[j], acc[i][j]);
    }
    __syncwarp();
  }
  // block reduction through smem (the staging buffers are free now), then ONE atomic per output and block
  __syncthreads();
  float* red = &xs[0][0][0];   // [kSwWarps][32 co][32 k] floats == sizeof(xs)
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) red[(warp * 32 + cg * 4 + i) * 32 + kg * 8 + j] = acc[i][j];
  __syncthreads();
  for (int o = threadIdx.x; o < 32 * 32; o += kSwWarps * 32) {
    const int co = o >> 5, k = o & 31;
    if (k < 27) {
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < kSwWarps; ++w) sum += red[(w * 32 + co) * 32 + k];
      atomicAdd(dW + static_cast<long>(co0 + co) * 27 + k, sum);
    }
  }
}
}  // namespace

cudaError_t stem_wgrad_launch(const float* x, int N, int H, int W, View dy, float* dW, int num_sms, cudaStream_t st) {
  if (dy.C % 32 != 0 || dy.ps % 8 != 0) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(dW, 0, static_cast<size_t>(dy.C) * 27 * sizeof(float), st);
  if (e != cudaSuccess) return e;
  const long chunks = (static_cast<long>(N) * dy.H * dy.W + 31) / 32;
  long bx = static_cast<long>(num_sms) * 4;   // one resident wave (96 registers x 128 threads: 5 blocks per SM)
  if (bx * kSwWarps > chunks) bx = (chunks + kSwWarps - 1) / kSwWarps;
  stem_wgrad_kernel<<<dim3(static_cast<unsigned>(bx), dy.C / 32, 1), kSwWarps * 32, 0, st>>>(x, N, H, W, dy, dW);
  return cudaGetLastError();
}

cudaError_t nchw_to_nhwc_launch(const float* x, int N, int C, int H, int W, View out, cudaStream_t st) {
  nchw_to_nhwc_kernel<<<blocks_for(static_cast<long>(N) * H * W, 256), 256, 0, st>>>(x, N, C, H, W, out);
  return cudaGetLastError();
}

cudaError_t upsample_transpose_launch(View dhi, View dlow, int accumulate, cudaStream_t st) {
  const long total = static_cast<long>(dlow.N) * dlow.H * dlow.W * (dlow.C / 8);
  // rows of the gather window per output: split them over 32 / 8 cooperating lanes when the window is tall
  const int rows = 2 * ((dhi.H + dlow.H - 1) / dlow.H) + 3;
  if (rows >= 24 || total < 4096) upsample_transpose_kernel<32><<<blocks_for(total * 32, 256), 256, 0, st>>>(dhi, dlow, accumulate);
  else if (rows >= 16) upsample_transpose_kernel<8><<<blocks_for(total * 8, 256), 256, 0, st>>>(dhi, dlow, accumulate);
  else upsample_transpose_kernel<1><<<blocks_for(total, 256), 256, 0, st>>>(dhi, dlow, accumulate);
  return cudaGetLastError();
}

cudaError_t pool_transpose_launch(View dy, View dx, int k, int stride, int pad, int accumulate, cudaStream_t st) {
  const long total = static_cast<long>(dx.N) * dx.H * dx.W * (dx.C / 8);
  pool_transpose_kernel<<<blocks_for(total, 256), 256, 0, st>>>(dy, dx, k, stride, pad, accumulate);
  return cudaGetLastError();
}

cudaError_t pag_train_fwd_launch(View x, View xk, View yq, View y, View out, float* gate, cudaStream_t st) {
  const int LPX = x.C / 8;
  const long total = static_cast<long>(x.N) * x.H * x.W * LPX;
  const unsigned nb = blocks_for(total, 256);
  switch (LPX) {
    case 2: pag_train_fwd_kernel<2><<<nb, 256, 0, st>>>(x, xk, yq, y, out, gate); break;
    case 4: pag_train_fwd_kernel<4><<<nb, 256, 0, st>>>(x, xk, yq, y, out, gate); break;
    case 8: pag_train_fwd_kernel<8><<<nb, 256, 0, st>>>(x, xk, yq, y, out, gate); break;
    case 16: pag_train_fwd_kernel<16><<<nb, 256, 0, st>>>(x, xk, yq, y, out, gate); break;
    case 32: pag_train_fwd_kernel<32><<<nb, 256, 0, st>>>(x, xk, yq, y, out, gate); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t pag_train_bwd_launch(View x, View xk, View yq, View y, View out, View dout, const float* gate, View dx,
                                 int acc_dx, View dxk, View t1, View t2, cudaStream_t st) {
  const int LPX = x.C / 8;
  const long total = static_cast<long>(x.N) * x.H * x.W * LPX;
  const unsigned nb = blocks_for(total, 256);
  switch (LPX) {
    case 2: pag_train_bwd_kernel<2><<<nb, 256, 0, st>>>(x, xk, yq, y, out, dout, gate, dx, acc_dx, dxk, t1, t2); break;
    case 4: pag_train_bwd_kernel<4><<<nb, 256, 0, st>>>(x, xk, yq, y, out, dout, gate, dx, acc_dx, dxk, t1, t2); break;
    case 8: pag_train_bwd_kernel<8><<<nb, 256, 0, st>>>(x, xk, yq, y, out, dout, gate, dx, acc_dx, dxk, t1, t2); break;
    case 16: pag_train_bwd_kernel<16><<<nb, 256, 0, st>>>(x, xk, yq, y, out, dout, gate, dx, acc_dx, dxk, t1, t2); break;
    case 32: pag_train_bwd_kernel<32><<<nb, 256, 0, st>>>(x, xk, yq, y, out, dout, gate, dx, acc_dx, dxk, t1, t2); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t lightbag_bwd_launch(View p, View il, View d, View duv, View dp, int acc_dp, View dd, int acc_dd, View ti,
                                cudaStream_t st) {
  const long total = static_cast<long>(p.N) * p.H * p.W * (p.C / 8);
  lightbag_bwd_kernel<<<blocks_for(total, 256), 256, 0, st>>>(p, il, d, duv, dp, acc_dp, dd, acc_dd, ti);
  return cudaGetLastError();
}

cudaError_t bag_train_fwd_launch(View p, View il, View d, View out, cudaStream_t st) {
  const long total = static_cast<long>(p.N) * p.H * p.W * (p.C / 8);
  bag_train_fwd_kernel<<<blocks_for(total, 256), 256, 0, st>>>(p, il, d, out);
  return cudaGetLastError();
}
cudaError_t bag_train_bwd_launch(View p, View il, View d, View dout, View dp, int acc_dp, View dd, int acc_dd, View ti,
                                 cudaStream_t st) {
  const long total = static_cast<long>(p.N) * p.H * p.W * (p.C / 8);
  bag_train_bwd_kernel<<<blocks_for(total, 256), 256, 0, st>>>(p, il, d, dout, dp, acc_dp, dd, acc_dd, ti);
  return cudaGetLastError();
}

cudaError_t sgd_step_launch(float* p, const float* g, float* buf, long n, float lr, float momentum, float dampening, float wd,
                            int nesterov, int first, float grad_scale, cudaStream_t st) {
  if (n % 4 || (reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(buf)) % 16)
    return cudaErrorInvalidValue;
  const long n4 = n / 4;
  if (n4 == 0) return cudaSuccess;
  sgd_step_kernel<<<blocks_for(n4, 256), 256, 0, st>>>(p, g, buf, n4, lr, momentum, dampening, wd, nesterov, first, grad_scale);
  return cudaGetLastError();
}

cudaError_t add_sums_launch(double* sums, float* dst, int C, int Cacc, cudaStream_t st) {
  add_sums_kernel<<<(2 * Cacc + 127) / 128, 128, 0, st>>>(sums, dst, C, Cacc);
  return cudaGetLastError();
}

cudaError_t masked_add_launch(View a, View mask, View out, int accumulate, cudaStream_t st) {
  const long total = static_cast<long>(out.N) * out.H * out.W * (out.C / 8);
  masked_add_kernel<<<blocks_for(total, 256), 256, 0, st>>>(a, mask, out, accumulate);
  return cudaGetLastError();
}

}  // namespace pidnet
