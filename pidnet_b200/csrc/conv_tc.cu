// tcgen05 / TMEM / TMA implicit-GEMM convolution kernel -- see conv_tc.cuh for the contract.
#include "conv_tc.cuh"
#include "ptx.cuh"

namespace pidnet {

namespace {

constexpr int kThreads = 128;  // warp 0: TMA producer, warp 1: MMA issuer, warp 2: TMEM alloc; all 4: epilogue
constexpr int kTileM = 128;

template <int BN, int BK>
struct Cfg {
  static constexpr int kABytes = kTileM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // pipeline depth: keep >= 2 CTAs per SM where the tile allows it (epilogue of one overlaps the
  // main loop of the other); the 128-wide tile takes the SM alone with a deeper ring.
  // The 128-wide tile used to take the SM alone (4 stages + staging = 162 KB): with one non-persistent CTA per SM nothing
  // overlapped its prologue (TMEM alloc, first TMA round trip) and epilogue.  It now runs 3 stages and ALIASES the output
  // staging tile onto the (by then drained) stage ring, which fits two CTAs per SM.
  static constexpr bool kAliasOut = (BN == 128 && BK == 64);
  static constexpr int kStages = kAliasOut ? 3 : ((BN == 128) ? 4 : ((BN == 64 && BK == 64) ? 3 : 4));
  static constexpr int kSlabC = BN < 64 ? BN : 64;  // channels per staged output slab
  static constexpr int kSlabRowBytes = kSlabC * 2;  // == swizzle span of tmD / tmR
  static constexpr int kSlabBytes = kTileM * kSlabRowBytes;
  static constexpr int kNumSlabs = BN / kSlabC;
  static constexpr int kOutBytes = kTileM * BN * 2;
  static constexpr int kBarBytes = 256;
  static constexpr int kOutOffset = kAliasOut ? 0 : kStages * kStageBytes;
  static constexpr int kTailOffset = kAliasOut ? kStages * kStageBytes : kStages * kStageBytes + kOutBytes;   // bias, barriers
  static_assert(!kAliasOut || kStages * kStageBytes >= kOutBytes, "staging tile must fit in the stage ring");
  static constexpr int kSmem = kTailOffset + BN * 4 + kBarBytes + 1024 /*align slack*/;
};

__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int BN, int BK>
__global__ void __launch_bounds__(kThreads, Cfg<BN, BK>::kAliasOut ? 2 : 1) conv_tc_kernel(const __grid_constant__ ConvParams p) {
  using C = Cfg<BN, BK>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const uint32_t stage_base = smem_base;
  const uint32_t out_base = smem_base + C::kOutOffset;
  uint8_t* out_gen = smem_gen + C::kOutOffset;
  float* bias_s = reinterpret_cast<float*>(smem_gen + C::kTailOffset);
  const uint32_t bar_base = smem_base + C::kTailOffset + BN * 4;
  // barriers: full[kStages], empty[kStages], acc_full, res_full ; then the TMEM base address word
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::kStages + s); };
  const uint32_t acc_bar = bar_base + 8u * (2 * C::kStages);
  const uint32_t res_bar = acc_bar + 8u;
  const uint32_t tmem_slot = res_bar + 8u;
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + C::kTailOffset + BN * 4 + 8 * (2 * C::kStages + 2));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- tile coordinates
  const int m_tile = blockIdx.x;
  const int n_tile = blockIdx.y;
  const int per_group = p.tiles_w * p.tiles_h;
  const int tn = m_tile / per_group;
  const int rem = m_tile - tn * per_group;
  const int th = rem / p.tiles_w;
  const int tw = rem - th * p.tiles_w;
  const int w0 = tw * p.TW, h0 = th * p.TH, n0 = tn * p.TN;
  const int c_out0 = n_tile * BN;

  int num_k = 0;
  for (int s = 0; s < p.nsrc; ++s) num_k += p.src[s].ntaps * p.src[s].chunks;

  // ---- one-time setup (overlaps the previous kernel's tail under programmatic dependent launch, see ptx.cuh)
  if (threadIdx.x == 0) pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(acc_bar, 1);
    mbar_init(res_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<BN>(tmem_slot);
  pdl_wait();   // global memory is touched only from here on
  if (threadIdx.x < BN) bias_s[threadIdx.x] = p.bias[c_out0 + threadIdx.x];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot_gen;

  if (warp == 0) {
    // ======================= TMA producer =======================
    if (elect_one()) {
      tma_prefetch_desc(&p.tmB);
      auto load_residual = [&] {
        mbar_arrive_expect_tx(res_bar, C::kOutBytes);
        for (int sl = 0; sl < C::kNumSlabs; ++sl)
          tma_load_4d(out_base + sl * C::kSlabBytes, &p.tmR, res_bar, c_out0 + sl * C::kSlabC, w0, h0, n0);
      };
      if (p.has_res && !C::kAliasOut) load_residual();
      int ks = 0;
      for (int s = 0; s < p.nsrc; ++s) {
        const ConvSrc& src = p.src[s];
        for (int t = 0; t < src.ntaps; ++t) {
          const uint32_t tap = src.taps[t];
          const CUtensorMap* mapA = &p.tmA[tap & 0xFF];
          const int dh = static_cast<int>((tap >> 8) & 0xFF) - 8;
          const int dw = static_cast<int>((tap >> 16) & 0xFF) - 8;
          for (int cc = 0; cc < src.chunks; ++cc, ++ks) {
            const int st = ks % C::kStages;
            const uint32_t ph = (ks / C::kStages) & 1;
            mbar_wait(empty_bar(st), ph ^ 1);
            mbar_arrive_expect_tx(full_bar(st), C::kStageBytes);
            const uint32_t a_dst = stage_base + st * C::kStageBytes;
            tma_load_4d(a_dst, mapA, full_bar(st), cc * BK, w0 + dw, h0 + dh, n0);
            tma_load_2d(a_dst + C::kABytes, &p.tmB, full_bar(st), ks * BK, c_out0);
          }
        }
      }
      if (p.has_res && C::kAliasOut) {   // the staging tile aliases the ring: wait until every MMA has read its stage
        mbar_wait(acc_bar, 0);
        load_residual();
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(kTileM, BN);
      for (int ks = 0; ks < num_k; ++ks) {
        const int st = ks % C::kStages;
        const uint32_t ph = (ks / C::kStages) & 1;
        mbar_wait(full_bar(st), ph);
        tc_fence_after();
        const uint32_t a_addr = stage_base + st * C::kStageBytes;
        const uint64_t a_desc = make_kmajor_desc(a_addr, BK * 2);
        const uint64_t b_desc = make_kmajor_desc(a_addr + C::kABytes, BK * 2);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advance 16 elements (32 B) along K inside the swizzle atom: +2 in the (addr >> 4) field
          umma_bf16(tmem_acc, a_desc + 2 * k, b_desc + 2 * k, idesc, (ks | k) != 0 ? 1u : 0u);
        }
        umma_commit(empty_bar(st));  // frees the smem stage when these MMAs retire
      }
      umma_commit(acc_bar);  // accumulator complete
    }
    __syncwarp();
  }

  // ======================= epilogue (all 4 warps; warp w owns TMEM lanes 32w..32w+31) =======================
  mbar_wait(acc_bar, 0);
  tc_fence_after();
  if (p.has_res) mbar_wait(res_bar, 0);

  const int row = threadIdx.x;  // tile row == TMEM lane
  const uint32_t t_row = tmem_acc + (static_cast<uint32_t>(warp * 32) << 16);

  if (p.out_mode == kOutNHWCbf16 || p.out_mode == kOutNHWCsplit) {
    const uint32_t swz = (C::kSlabRowBytes == 128) ? (row & 7) : (C::kSlabRowBytes == 64 ? ((row >> 1) & 3) : ((row >> 2) & 1));
    const int npass = p.out_mode == kOutNHWCsplit ? 2 : 1;   // split storage: pass 0 writes hi = bf16(y), pass 1 lo = bf16(y - hi)
    for (int pass = 0; pass < npass; ++pass) {
#pragma unroll
      for (int g = 0; g < BN / 32; ++g) {
        uint32_t v[32];
        tmem_ld32(t_row + g * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = g * 32 + j * 8;
          const int slab = c / C::kSlabC;
          const int chunk = (c % C::kSlabC) / 8;
          uint4* ptr = reinterpret_cast<uint4*>(out_gen + slab * C::kSlabBytes + row * C::kSlabRowBytes +
                                                ((chunk ^ swz) << 4));
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[j * 8 + e]) + bias_s[c + e];
          if (p.has_res) {
            const uint4 r = *ptr;
            f[0] += bf16lo(r.x); f[1] += bf16hi(r.x); f[2] += bf16lo(r.y); f[3] += bf16hi(r.y);
            f[4] += bf16lo(r.z); f[5] += bf16hi(r.z); f[6] += bf16lo(r.w); f[7] += bf16hi(r.w);
          }
          if (p.relu) {
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
          }
          if (pass == 1) {
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] -= __bfloat162float(__float2bfloat16_rn(f[e]));
          }
          uint4 o;
          o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]);
          o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
          *ptr = o;
        }
      }
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the TMA store
      tc_fence_before();
      __syncthreads();
      if (threadIdx.x == 0) {
        for (int sl = 0; sl < C::kNumSlabs; ++sl) {
          if (c_out0 + sl * C::kSlabC < p.Cout)
            tma_store_4d(&p.tmD, out_base + sl * C::kSlabBytes, pass * p.Cout + c_out0 + sl * C::kSlabC, w0, h0, n0);
        }
        tma_store_commit();
        tma_store_wait_all();
      }
      if (pass + 1 < npass) {   // the staging tile is rewritten by the second pass: wait until the store has read it
        __syncthreads();
        tc_fence_after();
      }
    }
  } else {
    // fp32 NCHW planes (logits): thread == pixel; consecutive lanes == consecutive w -> coalesced per plane
    const int iw = row % p.TW;
    const int ih = (row / p.TW) % p.TH;
    const int in = row / (p.TW * p.TH);
    const int w = w0 + iw, h = h0 + ih, n = n0 + in;
    const bool ok = (w < p.Wo) && (h < p.Ho) && (n < p.N);
    const size_t plane = static_cast<size_t>(p.Ho) * p.Wo;
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
    __syncthreads();
  }
  if (warp == 2) tmem_dealloc<BN>(tmem_acc);
}

template <int BN, int BK>
cudaError_t launch_inst(const ConvLaunch& L, cudaStream_t stream) {
  return launch_pdl(conv_tc_kernel<BN, BK>, L.grid, dim3(kThreads, 1, 1), Cfg<BN, BK>::kSmem, stream, L.p);
}

template <int BN, int BK>
cudaError_t init_inst() {
  return cudaFuncSetAttribute(conv_tc_kernel<BN, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              Cfg<BN, BK>::kSmem);
}

}  // namespace

#define PIDNET_CONV_DISPATCH(FN, ...)                                   \
  do {                                                                  \
    if (BN == 32 && BK == 32) return FN<32, 32>(__VA_ARGS__);           \
    if (BN == 32 && BK == 64) return FN<32, 64>(__VA_ARGS__);           \
    if (BN == 64 && BK == 32) return FN<64, 32>(__VA_ARGS__);           \
    if (BN == 64 && BK == 64) return FN<64, 64>(__VA_ARGS__);           \
    if (BN == 128 && BK == 32) return FN<128, 32>(__VA_ARGS__);         \
    if (BN == 128 && BK == 64) return FN<128, 64>(__VA_ARGS__);         \
  } while (0)

cudaError_t conv_tc_launch(const ConvLaunch& L, cudaStream_t stream) {
  const int BN = L.BN, BK = L.BK;
  PIDNET_CONV_DISPATCH(launch_inst, L, stream);
  return cudaErrorInvalidValue;
}

cudaError_t conv_tc_init() {
  cudaError_t e;
  if ((e = init_inst<32, 32>()) != cudaSuccess) return e;
  if ((e = init_inst<32, 64>()) != cudaSuccess) return e;
  if ((e = init_inst<64, 32>()) != cudaSuccess) return e;
  if ((e = init_inst<64, 64>()) != cudaSuccess) return e;
  if ((e = init_inst<128, 32>()) != cudaSuccess) return e;
  if ((e = init_inst<128, 64>()) != cudaSuccess) return e;
  return cudaSuccess;
}

static size_t smem_of(int BN, int BK) {
  if (BN == 32 && BK == 32) return Cfg<32, 32>::kSmem;
  if (BN == 32 && BK == 64) return Cfg<32, 64>::kSmem;
  if (BN == 64 && BK == 32) return Cfg<64, 32>::kSmem;
  if (BN == 64 && BK == 64) return Cfg<64, 64>::kSmem;
  if (BN == 128 && BK == 32) return Cfg<128, 32>::kSmem;
  if (BN == 128 && BK == 64) return Cfg<128, 64>::kSmem;
  return 0;
}
size_t conv_tc_smem_bytes(int BN, int BK) { return smem_of(BN, BK); }

}  // namespace pidnet
