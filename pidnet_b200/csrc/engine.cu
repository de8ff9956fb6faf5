// pidnet_b200 engine: eval-mode planner for PIDNet (BN folding, weight packing, HBM arena, TMA tensor
// maps, three branch lanes joined by events, optional CUDA-graph replay) behind the C ABI of
// include/pidnet_b200.h.  The dataflow follows models/pidnet.py:136-182 of the reference; the fusion
// plan is described in DESIGN.md.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/pidnet_b200.h"
#include "conv_tc.cuh"
#include "kernels.cuh"
#include "criterion.cuh"
#include "train_kernels.cuh"

#ifdef PIDNET_PROBES   // hardware probes (probe.cu): only in the tools' libpidnet_b200_probe.so, never in the product library
namespace pidnet {
struct ProbeParams {
  CUtensorMap tmA, tmB;
  int r, s, mode;
  float* out;
};
cudaError_t halo_probe_launch(const ProbeParams& p, cudaStream_t st);
cudaError_t mma_rate_launch(int N, int iters, int distinct, int blocks, long long* out, cudaStream_t st);
struct MnProbeParams {
  CUtensorMap tmA, tmB;
  int lbo_a, sbo, variant;
  float* out;
};
cudaError_t mn_probe_launch(const MnProbeParams& p, cudaStream_t st);
struct PairProbeParams {
  CUtensorMap tmA, tmB;
  int swap_b;
  float* out;
};
cudaError_t pair_probe_launch(const PairProbeParams& p, cudaStream_t st);
cudaError_t mma_rate_pair_launch(int N, int iters, int distinct, int pairs, long long* out, cudaStream_t st);
}  // namespace pidnet
#endif

namespace pidnet {

// ----------------------------------------------------------------------------------------- errors
static thread_local std::string g_err;
struct Err : std::runtime_error {
  using std::runtime_error::runtime_error;
};
[[noreturn]] static void fail(const std::string& m) { throw Err(m); }
#define CK(call)                                                                                      \
  do {                                                                                                \
    cudaError_t e__ = (call);                                                                         \
    if (e__ != cudaSuccess) fail(std::string(#call) + ": " + cudaGetErrorString(e__));                \
  } while (0)

// ----------------------------------------------------------------------------------------- tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess) fail("cuTensorMapEncodeTiled entry point not available");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
static CUtensorMapSwizzle swizzle_for(int bytes) {
  switch (bytes) {
    case 128: return CU_TENSOR_MAP_SWIZZLE_128B;
    case 64: return CU_TENSOR_MAP_SWIZZLE_64B;
    case 32: return CU_TENSOR_MAP_SWIZZLE_32B;
    case 0: return CU_TENSOR_MAP_SWIZZLE_NONE;
  }
  fail("bad swizzle span");
}
// bf16 tensor of rank `rank` (dims fastest first, strides in BYTES for dims 1..rank-1)
static CUtensorMap encode_map(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_b,
                              const uint32_t* box, int swizzle_bytes,
                              CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16) {
  CUtensorMap m;
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (dims[i] == 0) fail("tensor map: zero dim");
    if (box[i] == 0 || box[i] > 256) fail("tensor map: bad box");
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gs[i] = strides_b[i];
    if (gs[i] % 16 != 0) fail("tensor map: stride not multiple of 16 B");
  }
  if (reinterpret_cast<uintptr_t>(base) % 16 != 0) fail("tensor map: base not 16 B aligned");
  CUresult r = get_encode()(&m, dtype, rank, const_cast<void*>(base), gd, gs, bx, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(swizzle_bytes),
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) fail("cuTensorMapEncodeTiled failed with code " + std::to_string(static_cast<int>(r)));
  return m;
}

// ----------------------------------------------------------------------------------------- host helpers
static inline uint16_t f2bf(float f) {  // round to nearest even
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return static_cast<uint16_t>(u >> 16);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}
static inline float bf2f(uint16_t h) {
  uint32_t u = static_cast<uint32_t>(h) << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct HostParam {
  std::vector<float> data;
  std::vector<int64_t> shape;
};
struct Affine {
  std::vector<float> s, t;
};

struct T {  // NHWC bf16 view
  bf16* ptr = nullptr;
  int N = 0, H = 0, W = 0, C = 0;
  long ps = 0;    // pixel stride (elements)
  int prod = -1;  // index of the producing op (-1: external)
  long sH = 0, sN = 0;  // explicit row / image strides in elements (0: dense, derived from W and H); sW == ps
  View view() const { return View{ptr, N, H, W, C, ps}; }
  long row_stride() const { return sH ? sH : static_cast<long>(W) * ps; }
  long img_stride() const { return sN ? sN : static_cast<long>(H) * W * ps; }
  bool dense() const { return sH == 0 && sN == 0; }
};

struct RunArgs {
  const float* x;
  float* out[3];  // main, p, d
  // alternative input: uint8 HWC BGR frames + the input_transform constants (pidnet_forward_u8); x is then null
  const uint8_t* x_u8 = nullptr;
  const float* lut = nullptr;   // device [3][256]: input_transform of every byte value per model channel
};

struct Op {
  std::string name;
  std::string kernel;   // kernel family label (profiling / roofline)
  double flops = 0;     // algorithmic FLOPs (2*MACs, no padding waste)
  double bytes = 0;     // algorithmic bytes: every distinct input and the output once
  int lane = 0;
  std::vector<int> deps;  // producer ops on other lanes
  std::vector<float*> gwrites;   // training backward: parameter-gradient pointers this op finalises (bucketed all-reduce)
  bool record = false;
  std::function<cudaError_t(cudaStream_t, const RunArgs&)> fn;
};

struct TapSpec {
  int dh, dw;  // input offset (in pixels of `in`) relative to the output pixel
  int r, s;    // which tap of the weight tensor multiplies it
};
struct ConvSrcSpec {
  T in;
  std::vector<float> w;  // [Cout][Cin][k][k] fp32, BN-folded (eval); empty when dev_w is set
  int k = 1, stride = 1;
  // training: weights live on the device as fp32 [Co][Cin_total][k][k] and are re-packed every step
  const float* dev_w = nullptr;
  int dev_cin_total = 0, dev_ci_off = 0;
  int dev_cin = 0;               // valid channels of `in` for packing (0: in.C); e.g. 3 of the 8-channel padded image
  bool dgrad = false;            // pack transposed (rows = ci, channels = co): this conv computes an input gradient
  std::vector<TapSpec> xtaps;    // explicit tap list (stride-2 dgrad parities); stride must be 1
};

// ----------------------------------------------------------------------------------------- builder
static int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v && *v ? std::atoi(v) : dflt;
}

struct Builder {
  bool dry = true;
  int conv_impl = 0;
  int use_ws = 1;   // weight-stationary halo-patch kernel for 3x3 stride-1 convs
  int use_stem2 = env_int("PIDNET_STEM2", 2);       // fused conv1.0 -> conv1.3 kernel (stem2_tc.cu): 1 lock-step, 2 pipelined (C = 32)
  int use_pyramid = env_int("PIDNET_POOL_PYRAMID", 1);   // PAPPM / DAPPM pooled branches from one summed-area-table kernel
  int use_pair = env_int("PIDNET_WS_PAIR", 1);      // CTA pairs (tcgen05 cta_group::2) where conv3_ws has the instance
  int ws_stages = env_int("PIDNET_WS_STAGES", 3);   // rotating staging buffers of the weight-stationary kernels (2 or 3)
  int num_sms = 148;
  size_t act_cur = 0, wt_cur = 0;
  uint8_t* act_base = nullptr;
  uint8_t* wt_base = nullptr;
  std::vector<uint8_t> wt_host;
  std::vector<Op> ops;
  int lane = 0;
  double flops = 0;
  std::map<std::string, T> named;
  const float* dev_bias_override = nullptr;   // training: conv bias read straight from the fp32 parameter
  bool split_next = false;   // the next conv() stores its result in split-bf16 form ([hi | lo], 2*Cout channels; conv_tc.cuh)
  // training: the next conv() feeds a BatchNorm whose fp64 accumulator table is `stats_next` ([kBnStatReplicas][2 * Cout]); if the
  // conv lands on a kernel that can accumulate the batch statistics in its epilogue (conv3_ws.cu) it does, and stats_taken says so
  double* stats_next = nullptr;
  bool stats_taken = false;
  int stats_kinds = 3;   // bit 0: 3x3 halo-patch kernel, bit 1: 1x1 flat-tile kernel
  std::vector<PackJob> pack_jobs;   // training: device-side weight packing, run at the start of every step

  void reset(bool dry_) {
    pack_jobs.clear();
    dry = dry_;
    act_cur = wt_cur = 0;
    ops.clear();
    named.clear();
    lane = 0;
    flops = 0;
    if (!dry) wt_host.assign(wt_host.size(), 0);
  }
  void* alloc_act(size_t bytes) {
    size_t off = act_cur;
    act_cur = align_up(act_cur + bytes, 1024);
    return act_base + off;
  }
  // device copy of host data (uploaded in one memcpy when the plan is finished)
  void* alloc_wt(const void* src, size_t bytes) {
    size_t off = wt_cur;
    wt_cur = align_up(wt_cur + bytes, 256);
    if (!dry) {
      if (wt_cur > wt_host.size()) fail("weight arena overflow");
      std::memcpy(wt_host.data() + off, src, bytes);
    }
    return wt_base + off;
  }
  const float* upload_f32(const std::vector<float>& v) {
    return reinterpret_cast<const float*>(alloc_wt(v.data(), v.size() * sizeof(float)));
  }
  T new_tensor(int N, int H, int W, int C) {
    if (C % 8 != 0) fail("channel count must be a multiple of 8 (got " + std::to_string(C) + ")");
    T t;
    t.N = N; t.H = H; t.W = W; t.C = C; t.ps = C;
    t.ptr = reinterpret_cast<bf16*>(alloc_act(static_cast<size_t>(N) * H * W * C * 2));
    return t;
  }
  static T slice(const T& t, int coff, int C) {
    if (coff % 8 != 0 || C % 8 != 0 || coff + C > t.C) fail("bad channel slice");
    T s = t;
    s.ptr = t.ptr + coff;
    s.C = C;
    return s;
  }
  int add_op(const std::string& name, std::vector<const T*> inputs,
             std::function<cudaError_t(cudaStream_t, const RunArgs&)> fn) {
    Op op;
    op.name = name;
    op.lane = lane;
    for (const T* t : inputs)
      if (t && t->prod >= 0 && ops[t->prod].lane != lane) {
        op.deps.push_back(t->prod);
        ops[t->prod].record = true;
      }
    op.fn = std::move(fn);
    ops.push_back(std::move(op));
    return static_cast<int>(ops.size()) - 1;
  }

  // --------------------------------------------------------------------------------- conv
  static void choose_tile(int N, int H, int W, bool nchw, int& TN, int& TH, int& TW) {
    long best = -1;
    for (int tw = 128; tw >= 1; tw >>= 1) {
      for (int th = 128 / tw; th >= 1; th >>= 1) {
        const int tn = 128 / (tw * th);
        if (nchw && tw < 32 && tw < W) continue;
        if (tn > 1 && (th < H || tw < W)) continue;  // batch several images only when one tile covers an image
        const long tiles = static_cast<long>(cdiv(W, tw)) * cdiv(H, th) * cdiv(N, tn);
        const long cost = tiles * 100000 + static_cast<long>(th + 2) * (tw + 2) * 10 + (128 - tw) / 16;
        if (best < 0 || cost < best) {
          best = cost;
          TN = tn; TH = th; TW = tw;
        }
      }
    }
  }

  // out = act( sum_src conv(src) + bias (+res) ); writes `out` (bf16 NHWC view) or the fp32 NCHW runtime
  // Would a dense single-source 3x3 stride-1 conv Cin -> Cout run on the weight-stationary halo kernel (same rule as below)?
  bool ws3_eligible(int Cin, int Cout) const {
    if (conv_impl != 0 || !use_ws) return false;
    const int BK = Cin > 32 ? 64 : 32;
    const int chunks = cdiv(Cin, BK);
    for (int bn = 128; bn >= 32; bn >>= 1) {
      if (bn > 32 && bn / 2 >= Cout) continue;
      if (bn == 32 && Cout > 32) continue;
      size_t smem = 0;
      if (conv3_ws_plan(0, bn, BK, chunks, 0, ws_stages, nullptr, &smem) >= 2) return true;
    }
    return false;
  }
  // output `out_slot` (0..2).  Returns the output view (invalid ptr for NCHW slots).
  T conv(const std::string& name, const std::vector<ConvSrcSpec>& srcs, const std::vector<float>& bias, int Cout,
         bool relu, const T* res, const T* out_view, int out_slot) {
    if (srcs.empty() || srcs.size() > 2) fail(name + ": 1 or 2 sources supported");
    const bool split = split_next;
    split_next = false;
    double* const stats = stats_next;
    stats_next = nullptr;
    stats_taken = false;
    if (split && (conv_impl != 0 || res || out_view || out_slot >= 0)) fail(name + ": split-bf16 output needs the tcgen05 path, no residual and its own tensor");
    const T& in0 = srcs[0].in;
    const int N = in0.N;
    const bool xt = !srcs[0].xtaps.empty();
    if (xt && (!out_view || srcs.size() != 1)) fail(name + ": explicit taps need one source and an output view");
    const int Ho = xt ? out_view->H : cdiv(in0.H, srcs[0].stride), Wo = xt ? out_view->W : cdiv(in0.W, srcs[0].stride);
    bool flat = out_slot < 0 && !xt && (!out_view || out_view->dense()) && (!res || res->dense());
    int BK = 32;
    for (const auto& s : srcs) {
      if (s.k != 1 && s.k != 3) fail(name + ": kernel size must be 1 or 3");
      if (s.stride != 1 && s.stride != 2) fail(name + ": stride must be 1 or 2");
      if (!xt && (cdiv(s.in.H, s.stride) != Ho || cdiv(s.in.W, s.stride) != Wo || s.in.N != N))
        fail(name + ": source geometry mismatch");
      if (!s.dev_w && static_cast<long>(s.w.size()) != static_cast<long>(Cout) * s.in.C * s.k * s.k)
        fail(name + ": weight size mismatch");
      if (!s.in.dense()) fail(name + ": conv sources must be dense NHWC views");
      if (s.k != 1 || s.stride != 1) flat = false;
      if (s.in.C > 32) BK = 64;
      flops += 2.0 * N * Ho * Wo * Cout * s.in.C * s.k * s.k;
    }
    T out;
    if (out_slot < 0) {
      out = out_view ? *out_view : new_tensor(N, Ho, Wo, split ? 2 * Cout : Cout);
      if (out.N != N || out.H != Ho || out.W != Wo || out.C != (split ? 2 * Cout : Cout)) fail(name + ": output view mismatch");
    } else {
      out.N = N; out.H = Ho; out.W = Wo; out.C = Cout;
    }
    if (res && (res->N != N || res->H != Ho || res->W != Wo || res->C != Cout)) fail(name + ": residual mismatch");

    // logical tiling geometry
    const int gN = flat ? 1 : N, gH = flat ? 1 : Ho, gW = flat ? N * Ho * Wo : Wo;
    int TN, TH, TW;
    choose_tile(gN, gH, gW, out_slot >= 0, TN, TH, TW);
    const int tiles_w = cdiv(gW, TW), tiles_h = cdiv(gH, TH), tiles_n = cdiv(gN, TN);
    const int m_tiles = tiles_w * tiles_h * tiles_n;
    int BN = Cout >= 128 ? 128 : (Cout > 32 ? 64 : 32);
    while (BN > 32 && static_cast<long>(m_tiles) * cdiv(Cout, BN) < num_sms) BN >>= 1;
    // persistent weight-stationary kernels (conv3_ws.cu): mode 0 = single-source 3x3 stride-1 (halo patch),
    // mode 1 = 1x1 stride-1 (<= 2 sources) over flattened pixels; used when the Cout tile's weights fit in smem
    bool ws = false;
    int ws_mode = 0, ws_np = 0, ws_chunks = 0, ws_nstage = 2, ws_pair = 0;
    size_t ws_smem = 0;
    if (conv_impl == 0 && use_ws && !split && !(res && out_slot >= 0)) {
      bool all1x1 = true;
      for (const auto& sp : srcs) {
        ws_chunks += cdiv(sp.in.C, BK);
        if (sp.k != 1 || sp.stride != 1) all1x1 = false;
      }
      const bool strided_io = xt || (out_view && !out_view->dense()) || (res && !res->dense());
      const bool m0 = !strided_io && srcs.size() == 1 && srcs[0].k == 3 && srcs[0].stride == 1;
      if (strided_io) all1x1 = false;
      if (m0 || all1x1) {
        ws_mode = m0 ? 0 : 1;
        for (int bn = 128; bn >= 32 && !ws; bn >>= 1) {
          if (bn > 32 && bn / 2 >= Cout) continue;      // tile wider than needed
          if (m0 && bn == 32 && Cout > 32) continue;    // N=32 MMAs re-reading the patch per Cout tile lose to the generic kernel
          const int np = conv3_ws_plan(ws_mode, bn, BK, ws_chunks, 0, ws_stages, &ws_nstage, &ws_smem);
          if (np >= (m0 ? 2 : 4)) {
            ws = true;
            ws_np = np;
            BN = bn;
          }
        }
      }
    }
    // CTA pairs (conv3_ws PAIR instance): half of the weights per CTA and M = 256 MMAs.  Used where the layer is MMA-bound
    // and the single-CTA ring cannot prefetch, i.e. Cin >= 128 (two or more chunks): measured 0.098 -> 0.069 ms (C = 128
    // @64x128) and 0.348 -> 0.238 ms (final_layer.conv1, 1.30 PFLOP/s).  The Cin = 64 layers stream more than they compute
    // and lose ~10-20 % to the pair's lock step (either CTA's memory stall holds both), so they stay single-CTA
    // (use_pair = 2 forces pairs wherever the instance exists; it also admits layers whose weights only fit when halved).
    if (conv_impl == 0 && use_ws && !split && use_pair && ws_mode == 0 && !(res && out_slot >= 0) && srcs.size() == 1 &&
        srcs[0].k == 3 && srcs[0].stride == 1 && Cout > 32 && conv3_ws_pair_available(0, 64, BK) &&
        (ws ? BN == 64 : use_pair >= 2) && (ws_chunks >= 2 || use_pair >= 2)) {
      const bool strided_io = xt || (out_view && !out_view->dense()) || (res && !res->dense());
      const long mt = static_cast<long>(N) * cdiv(Wo, 8) * cdiv(Ho, 16);
      int nst = 2;
      size_t sm = 0;
      const int np = strided_io ? 0 : conv3_ws_plan(0, 64, BK, ws_chunks, 1, ws_stages, &nst, &sm);
      if (np >= 2 && mt >= 2 && num_sms / cdiv(Cout, 64) >= 2) {
        ws = true; BN = 64;
        ws_pair = 1; ws_np = np; ws_nstage = nst; ws_smem = sm;
      }
    }
    const int n_tiles = cdiv(Cout, BN);
    const int Cout_pad = n_tiles * BN;

    ConvLaunch L;
    std::memset(&L, 0, sizeof(L));
    ConvRefParams R;
    std::memset(&R, 0, sizeof(R));
    L.BN = BN; L.BK = BK;
    L.grid = dim3(m_tiles, n_tiles, 1);
    ConvParams& p = L.p;
    p.nsrc = static_cast<int>(srcs.size());
    p.tiles_w = tiles_w; p.tiles_h = tiles_h;
    p.TW = TW; p.TH = TH; p.TN = TN;
    p.N = gN; p.Ho = gH; p.Wo = gW; p.Cout = Cout;
    p.relu = relu ? 1 : 0;
    p.has_res = res ? 1 : 0;
    p.out_mode = out_slot >= 0 ? kOutNCHWf32 : (split ? kOutNHWCsplit : kOutNHWCbf16);

    // ---- activation maps + tap tables
    int nmaps = 0;
    struct TapW { int r, s; };
    std::vector<std::vector<TapW>> kept(srcs.size());
    for (size_t si = 0; si < srcs.size(); ++si) {
      const ConvSrcSpec& s = srcs[si];
      const T& in = s.in;
      ConvSrc& cs = p.src[si];
      cs.chunks = cdiv(in.C, BK);
      cs.ntaps = 0;
      int map_of[2][2] = {{-1, -1}, {-1, -1}};
      struct Cand { int hp, wp, dh, dw, r, q; };
      std::vector<Cand> cands;
      if (!s.xtaps.empty()) {
        if (s.stride != 1) fail(name + ": explicit taps require stride 1");
        for (const TapSpec& t : s.xtaps) cands.push_back(Cand{0, 0, t.dh, t.dw, t.r, t.s});
      } else {
        for (int r = 0; r < s.k; ++r)
          for (int q = 0; q < s.k; ++q) {
            Cand c{0, 0, 0, 0, r, q};
            if (s.k == 3) {
              if (s.stride == 1) { c.dh = r - 1; c.dw = q - 1; }
              else { c.hp = (r == 1) ? 0 : 1; c.dh = (r == 0) ? -1 : 0; c.wp = (q == 1) ? 0 : 1; c.dw = (q == 0) ? -1 : 0; }
            }
            cands.push_back(c);
          }
      }
      for (const Cand& cd : cands) {
        {
          const int hp = cd.hp, wp = cd.wp, dh = cd.dh, dw = cd.dw, r = cd.r, q = cd.q;
          // geometry of the (sub-)lattice this tap reads
          const int st = s.stride;
          const int LW = (in.W - wp + st - 1) / st, LH = (in.H - hp + st - 1) / st;
          if (LW <= 0 || LH <= 0) continue;  // tap lies entirely in the padding
          if (map_of[hp][wp] < 0) {
            if (nmaps >= kConvMaxMaps) fail(name + ": too many tensor maps");
            const bf16* base = in.ptr + (static_cast<long>(hp) * in.W + wp) * in.ps;
            RefMap rm;
            rm.ptr = base; rm.C = in.C;
            if (flat) {
              rm.W = N * in.H * in.W; rm.H = 1; rm.N = 1;
              rm.sW = in.ps; rm.sH = 0; rm.sN = 0;
            } else {
              rm.W = LW; rm.H = LH; rm.N = N;
              rm.sW = static_cast<long>(st) * in.ps;
              rm.sH = static_cast<long>(st) * in.W * in.ps;
              rm.sN = static_cast<long>(in.H) * in.W * in.ps;
            }
            R.maps[nmaps] = rm;
            if (!dry) {
              const uint64_t big = static_cast<uint64_t>(rm.W) * rm.sW * 2;
              uint64_t dims[4] = {static_cast<uint64_t>(rm.C), static_cast<uint64_t>(rm.W),
                                  static_cast<uint64_t>(rm.H), static_cast<uint64_t>(rm.N)};
              uint64_t strides[3] = {static_cast<uint64_t>(rm.sW) * 2,
                                     flat ? big : static_cast<uint64_t>(rm.sH) * 2,
                                     flat ? big : static_cast<uint64_t>(rm.sN) * 2};
              uint32_t box[4] = {static_cast<uint32_t>(BK), static_cast<uint32_t>(TW), static_cast<uint32_t>(TH),
                                 static_cast<uint32_t>(TN)};
              p.tmA[nmaps] = encode_map(base, 4, dims, strides, box, BK * 2);
            }
            map_of[hp][wp] = nmaps++;
          }
          if (cs.ntaps >= kConvMaxTaps) fail(name + ": too many taps");
          cs.taps[cs.ntaps++] = static_cast<uint32_t>(map_of[hp][wp]) | (static_cast<uint32_t>(dh + 8) << 8) |
                                (static_cast<uint32_t>(dw + 8) << 16);
          kept[si].push_back(TapW{r, q});
        }
      }
      if (cs.ntaps == 0) fail(name + ": source has no taps");
    }

    // ---- packed weights [Cout_pad][Ktot] bf16, K order == the kernel's (src, tap, chunk, channel)
    long Ktot = 0;
    for (int si = 0; si < p.nsrc; ++si) Ktot += static_cast<long>(p.src[si].ntaps) * p.src[si].chunks * BK;
    std::vector<uint16_t> wpk;
    if (!dry) {
      wpk.assign(static_cast<size_t>(Cout_pad) * Ktot, 0);
      long kofs = 0;
      for (size_t si = 0; si < srcs.size(); ++si) {
        const ConvSrcSpec& s = srcs[si];
        const int Cin = s.in.C, kk = s.k * s.k, chunks = p.src[si].chunks;
        if (s.dev_w) {   // packed on the device every step (see pack_jobs below)
          kofs += static_cast<long>(kept[si].size()) * chunks * BK;
          continue;
        }
        for (size_t ti = 0; ti < kept[si].size(); ++ti) {
          const int r = kept[si][ti].r, q = kept[si][ti].s;
          for (int co = 0; co < Cout; ++co) {
            uint16_t* dst = wpk.data() + static_cast<size_t>(co) * Ktot + kofs + static_cast<long>(ti) * chunks * BK;
            const float* wsrc = s.w.data() + static_cast<size_t>(co) * Cin * kk + r * s.k + q;
            for (int ci = 0; ci < Cin; ++ci) dst[ci] = f2bf(wsrc[static_cast<size_t>(ci) * kk]);
          }
        }
        kofs += static_cast<long>(kept[si].size()) * chunks * BK;
      }
    }
    const bf16* wdev =
        reinterpret_cast<const bf16*>(alloc_wt(wpk.data(), static_cast<size_t>(Cout_pad) * Ktot * 2));
    {
      long kofs = 0;
      for (size_t si = 0; si < srcs.size(); ++si) {
        const ConvSrcSpec& s = srcs[si];
        const int chunks = p.src[si].chunks;
        if (s.dev_w) {
          PackJob j;
          std::memset(&j, 0, sizeof(j));
          j.src = s.dev_w;
          j.dst = const_cast<bf16*>(wdev);
          j.Ktot = Ktot; j.kofs = kofs; j.rows_pad = Cout_pad;
          j.k = s.k; j.BK = BK; j.chunks = chunks; j.ntaps = static_cast<int>(kept[si].size());
          j.dgrad = s.dgrad ? 1 : 0;
          j.Cin_total = s.dev_cin_total; j.ci_off = s.dev_ci_off;
          const int src_ch = s.dev_cin ? s.dev_cin : s.in.C;   // valid channels of the source tensor
          j.Cout = s.dgrad ? src_ch : Cout;   // ORIGINAL conv's Cout / Cin (dgrad swaps the GEMM roles)
          j.Cin = s.dgrad ? Cout : src_ch;
          for (size_t ti = 0; ti < kept[si].size(); ++ti) {
            int r = kept[si][ti].r, q = kept[si][ti].s;
            if (s.dgrad && s.xtaps.empty()) { r = s.k - 1 - r; q = s.k - 1 - q; }   // mirrored filter
            j.taps[ti] = static_cast<unsigned char>((r << 4) | q);
          }
          if (!dry) pack_jobs.push_back(j);
        }
        kofs += static_cast<long>(kept[si].size()) * chunks * BK;
      }
    }
    std::vector<float> bpad(Cout_pad, 0.f);
    for (int c = 0; c < Cout && c < static_cast<int>(bias.size()); ++c) bpad[c] = bias[c];
    const float* bdev = dev_bias_override ? dev_bias_override : upload_f32(bpad);
    p.bias = bdev;

    if (!dry) {
      uint64_t dims[2] = {static_cast<uint64_t>(Ktot), static_cast<uint64_t>(Cout_pad)};
      uint64_t strides[1] = {static_cast<uint64_t>(Ktot) * 2};
      uint32_t box[2] = {static_cast<uint32_t>(BK), static_cast<uint32_t>(BN)};
      p.tmB = encode_map(wdev, 2, dims, strides, box, BK * 2);
      const int SC = BN < 64 ? BN : 64;
      auto out_map = [&](const T& t) {
        uint64_t dims4[4], str[3];
        dims4[0] = static_cast<uint64_t>(t.C);
        if (flat) {
          const uint64_t rows = static_cast<uint64_t>(N) * Ho * Wo;
          dims4[1] = rows; dims4[2] = 1; dims4[3] = 1;
          str[0] = static_cast<uint64_t>(t.ps) * 2; str[1] = rows * t.ps * 2; str[2] = rows * t.ps * 2;
        } else {
          dims4[1] = Wo; dims4[2] = Ho; dims4[3] = N;
          str[0] = static_cast<uint64_t>(t.ps) * 2;
          str[1] = static_cast<uint64_t>(t.row_stride()) * 2;
          str[2] = static_cast<uint64_t>(t.img_stride()) * 2;
        }
        uint32_t box4[4] = {static_cast<uint32_t>(SC), static_cast<uint32_t>(TW), static_cast<uint32_t>(TH),
                            static_cast<uint32_t>(TN)};
        return encode_map(t.ptr, 4, dims4, str, box4, SC * 2);
      };
      if (out_slot < 0) p.tmD = out_map(out);
      if (res) p.tmR = out_map(*res);
    }
    Conv3Launch L3;
    std::memset(&L3, 0, sizeof(L3));
    if (ws) {
      Conv3Params& q = L3.p;
      L3.BN = BN; L3.CK = BK; L3.mode = ws_mode; L3.smem_bytes = ws_smem; L3.pair = ws_pair;
      q.nstage = ws_nstage;
      q.chunks = ws_chunks;
      q.chunks0 = cdiv(in0.C, BK);
      q.npatch = ws_np;
      q.N = N; q.Cout = Cout; q.Ho = Ho; q.Wo = Wo;
      q.relu = p.relu; q.has_res = p.has_res; q.out_mode = p.out_mode;
      q.bias = bdev;
      if (stats && !res && out_slot < 0 && !relu && conv_impl == 0 && ((stats_kinds >> ws_mode) & 1)) {   // raw conv output -> BatchNorm: statistics in the store warp
        q.stats = stats; q.stats_C = Cout;
        stats_taken = true;
      }
      const long rows = static_cast<long>(N) * Ho * Wo;
      long mt;
      if (ws_mode == 0) {
        q.tiles_w = cdiv(Wo, 8); q.tiles_h = cdiv(Ho, 16);
        mt = static_cast<long>(N) * q.tiles_w * q.tiles_h;
      } else {
        mt = (rows + 127) / 128;
        q.tiles_w = static_cast<int>(mt); q.tiles_h = 1;
      }
      long gx = std::max<long>(1, num_sms / n_tiles);
      if (gx > mt) gx = mt;
      if (ws_pair) gx = std::max<long>(2, gx & ~1L);   // whole pairs (the last pair's second CTA may get a tile past the end)
      L3.grid = dim3(static_cast<unsigned>(gx), n_tiles, 1);
      if (!dry) {
        auto nhwc_map = [&](const T& t, int boxc, int bw, int bh) {
          uint64_t dims4[4] = {static_cast<uint64_t>(t.C), static_cast<uint64_t>(t.W), static_cast<uint64_t>(t.H),
                               static_cast<uint64_t>(t.N)};
          uint64_t str[3] = {static_cast<uint64_t>(t.ps) * 2, static_cast<uint64_t>(t.W) * t.ps * 2,
                             static_cast<uint64_t>(t.H) * t.W * t.ps * 2};
          uint32_t box4[4] = {static_cast<uint32_t>(boxc), static_cast<uint32_t>(bw), static_cast<uint32_t>(bh), 1};
          return encode_map(t.ptr, 4, dims4, str, box4, boxc * 2);
        };
        auto flat_map = [&](const T& t, int boxc) {
          const uint64_t r = static_cast<uint64_t>(t.N) * t.H * t.W;
          uint64_t dims4[4] = {static_cast<uint64_t>(t.C), r, 1, 1};
          uint64_t str[3] = {static_cast<uint64_t>(t.ps) * 2, r * t.ps * 2, r * t.ps * 2};
          uint32_t box4[4] = {static_cast<uint32_t>(boxc), 128, 1, 1};
          return encode_map(t.ptr, 4, dims4, str, box4, boxc * 2);
        };
        const int SC = BN < 64 ? BN : 64;
        q.tmW = p.tmB;
        if (ws_pair) {   // each CTA of a pair loads its half of the Cout tile's rows
          uint64_t wd[2] = {static_cast<uint64_t>(Ktot), static_cast<uint64_t>(Cout_pad)};
          uint64_t wst[1] = {static_cast<uint64_t>(Ktot) * 2};
          uint32_t wbox[2] = {static_cast<uint32_t>(BK), static_cast<uint32_t>(BN / 2)};
          q.tmW = encode_map(wdev, 2, wd, wst, wbox, BK * 2);
        }
        if (ws_mode == 0) {
          q.tmA = nhwc_map(in0, BK, 10, 18);
          if (out_slot < 0) q.tmD = nhwc_map(out, SC, 8, 16);
          if (res) q.tmR = nhwc_map(*res, SC, 8, 16);
        } else {
          q.tmA = flat_map(in0, BK);
          if (srcs.size() > 1) q.tmA2 = flat_map(srcs[1].in, BK);
          if (out_slot < 0) q.tmD = flat_map(out, SC);
          if (res) q.tmR = flat_map(*res, SC);
        }
      }
    }
    // ---- SIMT restatement parameters
    for (int si = 0; si < p.nsrc; ++si) R.src[si] = p.src[si];
    R.nsrc = p.nsrc; R.BK = BK; R.wpk = wdev; R.Ktot = Ktot; R.bias = bdev;
    R.N = gN; R.Ho = gH; R.Wo = gW; R.Cout = Cout; R.relu = p.relu; R.out_mode = p.out_mode;
    if (res) {
      R.res.ptr = res->ptr;
      R.res.sW = res->ps; R.res.sH = flat ? 0 : res->row_stride();
      R.res.sN = flat ? 0 : res->img_stride();
    }
    if (out_slot < 0) {
      R.out = out.ptr;
      R.o_sW = out.ps; R.o_sH = flat ? 0 : out.row_stride();
      R.o_sN = flat ? 0 : out.img_stride();
    }

    std::vector<const T*> ins;
    for (const auto& s : srcs) ins.push_back(&s.in);
    if (res) ins.push_back(res);
    const int impl = conv_impl;
    const int idx = add_op(name, ins, [L, L3, ws, R, impl, out_slot](cudaStream_t st, const RunArgs& a) mutable {
      if (impl == 0 && ws) {
        if (out_slot >= 0) L3.p.out_f32 = a.out[out_slot];
        return conv3_ws_launch(L3, st);
      }
      if (impl == 0) {
        if (out_slot >= 0) L.p.out_f32 = a.out[out_slot];
        return conv_tc_launch(L, st);
      }
      if (out_slot >= 0) R.out_f32 = a.out[out_slot];
      return conv_ref_launch(R, st);
    });
    {
      Op& op = ops[idx];
      char lab[64];
      if (impl == 0 && ws) std::snprintf(lab, sizeof(lab), "%s<BN=%d,CK=%d>", ws_mode == 0 ? (ws_pair ? "conv3_ws_pair" : "conv3_ws") : "conv1_ws", BN, BK);
      else if (impl == 0) std::snprintf(lab, sizeof(lab), "conv_tc<BN=%d,BK=%d>", BN, BK);
      else std::snprintf(lab, sizeof(lab), "conv_ref");
      op.kernel = lab;
      for (const auto& s : srcs) {
        op.flops += 2.0 * N * Ho * Wo * Cout * s.in.C * s.k * s.k;
        op.bytes += 2.0 * s.in.N * s.in.H * s.in.W * s.in.C + 2.0 * Cout * s.in.C * s.k * s.k;
      }
      if (res) op.bytes += 2.0 * N * Ho * Wo * Cout;
      op.bytes += (out_slot >= 0 ? 4.0 : 2.0) * N * Ho * Wo * Cout;
    }
    out.prod = idx;
    return out;
  }
  static double tbytes(const T& t) { return 2.0 * t.N * t.H * t.W * t.C; }
  void label(int idx, const char* kernel, double bytes, double flops = 0) {
    ops[idx].kernel = kernel;
    ops[idx].bytes = bytes;
    ops[idx].flops = flops;
  }
};

// ----------------------------------------------------------------------------------------- engine
struct Engine {
  pidnet_cfg cfg{};
  std::map<std::string, HostParam> params;
  Builder b;
  bool planned = false;
  int N = 0, H = 0, W = 0;
  int lanes = 3;
  int conv_impl = 0;
  int use_ws = 1;
  int use_pair = -1, ws_stages = -1, use_stem2 = -1, use_pyramid = -1;   // -1: builder default (environment / built-in)
  int fp32_head = 0;   // final_layer (the logits path) in split-bf16 arithmetic: weights and the hidden tensor as hi + lo pairs
  cudaStream_t side[2] = {nullptr, nullptr};
  cudaStream_t cap_stream = nullptr;  // capture origin (the caller's stream may be the legacy default stream)
  // uint8 input path: per-channel table of the reference's input_transform (datasets/base_dataset.py:36-44), evaluated
  // with numpy's arithmetic: `image / 255.0` is float32; `image -= mean` and `image /= std` (python-float lists) run in
  // float64 and are rounded back to float32 after each statement
  float* dev_lut = nullptr;
  std::vector<float> host_lut;
  double lut_mean[3] = {0, 0, 0}, lut_std[3] = {0, 0, 0};
  const float* u8_lut(const double* mean, const double* stdv, cudaStream_t st) {
    for (int c = 0; c < 3; ++c)
      if (!(stdv[c] > 0.0)) fail("pidnet_forward_u8: std must be positive");
    const bool same = dev_lut && std::memcmp(mean, lut_mean, sizeof(lut_mean)) == 0 && std::memcmp(stdv, lut_std, sizeof(lut_std)) == 0;
    if (same) return dev_lut;
    if (!dev_lut) CK(cudaMalloc(&dev_lut, 3 * 256 * sizeof(float)));
    CK(cudaStreamSynchronize(st));   // the previous table may still be in use (rare path: the constants changed)
    host_lut.resize(3 * 256);
    for (int c = 0; c < 3; ++c)
      for (int b = 0; b < 256; ++b) {
        const float a = static_cast<float>(b) / 255.0f;
        const float m = static_cast<float>(static_cast<double>(a) - mean[c]);
        host_lut[c * 256 + b] = static_cast<float>(static_cast<double>(m) / stdv[c]);
      }
    CK(cudaMemcpy(dev_lut, host_lut.data(), host_lut.size() * sizeof(float), cudaMemcpyHostToDevice));
    std::memcpy(lut_mean, mean, sizeof(lut_mean));
    std::memcpy(lut_std, stdv, sizeof(lut_std));
    return dev_lut;
  }
  std::vector<cudaEvent_t> events;  // one per op that records
  std::vector<int> ev_of_op;
  cudaEvent_t ev_start = nullptr, ev_join[2] = {nullptr, nullptr};
  // CUDA-graph cache keyed by the runtime pointers (a few entries: double-buffered callers)
  struct GraphEntry {
    RunArgs args;
    cudaGraphExec_t exec;
  };
  std::vector<GraphEntry> graphs;
  static constexpr size_t kMaxGraphs = 8;   // (input, output) pointer sets of double-buffered callers whose outputs come from a caching allocator

  ~Engine() { release(); }
  void release() {
    for (auto& g : graphs) cudaGraphExecDestroy(g.exec);
    graphs.clear();
    for (auto e : events) cudaEventDestroy(e);
    events.clear();
    if (ev_start) { cudaEventDestroy(ev_start); ev_start = nullptr; }
    for (int i = 0; i < 2; ++i) {
      if (ev_join[i]) { cudaEventDestroy(ev_join[i]); ev_join[i] = nullptr; }
      if (side[i]) { cudaStreamDestroy(side[i]); side[i] = nullptr; }
    }
    if (cap_stream) { cudaStreamDestroy(cap_stream); cap_stream = nullptr; }
    if (b.act_base) { cudaFree(b.act_base); b.act_base = nullptr; }
    if (b.wt_base) { cudaFree(b.wt_base); b.wt_base = nullptr; }
    if (dev_lut) { cudaFree(dev_lut); dev_lut = nullptr; }
    planned = false;
  }

  // ---- parameters
  const HostParam& P(const std::string& k) const {
    auto it = params.find(k);
    if (it == params.end()) fail("missing parameter '" + k + "'");
    return it->second;
  }
  bool has(const std::string& k) const { return params.count(k) != 0; }
  Affine bn(const std::string& p) const {  // eval BatchNorm as y = s*x + t  (eps 1e-5)
    const auto &g = P(p + ".weight").data, &be = P(p + ".bias").data, &mu = P(p + ".running_mean").data,
               &var = P(p + ".running_var").data;
    Affine a;
    a.s.resize(g.size());
    a.t.resize(g.size());
    for (size_t i = 0; i < g.size(); ++i) {
      const double s = static_cast<double>(g[i]) / std::sqrt(static_cast<double>(var[i]) + 1e-5);
      a.s[i] = static_cast<float>(s);
      a.t[i] = static_cast<float>(static_cast<double>(be[i]) - static_cast<double>(mu[i]) * s);
    }
    return a;
  }
  static Affine slice(const Affine& a, int off, int n) {
    Affine r;
    r.s.assign(a.s.begin() + off, a.s.begin() + off + n);
    r.t.assign(a.t.begin() + off, a.t.begin() + off + n);
    return r;
  }
  // conv weight scaled per output channel by `post` (the BN that FOLLOWS the conv); bias folded likewise
  void fold(const std::string& conv, const Affine* post, std::vector<float>& w, std::vector<float>& bias) const {
    const HostParam& hw = P(conv + ".weight");
    const int Cout = static_cast<int>(hw.shape[0]);
    const size_t per = hw.data.size() / Cout;
    w = hw.data;
    bias.assign(Cout, 0.f);
    if (has(conv + ".bias")) bias = P(conv + ".bias").data;
    if (post) {
      for (int co = 0; co < Cout; ++co) {
        for (size_t i = 0; i < per; ++i) w[co * per + i] *= post->s[co];
        bias[co] = bias[co] * post->s[co] + post->t[co];
      }
    }
  }
  ConvSrcSpec src_of(const T& in, const std::string& conv, const Affine* post, int k, int stride,
                     std::vector<float>* bias_acc) const {
    ConvSrcSpec s;
    s.in = in; s.k = k; s.stride = stride;
    std::vector<float> bias;
    fold(conv, post, s.w, bias);
    const HostParam& hw = P(conv + ".weight");
    if (hw.shape.size() != 4 || hw.shape[1] != in.C || hw.shape[2] != k)
      fail("weight '" + conv + "' does not match its input (" + std::to_string(in.C) + " ch)");
    if (bias_acc) {
      if (bias_acc->empty()) *bias_acc = bias;
      else for (size_t i = 0; i < bias.size(); ++i) (*bias_acc)[i] += bias[i];
    }
    return s;
  }

  // ---- blocks (fused forms of model_utils.py BasicBlock / Bottleneck / segmenthead)
  T conv_bn(const std::string& name, const T& x, const std::string& conv, const std::string& bnp, int k, int stride,
            bool relu) {
    Affine a = bn(bnp);
    std::vector<float> bias;
    ConvSrcSpec s = src_of(x, conv, &a, k, stride, &bias);
    return b.conv(name, {s}, bias, static_cast<int>(P(conv + ".weight").shape[0]), relu, nullptr, nullptr, -1);
  }
  // tail conv of a residual block: conv+bn (+ 1x1 downsample conv+bn as extra K slices | + identity residual)
  T block_tail(const std::string& p, const T& mid, const std::string& conv, const std::string& bnp, int k, const T& x,
               int stride, bool relu) {
    Affine a = bn(p + "." + bnp);
    std::vector<float> bias;
    std::vector<ConvSrcSpec> srcs;
    srcs.push_back(src_of(mid, p + "." + conv, &a, k, 1, &bias));
    const int Cout = static_cast<int>(P(p + "." + conv + ".weight").shape[0]);
    if (has(p + ".downsample.0.weight")) {
      if (k == 3 && Cout <= 64 && b.ws3_eligible(mid.C, Cout)) {   // (wider tails re-read the patch per Cout tile: not worth it)
        // the 3x3 tail can run on the weight-stationary halo kernel, which takes ONE source: compute the 1x1 downsample
        // branch on its own and feed it as the residual tile (measured: 170 -> ~110 us for layer3_d, 193 -> ~135 us for
        // layer2.0 at batch 32), instead of K-concatenating it into the generic tap-by-tap kernel
        T r = conv_bn(p + ".downsample", x, p + ".downsample.0", p + ".downsample.1", 1, stride, false);
        return b.conv(p + "." + conv + "+dsres", srcs, bias, Cout, relu, &r, nullptr, -1);
      }
      Affine d = bn(p + ".downsample.1");
      srcs.push_back(src_of(x, p + ".downsample.0", &d, 1, stride, &bias));
      return b.conv(p + "." + conv + "+ds", srcs, bias, Cout, relu, nullptr, nullptr, -1);
    }
    return b.conv(p + "." + conv + "+res", srcs, bias, Cout, relu, &x, nullptr, -1);
  }
  T basic_block(const std::string& p, const T& x, int stride, bool relu_out) {
    T mid = conv_bn(p + ".conv1", x, p + ".conv1", p + ".bn1", 3, stride, true);
    return block_tail(p, mid, "conv2", "bn2", 3, x, stride, relu_out);
  }
  T bottleneck(const std::string& p, const T& x, int stride, bool relu_out) {
    T a = conv_bn(p + ".conv1", x, p + ".conv1", p + ".bn1", 1, 1, true);
    T c = conv_bn(p + ".conv2", a, p + ".conv2", p + ".bn2", 3, stride, true);
    return block_tail(p, c, "conv3", "bn3", 1, x, stride, relu_out);
  }
  // _make_layer (pidnet.py:103-121); `relu_last`: what the single stored form of the layer output must be
  T layer(const std::string& p, T x, bool bott, int blocks, int stride, bool relu_last) {
    for (int i = 0; i < blocks; ++i) {
      const bool last = i == blocks - 1;
      // block 0 keeps the block default (BasicBlock: relu, Bottleneck: no relu); middle blocks relu
      bool relu = last ? relu_last : (i == 0 ? !bott : true);
      if (blocks == 1) relu = relu_last;
      const std::string bp = p + "." + std::to_string(i);
      x = bott ? bottleneck(bp, x, i == 0 ? stride : 1, relu) : basic_block(bp, x, i == 0 ? stride : 1, relu);
    }
    return x;
  }
  T affine_relu(const std::string& name, const T& x, const Affine& a, const T* out_view) {
    T out = out_view ? *out_view : b.new_tensor(x.N, x.H, x.W, x.C);
    const float* s = b.upload_f32(a.s);
    const float* t = b.upload_f32(a.t);
    View xv = x.view(), ov = out.view();
    out.prod = b.add_op(name, {&x}, [xv, ov, s, t](cudaStream_t st, const RunArgs&) {
      return upadd_launch(xv, View{nullptr, 0, 0, 0, 0, 0}, ov, s, t, 1, st);
    });
    b.label(out.prod, "upadd", Builder::tbytes(x) + Builder::tbytes(out));
    return out;
  }
  // out = act(s*(a + U(blow)) + t)
  T upadd(const std::string& name, const T& a, const T& blow, const Affine* aff, bool relu, const T* out_view) {
    T out = out_view ? *out_view : b.new_tensor(a.N, a.H, a.W, a.C);
    const float* s = aff ? b.upload_f32(aff->s) : nullptr;
    const float* t = aff ? b.upload_f32(aff->t) : nullptr;
    View av = a.view(), bv = blow.view(), ov = out.view();
    const int r = relu ? 1 : 0;
    out.prod = b.add_op(name, {&a, &blow}, [av, bv, ov, s, t, r](cudaStream_t st, const RunArgs&) {
      return upadd_launch(av, bv, ov, s, t, r, st);
    });
    b.label(out.prod, "upadd", Builder::tbytes(a) + Builder::tbytes(blow) + Builder::tbytes(out));
    return out;
  }
  // segmenthead in split-bf16 arithmetic ("fp32_head"): every fp32 weight w is used as w_hi + w_lo (two bf16 values, ~16
  // mantissa bits) and the hidden tensor h = relu(bn2(conv1(x))) is stored as h_hi + h_lo, so that
  //   conv1:  y = x * w_hi + x * w_lo                              (x is the bf16 tensor the network produced: exact)
  //   conv2:  z = h_hi * w_hi + h_lo * w_hi + h_hi * w_lo (+ bias)  (the dropped h_lo * w_lo term is ~2^-16 relative)
  // are plain K-concatenated tcgen05 GEMMs with fp32 accumulation: given its input, the head matches fp32 arithmetic to ~1e-5
  // (the bf16 head: ~3e-3).  Costs 2x / 3x the MMAs of the two head convs.
  void seghead_tail_split(const std::string& p, const T& xin, int slot) {
    auto lo_of = [](const std::vector<float>& w) {
      std::vector<float> r(w.size());
      for (size_t i = 0; i < w.size(); ++i) r[i] = w[i] - bf2f(f2bf(w[i]));
      return r;
    };
    const int Ch = static_cast<int>(P(p + ".conv1.weight").shape[0]);
    T hl;
    {
      Affine a = bn(p + ".bn2");
      std::vector<float> bias;
      ConvSrcSpec hi = src_of(xin, p + ".conv1", &a, 3, 1, &bias);
      ConvSrcSpec lo = hi;
      lo.w = lo_of(hi.w);
      b.split_next = true;
      hl = b.conv(p + ".conv1[hi+lo]", {hi, lo}, bias, Ch, true, nullptr, nullptr, -1);
    }
    {
      std::vector<float> w2, bias2;
      fold(p + ".conv2", nullptr, w2, bias2);                 // [ncls][Ch]
      const int ncls = static_cast<int>(P(p + ".conv2.weight").shape[0]);
      ConvSrcSpec s0, s1;
      s0.in = hl; s0.k = 1; s0.stride = 1;
      s0.w.resize(static_cast<size_t>(ncls) * 2 * Ch);
      for (int co = 0; co < ncls; ++co)
        for (int ci = 0; ci < Ch; ++ci)
          s0.w[static_cast<size_t>(co) * 2 * Ch + ci] = s0.w[static_cast<size_t>(co) * 2 * Ch + Ch + ci] = w2[static_cast<size_t>(co) * Ch + ci];
      s1.in = Builder::slice(hl, 0, Ch); s1.k = 1; s1.stride = 1;
      s1.w = lo_of(w2);
      b.conv(p + ".conv2[hi+lo]", {s0, s1}, bias2, ncls, false, nullptr, nullptr, slot);
    }
  }

  // segmenthead (model_utils.py:100-112): `xin` must already hold relu(bn1(x)); writes fp32 NCHW slot
  void seghead_tail(const std::string& p, const T& xin, int slot) {
    if (fp32_head && slot == 0 && conv_impl == 0) return seghead_tail_split(p, xin, slot);
    T h = conv_bn(p + ".conv1", xin, p + ".conv1", p + ".bn2", 3, 1, true);
    std::vector<float> bias;
    ConvSrcSpec s = src_of(h, p + ".conv2", nullptr, 1, 1, &bias);
    b.conv(p + ".conv2", {s}, bias, static_cast<int>(P(p + ".conv2.weight").shape[0]), false, nullptr, nullptr, slot);
  }

  // PagFM (model_utils.py:292-312) with compressionK folded in: one low-res 1x1 conv from the I-branch
  // tensor produces [y | z | t] (see DESIGN.md "PagFM algebra"), then the fuse kernel.
  T pag(const std::string& pg, const std::string& comp, const T& xhi, const T& ilow) {
    const int C2 = xhi.C, Ci = ilow.C, Pm = static_cast<int>(P(pg + ".f_x.0.weight").shape[0]);
    const Affine ac = bn(comp + ".1"), ax = bn(pg + ".f_x.1"), ay = bn(pg + ".f_y.1");
    const auto& Wc = P(comp + ".0.weight").data;     // [C2][Ci]
    const auto& Wx = P(pg + ".f_x.0.weight").data;   // [Pm][C2]
    const auto& Wy = P(pg + ".f_y.0.weight").data;   // [Pm][C2]
    // y = Yw I + yb
    std::vector<double> Yw(static_cast<size_t>(C2) * Ci), yb(C2);
    for (int c = 0; c < C2; ++c) {
      for (int i = 0; i < Ci; ++i) Yw[static_cast<size_t>(c) * Ci + i] = static_cast<double>(ac.s[c]) * Wc[static_cast<size_t>(c) * Ci + i];
      yb[c] = ac.t[c];
    }
    // G = Ax^T Ay  (C2 x C2), g0 = Ax^T ty ; r = tx^T Ay (1 x C2), r0 = tx . ty
    std::vector<double> G(static_cast<size_t>(C2) * C2, 0.0), g0(C2, 0.0), r(C2, 0.0);
    double r0 = 0.0;
    for (int m = 0; m < Pm; ++m) {
      for (int a = 0; a < C2; ++a) {
        const double axv = static_cast<double>(ax.s[m]) * Wx[static_cast<size_t>(m) * C2 + a];
        for (int c = 0; c < C2; ++c) G[static_cast<size_t>(a) * C2 + c] += axv * (static_cast<double>(ay.s[m]) * Wy[static_cast<size_t>(m) * C2 + c]);
        g0[a] += axv * ay.t[m];
      }
      for (int c = 0; c < C2; ++c) r[c] += static_cast<double>(ax.t[m]) * (static_cast<double>(ay.s[m]) * Wy[static_cast<size_t>(m) * C2 + c]);
      r0 += static_cast<double>(ax.t[m]) * ay.t[m];
    }
    const int CL = 2 * C2 + 8;
    ConvSrcSpec s;
    s.in = ilow; s.k = 1; s.stride = 1;
    s.w.assign(static_cast<size_t>(CL) * Ci, 0.f);
    std::vector<float> bias(CL, 0.f);
    for (int c = 0; c < C2; ++c) {
      for (int i = 0; i < Ci; ++i) s.w[static_cast<size_t>(c) * Ci + i] = static_cast<float>(Yw[static_cast<size_t>(c) * Ci + i]);
      bias[c] = static_cast<float>(yb[c]);
    }
    for (int a = 0; a < C2; ++a) {  // z = G y + g0
      double bz = g0[a];
      for (int c = 0; c < C2; ++c) bz += G[static_cast<size_t>(a) * C2 + c] * yb[c];
      bias[C2 + a] = static_cast<float>(bz);
      for (int i = 0; i < Ci; ++i) {
        double acc = 0.0;
        for (int c = 0; c < C2; ++c) acc += G[static_cast<size_t>(a) * C2 + c] * Yw[static_cast<size_t>(c) * Ci + i];
        s.w[static_cast<size_t>(C2 + a) * Ci + i] = static_cast<float>(acc);
      }
    }
    {  // t = r y + r0
      double bt = r0;
      for (int c = 0; c < C2; ++c) bt += r[c] * yb[c];
      bias[2 * C2] = static_cast<float>(bt);
      for (int i = 0; i < Ci; ++i) {
        double acc = 0.0;
        for (int c = 0; c < C2; ++c) acc += r[c] * Yw[static_cast<size_t>(c) * Ci + i];
        s.w[static_cast<size_t>(2 * C2) * Ci + i] = static_cast<float>(acc);
      }
    }
    T low = b.conv(pg + ".low", {s}, bias, CL, false, nullptr, nullptr, -1);
    T out = b.new_tensor(xhi.N, xhi.H, xhi.W, C2);
    View xv = xhi.view(), lv = low.view(), ov = out.view();
    out.prod = b.add_op(pg + ".fuse", {&xhi, &low}, [xv, lv, ov](cudaStream_t st, const RunArgs&) {
      return pag_fuse_launch(xv, lv, ov, 1, st);
    });
    b.label(out.prod, "pag_fuse", Builder::tbytes(xhi) + Builder::tbytes(low) + Builder::tbytes(out));
    return out;
  }

  T pool_affine(const std::string& name, const T& x, int k, int stride, int pad, const Affine& a) {
    const int oh = k == 0 ? 1 : (x.H + 2 * pad - k) / stride + 1;
    const int ow = k == 0 ? 1 : (x.W + 2 * pad - k) / stride + 1;
    if (oh < 1 || ow < 1) fail(name + ": pooled map is empty");
    T out = b.new_tensor(x.N, oh, ow, x.C);
    const float* s = b.upload_f32(a.s);
    const float* t = b.upload_f32(a.t);
    View xv = x.view(), ov = out.view();
    out.prod = b.add_op(name, {&x}, [xv, ov, k, stride, pad, s, t](cudaStream_t st, const RunArgs&) {
      return pool_affine_launch(xv, ov, k, stride, pad, s, t, 1, st);
    });
    b.label(out.prod, "pool_affine", Builder::tbytes(x) + Builder::tbytes(out));
    return out;
  }
  T conv_plain(const std::string& name, const T& x, const std::string& conv, const Affine* post, int k, bool relu,
               const T* out_view) {
    std::vector<float> bias;
    ConvSrcSpec s = src_of(x, conv, post, k, 1, &bias);
    return b.conv(name, {s}, bias, static_cast<int>(P(conv + ".weight").shape[0]), relu, nullptr, out_view, -1);
  }

  // PAPPM (model_utils.py:247-265) / DAPPM (:174-194).  Pre-activation BNs are applied by the producers.
  T spp(const T& x) {
    const bool large = cfg.m == 3;
    const int ppm = cfg.ppm_planes, outp = cfg.planes * 4;
    static const int pk[3] = {5, 9, 17}, pstr[3] = {2, 4, 8}, pp[3] = {2, 4, 8};
    // scale branches: BN -> ReLU (after the pool) then 1x1
    const bool pyramid = b.use_pyramid && x.C % 32 == 0 && static_cast<size_t>(x.H + 1) * (x.W + 1) * 128 <= 200 * 1024;
    T a0, sc_in;   // relu(bn(x)) operands of scale0 and of the shortcut: by-products of the pyramid kernel when it is used
    if (pyramid) {
      a0 = b.new_tensor(x.N, x.H, x.W, x.C);
      sc_in = b.new_tensor(x.N, x.H, x.W, x.C);
    } else {
      a0 = affine_relu("spp.scale0.bnrelu", x, bn("spp.scale0.0"), nullptr);
    }
    T ak[4];
    if (pyramid) {
      // all four pooled branches (+ their BN + ReLU) from one summed-area table per 32-channel slice, plus the two
      // full-resolution relu(bn(x)) operands: ONE launch reads x once for six consumers
      std::vector<float> sall, tall, se, te;
      View ov[4];
      for (int k = 0; k < 4; ++k) {
        const int oh = k < 3 ? (x.H + 2 * pp[k] - pk[k]) / pstr[k] + 1 : 1, ow = k < 3 ? (x.W + 2 * pp[k] - pk[k]) / pstr[k] + 1 : 1;
        if (oh < 1 || ow < 1) fail("spp: pooled map is empty");
        ak[k] = b.new_tensor(x.N, oh, ow, x.C);
        ov[k] = ak[k].view();
        const Affine a = bn("spp.scale" + std::to_string(k + 1) + ".1");
        sall.insert(sall.end(), a.s.begin(), a.s.end());
        tall.insert(tall.end(), a.t.begin(), a.t.end());
      }
      for (const char* key : {"spp.scale0.0", "spp.shortcut.0"}) {
        const Affine a = bn(key);
        se.insert(se.end(), a.s.begin(), a.s.end());
        te.insert(te.end(), a.t.begin(), a.t.end());
      }
      const float* sd = b.upload_f32(sall);
      const float* td = b.upload_f32(tall);
      const float* sed = b.upload_f32(se);
      const float* ted = b.upload_f32(te);
      const View xv = x.view();
      const View o0 = ov[0], o1 = ov[1], o2 = ov[2], o3 = ov[3], e0 = a0.view(), e1 = sc_in.view();
      const int op = b.add_op("spp.pools+bnrelu", {&x}, [xv, o0, o1, o2, o3, e0, e1, sd, td, sed, ted](cudaStream_t st, const RunArgs&) {
        const View outs[4] = {o0, o1, o2, o3};
        const View ews[2] = {e0, e1};
        return pool_pyramid_launch(xv, outs, sd, td, ews, sed, ted, st);
      });
      double ob = Builder::tbytes(a0) + Builder::tbytes(sc_in);
      for (int k = 0; k < 4; ++k) { ak[k].prod = op; ob += Builder::tbytes(ak[k]); }
      a0.prod = op; sc_in.prod = op;
      b.label(op, "pool_pyramid", Builder::tbytes(x) + ob);
    } else {
      for (int k = 0; k < 4; ++k) {
        const std::string sp = "spp.scale" + std::to_string(k + 1);
        ak[k] = k < 3 ? pool_affine(sp + ".pool", x, pk[k], pstr[k], pp[k], bn(sp + ".1"))
                      : pool_affine(sp + ".pool", x, 0, 1, 0, bn(sp + ".1"));
      }
    }
    T s0 = conv_plain("spp.scale0.conv", a0, "spp.scale0.2", nullptr, 1, false, nullptr);
    T sk[4];

    for (int k = 0; k < 4; ++k) {
      const std::string sp = "spp.scale" + std::to_string(k + 1);
      sk[k] = conv_plain(sp + ".conv", ak[k], sp + ".3", nullptr, 1, false, nullptr);
    }
    const Affine acomp = bn("spp.compression.0");
    T comp_in = b.new_tensor(x.N, x.H, x.W, 5 * ppm);
    if (!large) {
      const Affine asp = bn("spp.scale_process.0");
      T sp_in = b.new_tensor(x.N, x.H, x.W, 4 * ppm);
      {   // relu(bn_k(scale0 + U(scale_k))) for the four branches and relu(bn_0(scale0)) (the first slice of the compression
          // input) in ONE launch; job 4 is "null + U(scale0)" at scale 1, an exact identity interpolation
        std::vector<View> av(5), bv(5), ovs(5);
        std::vector<const float*> sd(5), td(5);
        double bytes = 0;
        for (int k = 0; k < 4; ++k) {
          T dst = Builder::slice(sp_in, k * ppm, ppm);
          Affine a = slice(asp, k * ppm, ppm);
          av[k] = s0.view(); bv[k] = sk[k].view(); ovs[k] = dst.view();
          sd[k] = b.upload_f32(a.s); td[k] = b.upload_f32(a.t);
          bytes += Builder::tbytes(s0) + Builder::tbytes(sk[k]) + Builder::tbytes(dst);
        }
        T c0 = Builder::slice(comp_in, 0, ppm);
        {
          Affine a = slice(acomp, 0, ppm);
          av[4] = View{nullptr, 0, 0, 0, 0, 0}; bv[4] = s0.view(); ovs[4] = c0.view();
          sd[4] = b.upload_f32(a.s); td[4] = b.upload_f32(a.t);
          bytes += Builder::tbytes(s0) + Builder::tbytes(c0);
        }
        sp_in.prod = b.add_op("spp.scale1-4.upadd+x_.bnrelu", {&s0, &sk[0], &sk[1], &sk[2], &sk[3]},
                              [av, bv, ovs, sd, td](cudaStream_t st, const RunArgs&) {
                                return upadd_batch_launch(5, av.data(), bv.data(), ovs.data(), sd.data(), td.data(), 1, st);
                              });
        comp_in.prod = sp_in.prod;
        b.label(sp_in.prod, "upadd", bytes);
      }
      // grouped 3x3 (groups=4, model_utils.py:230) as four convs on channel slices; epilogue applies the
      // compression BN slice + ReLU
      const HostParam& gw = P("spp.scale_process.2.weight");  // [4*ppm][ppm][3][3]
      for (int g = 0; g < 4; ++g) {
        ConvSrcSpec s;
        s.in = Builder::slice(sp_in, g * ppm, ppm);
        s.k = 3; s.stride = 1;
        const size_t per = static_cast<size_t>(ppm) * 9;
        s.w.assign(gw.data.begin() + static_cast<size_t>(g) * ppm * per, gw.data.begin() + static_cast<size_t>(g + 1) * ppm * per);
        std::vector<float> bias(ppm);
        for (int co = 0; co < ppm; ++co) {
          const float sc = acomp.s[ppm + g * ppm + co];
          for (size_t i = 0; i < per; ++i) s.w[co * per + i] *= sc;
          bias[co] = acomp.t[ppm + g * ppm + co];
        }
        T dst = Builder::slice(comp_in, ppm + g * ppm, ppm);
        T w = b.conv("spp.scale_process.g" + std::to_string(g), {s}, bias, ppm, true, nullptr, &dst, -1);
        comp_in.prod = w.prod;
      }
    } else {
      T prev = s0;
      T c0 = Builder::slice(comp_in, 0, ppm);
      affine_relu("spp.x0.bnrelu", s0, slice(acomp, 0, ppm), &c0);
      for (int k = 0; k < 4; ++k) {
        const std::string pr = "spp.process" + std::to_string(k + 1);
        Affine a = bn(pr + ".0");
        T yin = upadd(pr + ".upadd", prev, sk[k], &a, true, nullptr);
        prev = conv_plain(pr + ".conv", yin, pr + ".2", nullptr, 3, false, nullptr);
        T dst = Builder::slice(comp_in, (k + 1) * ppm, ppm);
        T w = affine_relu(pr + ".bnrelu", prev, slice(acomp, (k + 1) * ppm, ppm), &dst);
        comp_in.prod = w.prod;
      }
    }
    if (!pyramid) sc_in = affine_relu("spp.shortcut.bnrelu", x, bn("spp.shortcut.0"), nullptr);
    std::vector<float> bias;
    std::vector<ConvSrcSpec> srcs;
    srcs.push_back(src_of(comp_in, "spp.compression.2", nullptr, 1, 1, &bias));
    srcs.push_back(src_of(sc_in, "spp.shortcut.2", nullptr, 1, 1, &bias));
    return b.conv("spp.compression+shortcut", srcs, bias, outp, false, nullptr, nullptr, -1);
  }

  // ---- the net (pidnet.py:136-182)
  void build() {
    const int Pn = cfg.planes, m = cfg.m, n = cfg.n;
    const bool large = m == 3;
    const int L0 = 0, LP = lanes == 3 ? 1 : 0, LD = lanes == 3 ? 2 : 0;
    if (H % 8 || W % 8) fail("H and W must be multiples of 8");
    b.conv_impl = conv_impl;

    // stem: conv1.0 + BN + ReLU -> conv1.3 + BN + ReLU
    b.lane = L0;
    T x;
    const bool stem2 = conv_impl == 0 && b.use_stem2 && (Pn == 32 || Pn == 64) &&
                       P("conv1.3.weight").shape[0] == Pn && P("conv1.3.weight").shape[1] == Pn;
    if (stem2) {
      // ONE kernel; the half-resolution conv1.0 output (the largest tensor of the net) never reaches HBM
      const int H1 = cdiv(H, 2), W1 = cdiv(W, 2), H2 = cdiv(H1, 2), W2 = cdiv(W1, 2);
      x = b.new_tensor(N, H2, W2, Pn);
      Affine a1 = bn("conv1.1"), a2 = bn("conv1.4");
      std::vector<float> w1, bias1, w2, bias2;
      fold("conv1.0", &a1, w1, bias1);   // [P][3][3][3]
      fold("conv1.3", &a2, w2, bias2);   // [P][P][3][3]
      if (P("conv1.0.weight").shape[1] != 3) fail("conv1.0 must have 3 input channels");
      std::vector<uint16_t> w1s(static_cast<size_t>(Pn) * 32, 0);
      for (int co = 0; co < Pn; ++co)
        for (int m = 0; m < 9; ++m)        // m = ci * 3 + r; kernel K order (stem2_tc.cu P1): taps s = 1, 2 at 2m, 2m+1; s = 0 at 18 + m
          for (int sx = 0; sx < 3; ++sx) {
            const int k = sx == 0 ? 18 + m : 2 * m + (sx - 1);
            w1s[co * 32 + (((k / 8) ^ ((co >> 1) & 3)) * 8) + k % 8] = f2bf(w1[static_cast<size_t>(co) * 27 + m * 3 + sx]);
          }
      for (int co = 0; co < Pn; ++co) {   // K columns 27 / 28 (the im2col rows hold 1.0 there): bias as bf16 hi + lo
        const uint16_t hi = f2bf(bias1[co]);
        uint32_t hb = static_cast<uint32_t>(hi) << 16;
        float hf;
        std::memcpy(&hf, &hb, 4);
        w1s[co * 32 + ((3 ^ ((co >> 1) & 3)) * 8) + 3] = hi;
        w1s[co * 32 + ((3 ^ ((co >> 1) & 3)) * 8) + 4] = f2bf(bias1[co] - hf);
      }
      // conv1.3 tap tiles [tap][co][ci], rows of Pn*2 bytes, 16-byte chunks XOR-swizzled the way tcgen05 reads them
      std::vector<uint16_t> w2s(static_cast<size_t>(9) * Pn * Pn, 0);
      const int cmask = Pn == 32 ? 3 : 7;
      for (int tap = 0; tap < 9; ++tap)
        for (int co = 0; co < Pn; ++co)
          for (int ci = 0; ci < Pn; ++ci) {
            const int sw = Pn == 32 ? ((co >> 1) & 3) : (co & 7);
            const size_t pos = (static_cast<size_t>(tap) * Pn + co) * Pn + ((((ci >> 3) ^ sw) & cmask) << 3) + (ci & 7);
            w2s[pos] = f2bf(w2[(static_cast<size_t>(co) * Pn + ci) * 9 + tap]);
          }
      Stem2Params sp;
      std::memset(&sp, 0, sizeof(sp));
      sp.w1_swz = reinterpret_cast<const uint8_t*>(b.alloc_wt(w1s.data(), w1s.size() * 2));
      sp.w2_swz = reinterpret_cast<const uint8_t*>(b.alloc_wt(w2s.data(), w2s.size() * 2));
      sp.bias1 = b.upload_f32(bias1);
      sp.bias2 = b.upload_f32(bias2);
      sp.H = H; sp.W = W; sp.H1 = H1; sp.W1 = W1; sp.H2 = H2; sp.W2 = W2;
      sp.tiles_w = cdiv(W2, 8); sp.tiles_h = cdiv(H2, 16); sp.N = N;
      if (!b.dry) {
        uint64_t dims4[4] = {static_cast<uint64_t>(Pn), static_cast<uint64_t>(W2), static_cast<uint64_t>(H2),
                             static_cast<uint64_t>(N)};
        uint64_t str[3] = {static_cast<uint64_t>(x.ps) * 2, static_cast<uint64_t>(W2) * x.ps * 2,
                           static_cast<uint64_t>(H2) * W2 * x.ps * 2};
        uint32_t box4[4] = {static_cast<uint32_t>(Pn), 8, 16, 1};
        sp.tmD = encode_map(x.ptr, 4, dims4, str, box4, Pn * 2);
      }
      const int sms = b.num_sms;
      const double fl = 2.0 * N * H1 * W1 * Pn * 27 + 2.0 * N * H2 * W2 * Pn * Pn * 9;
      b.flops += fl;
      const int pipelined = b.use_stem2 >= 2 ? 1 : 0;
      const float* mapped_x = nullptr;
      x.prod = b.add_op("conv1.0+conv1.3", {}, [sp, Pn, sms, pipelined, mapped_x](cudaStream_t st, const RunArgs& a) mutable {
        sp.x = a.x; sp.x_u8 = a.x_u8; sp.lut = a.lut;
        if (pipelined && Pn == 32 && a.x && a.x != mapped_x) {
          // the pipelined kernel fetches the fp32 NCHW image with TMA: box {36 cols, 67 rows, 3 channels, 1 image}
          try {
            uint64_t dims4[4] = {static_cast<uint64_t>(sp.W), static_cast<uint64_t>(sp.H), 3, static_cast<uint64_t>(sp.N)};
            uint64_t str[3] = {static_cast<uint64_t>(sp.W) * 4, static_cast<uint64_t>(sp.H) * sp.W * 4,
                               static_cast<uint64_t>(sp.H) * sp.W * 12};
            uint32_t box4[4] = {36, 67, 3, 1};
            sp.tmX = encode_map(a.x, 4, dims4, str, box4, 0, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
          } catch (const std::exception&) {
            return cudaErrorMisalignedAddress;
          }
          mapped_x = a.x;
        }
        return stem2_tc_launch(sp, Pn, sms, pipelined, st);
      });
      b.label(x.prod, pipelined && Pn == 32 ? "stem3_tc" : "stem2_tc", 4.0 * N * 3 * H * W + Builder::tbytes(x), fl);
    } else {
      T x1 = b.new_tensor(N, cdiv(H, 2), cdiv(W, 2), Pn);
      {
        Affine a = bn("conv1.1");
        std::vector<float> w, bias;
        fold("conv1.0", &a, w, bias);  // [P][3][3][3]
        if (P("conv1.0.weight").shape[1] != 3) fail("conv1.0 must have 3 input channels");
        std::vector<float> wt(static_cast<size_t>(27) * Pn);
        for (int co = 0; co < Pn; ++co)
          for (int k = 0; k < 27; ++k) wt[static_cast<size_t>(k) * Pn + co] = w[static_cast<size_t>(co) * 27 + k];
        const float* wd = b.upload_f32(wt);
        const float* bd = b.upload_f32(bias);
        View ov = x1.view();
        const int n_ = N, h_ = H, w_ = W;
        b.flops += 2.0 * N * x1.H * x1.W * Pn * 27;
        const bool stem_tc = conv_impl == 0 && (Pn == 32 || Pn == 64);
        // tcgen05 stem: weights [Cout][32] bf16 K-major, pre-swizzled the way a SWIZZLE_64B tile sits in smem
        std::vector<uint16_t> wsw(static_cast<size_t>(Pn) * 32, 0);
        for (int co = 0; co < Pn; ++co)
          for (int k = 0; k < 27; ++k) {
            const int chunk = k / 8, within = k % 8;
            const int pos = co * 32 + ((chunk ^ ((co >> 1) & 3)) * 8) + within;
            wsw[pos] = f2bf(w[static_cast<size_t>(co) * 27 + k]);
          }
        const uint8_t* wswd = reinterpret_cast<const uint8_t*>(b.alloc_wt(wsw.data(), wsw.size() * 2));
        StemParams sp;
        std::memset(&sp, 0, sizeof(sp));
        sp.w_swz = wswd; sp.bias = bd; sp.H = H; sp.W = W; sp.Ho = x1.H; sp.Wo = x1.W; sp.relu = 1;
        sp.rows = static_cast<long>(N) * x1.H * x1.W;
        sp.tiles = (sp.rows + 127) / 128;
        if (stem_tc && !b.dry) {
          uint64_t dims4[4] = {static_cast<uint64_t>(Pn), static_cast<uint64_t>(sp.rows), 1, 1};
          uint64_t str[3] = {static_cast<uint64_t>(Pn) * 2, static_cast<uint64_t>(sp.rows) * Pn * 2,
                             static_cast<uint64_t>(sp.rows) * Pn * 2};
          uint32_t box4[4] = {static_cast<uint32_t>(Pn), 128, 1, 1};
          sp.tmD = encode_map(x1.ptr, 4, dims4, str, box4, Pn * 2);
        }
        const int sms = b.num_sms;
        x1.prod = b.add_op("conv1.0", {}, [ov, wd, bd, n_, h_, w_, sp, stem_tc, Pn, sms](cudaStream_t st, const RunArgs& a) mutable {
          if (stem_tc) {
            sp.x = a.x;
            sp.x_u8 = a.x_u8;
            sp.lut = a.lut;
            return stem_tc_launch(sp, Pn, sms, st);
          }
          if (a.x_u8) return cudaErrorNotSupported;   // the uint8 path exists in the tcgen05 stem only
          return stem_conv_launch(a.x, n_, h_, w_, ov, wd, bd, st);
        });
        b.label(x1.prod, stem_tc ? "stem_tc" : "stem_conv", 4.0 * N * 3 * H * W + Builder::tbytes(x1),
                2.0 * N * x1.H * x1.W * Pn * 27);
      }
      x = conv_bn("conv1.3", x1, "conv1.3", "conv1.4", 3, 2, true);
    }
    b.named["conv1"] = x;
    x = layer("layer1", x, false, m, 1, true);
    b.named["layer1"] = x;
    x = layer("layer2", x, false, m, 2, true);
    b.named["layer2"] = x;
    const T x8 = x;

    b.lane = LP;
    T xp = layer("layer3_", x8, false, m, 1, false);
    b.named["layer3_"] = xp;
    b.lane = LD;
    T xd = basic_block("layer3_d", x8, 1, false);
    b.named["layer3_d"] = xd;

    b.lane = L0;
    T xi = layer("layer3", x8, false, n, 2, true);
    b.named["layer3"] = xi;
    b.lane = LP;
    xp = pag("pag3", "compression3", xp, xi);
    b.named["pag3"] = xp;
    const T temp_p = xp;
    b.lane = LD;
    {
      T d3 = conv_bn("diff3", xi, "diff3.0", "diff3.1", 3, 1, false);
      xd = upadd("diff3.upadd", xd, d3, nullptr, true, nullptr);
      b.named["xd3"] = xd;
    }
    b.lane = L0;
    T xi4 = layer("layer4", xi, false, n, 2, true);
    b.named["layer4"] = xi4;
    b.lane = LP;
    xp = layer("layer4_", xp, false, m, 1, false);
    b.named["layer4_"] = xp;
    b.lane = LD;
    xd = large ? basic_block("layer4_d", xd, 1, false) : layer("layer4_d", xd, true, 1, 1, false);
    b.named["layer4_d"] = xd;
    b.lane = LP;
    xp = pag("pag4", "compression4", xp, xi4);
    b.named["pag4"] = xp;
    b.lane = LD;
    {
      T d4 = conv_bn("diff4", xi4, "diff4.0", "diff4.1", 3, 1, false);
      xd = upadd("diff4.upadd", xd, d4, nullptr, true, nullptr);
      b.named["xd4"] = xd;
    }
    const T temp_d = xd;
    b.lane = LP;
    xp = layer("layer5_", xp, true, 1, 1, false);
    b.named["layer5_"] = xp;
    if (cfg.augment) {
      T hp = affine_relu("seghead_p.bn1", temp_p, bn("seghead_p.bn1"), nullptr);
      seghead_tail("seghead_p", hp, 1);
    }
    b.lane = LD;
    xd = layer("layer5_d", xd, true, 1, 1, false);
    b.named["layer5_d"] = xd;
    if (cfg.augment) {
      T hd = affine_relu("seghead_d.bn1", temp_d, bn("seghead_d.bn1"), nullptr);
      seghead_tail("seghead_d", hd, 2);
    }
    b.lane = L0;
    T x5 = layer("layer5", xi4, true, 2, 2, false);
    b.named["layer5"] = x5;
    T sp = spp(x5);
    b.named["spp"] = sp;

    // dfm + final_layer.bn1/ReLU folded into its epilogue
    const Affine a1 = bn("final_layer.bn1");
    T f;
    if (!large) {
      const int C4 = 4 * Pn;
      T uv = b.new_tensor(N, xp.H, xp.W, 2 * C4);
      View pv = xp.view(), iv = sp.view(), dv = xd.view(), ov = uv.view();
      uv.prod = b.add_op("dfm.uv", {&xp, &sp, &xd}, [pv, iv, dv, ov](cudaStream_t st, const RunArgs&) {
        return lightbag_uv_launch(pv, iv, dv, ov, st);
      });
      b.label(uv.prod, "lightbag_uv", Builder::tbytes(xp) + Builder::tbytes(sp) + Builder::tbytes(xd) + Builder::tbytes(uv));
      const Affine ap = bn("dfm.conv_p.1"), ai = bn("dfm.conv_i.1");
      const auto &Wp = P("dfm.conv_p.0.weight").data, &Wi = P("dfm.conv_i.0.weight").data;
      ConvSrcSpec s;
      s.in = uv; s.k = 1; s.stride = 1;
      s.w.resize(static_cast<size_t>(C4) * 2 * C4);
      std::vector<float> bias(C4);
      for (int co = 0; co < C4; ++co) {
        for (int ci = 0; ci < C4; ++ci) {
          s.w[static_cast<size_t>(co) * 2 * C4 + ci] = a1.s[co] * ap.s[co] * Wp[static_cast<size_t>(co) * C4 + ci];
          s.w[static_cast<size_t>(co) * 2 * C4 + C4 + ci] = a1.s[co] * ai.s[co] * Wi[static_cast<size_t>(co) * C4 + ci];
        }
        bias[co] = a1.s[co] * (ap.t[co] + ai.t[co]) + a1.t[co];
      }
      f = b.conv("dfm.conv_p+conv_i", {s}, bias, C4, true, nullptr, nullptr, -1);
    } else {
      const Affine ab = bn("dfm.conv.0");
      T a = b.new_tensor(N, xp.H, xp.W, xp.C);
      const float* s_ = b.upload_f32(ab.s);
      const float* t_ = b.upload_f32(ab.t);
      View pv = xp.view(), iv = sp.view(), dv = xd.view(), ov = a.view();
      a.prod = b.add_op("dfm.blend", {&xp, &sp, &xd}, [pv, iv, dv, ov, s_, t_](cudaStream_t st, const RunArgs&) {
        return bag_blend_launch(pv, iv, dv, ov, s_, t_, st);
      });
      b.label(a.prod, "bag_blend", Builder::tbytes(xp) + Builder::tbytes(sp) + Builder::tbytes(xd) + Builder::tbytes(a));
      f = conv_plain("dfm.conv", a, "dfm.conv.2", &a1, 3, true, nullptr);
    }
    b.named["dfm"] = f;
    seghead_tail("final_layer", f, 0);
  }

  void plan(int N_, int H_, int W_) {
    release();
    N = N_; H = H_; W = W_;
    cudaDeviceProp prop;
    int dev = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaGetDeviceProperties(&prop, dev));
    b.num_sms = prop.multiProcessorCount;
    if (conv_impl == 0) {
      if (prop.major != 10) fail("the tcgen05 conv path needs an sm_100 GPU (found sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + ")");
      CK(conv_tc_init());
      CK(conv3_ws_init());
    }
    b.use_ws = use_ws;
    if (use_pair >= 0) b.use_pair = use_pair;
    if (ws_stages >= 0) b.ws_stages = ws_stages;
    if (use_stem2 >= 0) b.use_stem2 = use_stem2;
    if (use_pyramid >= 0) b.use_pyramid = use_pyramid;
    b.reset(true);
    build();  // dry pass: sizes only
    const size_t act_bytes = b.act_cur, wt_bytes = b.wt_cur;
    CK(cudaMalloc(&b.act_base, std::max<size_t>(act_bytes, 1024)));
    CK(cudaMalloc(&b.wt_base, std::max<size_t>(wt_bytes, 1024)));
    b.wt_host.assign(wt_bytes, 0);
    b.reset(false);
    build();
    if (b.act_cur != act_bytes || b.wt_cur != wt_bytes) fail("planner passes disagree");
    CK(cudaMemcpy(b.wt_base, b.wt_host.data(), wt_bytes, cudaMemcpyHostToDevice));
    CK(cudaMemset(b.act_base, 0, act_bytes));
    b.wt_host.clear();
    b.wt_host.shrink_to_fit();
    // streams / events
    for (int i = 0; i < 2; ++i) CK(cudaStreamCreateWithFlags(&side[i], cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&cap_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) CK(cudaEventCreateWithFlags(&ev_join[i], cudaEventDisableTiming));
    ev_of_op.assign(b.ops.size(), -1);
    for (size_t i = 0; i < b.ops.size(); ++i)
      if (b.ops[i].record) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ev_of_op[i] = static_cast<int>(events.size());
        events.push_back(e);
      }
    planned = true;
  }

  void enqueue(cudaStream_t stream, const RunArgs& a) {
    bool used[3] = {true, false, false};
    for (const Op& op : b.ops) used[op.lane] = true;
    CK(cudaEventRecord(ev_start, stream));
    for (int l = 1; l < 3; ++l)
      if (used[l]) CK(cudaStreamWaitEvent(side[l - 1], ev_start, 0));
    for (size_t i = 0; i < b.ops.size(); ++i) {
      const Op& op = b.ops[i];
      cudaStream_t s = op.lane == 0 ? stream : side[op.lane - 1];
      for (int d : op.deps) CK(cudaStreamWaitEvent(s, events[ev_of_op[d]], 0));
      cudaError_t e = op.fn(s, a);
      if (e != cudaSuccess) fail("launch of '" + op.name + "' failed: " + cudaGetErrorString(e));
      if (op.record) CK(cudaEventRecord(events[ev_of_op[i]], s));
    }
    for (int l = 1; l < 3; ++l)
      if (used[l]) {
        CK(cudaEventRecord(ev_join[l - 1], side[l - 1]));
        CK(cudaStreamWaitEvent(stream, ev_join[l - 1], 0));
      }
  }

  // one launch at a time on `stream`, CUDA events around each: per-op device time in ms
  void profile(cudaStream_t stream, const RunArgs& a, float* ms, int cap) {
    if (!planned) fail("pidnet_profile called before pidnet_plan");
    const int n = static_cast<int>(b.ops.size());
    if (cap < n) fail("profile buffer too small");
    std::vector<cudaEvent_t> ev(n + 1);
    for (auto& e : ev) CK(cudaEventCreate(&e));
    CK(cudaEventRecord(ev[0], stream));
    for (int i = 0; i < n; ++i) {
      cudaError_t e = b.ops[i].fn(stream, a);
      if (e != cudaSuccess) fail("launch of '" + b.ops[i].name + "' failed: " + cudaGetErrorString(e));
      CK(cudaEventRecord(ev[i + 1], stream));
    }
    CK(cudaStreamSynchronize(stream));
    for (int i = 0; i < n; ++i) CK(cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]));
    for (auto& e : ev) cudaEventDestroy(e);
  }

  void forward(cudaStream_t stream, const RunArgs& a, bool use_graph) {
    if (!planned) fail("pidnet_forward called before pidnet_plan");
    if ((!a.x && !a.x_u8) || !a.out[0]) fail("null input/output pointer");
    if (cfg.augment && (!a.out[1] || !a.out[2])) fail("augment=1 needs out_p and out_d");
    if (!use_graph) {
      enqueue(stream, a);
      return;
    }
    for (auto& g : graphs)
      if (std::memcmp(&g.args, &a, sizeof(RunArgs)) == 0) {
        CK(cudaGraphLaunch(g.exec, stream));
        return;
      }
    if (graphs.size() >= kMaxGraphs) {
      cudaGraphExecDestroy(graphs.front().exec);
      graphs.erase(graphs.begin());
    }
    cudaGraph_t g = nullptr;
    cudaGraphExec_t exec = nullptr;
    CK(cudaStreamBeginCapture(cap_stream, cudaStreamCaptureModeThreadLocal));
    try {
      enqueue(cap_stream, a);
    } catch (...) {
      cudaStreamEndCapture(cap_stream, &g);
      if (g) cudaGraphDestroy(g);
      throw;
    }
    CK(cudaStreamEndCapture(cap_stream, &g));
    cudaError_t e = cudaGraphInstantiate(&exec, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) fail(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
    graphs.push_back(GraphEntry{a, exec});
    CK(cudaGraphLaunch(exec, stream));
  }
};

// ----------------------------------------------------------------------------------------- C ABI glue
template <class F>
static int guard(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  } catch (...) {
    g_err = "unknown error";
    return -2;
  }
}

static T ext_tensor(const void* ptr, int N, int H, int W, int C) {
  T t;
  t.ptr = reinterpret_cast<bf16*>(const_cast<void*>(ptr));
  t.N = N; t.H = H; t.W = W; t.C = C; t.ps = C;
  return t;
}

#include "train.inc"

}  // namespace pidnet

using namespace pidnet;

struct pidnet_engine {
  Engine e;
};

#pragma GCC visibility push(default)
extern "C" {

const char* pidnet_last_error(void) { return g_err.c_str(); }
int pidnet_abi_version(void) { return 2; }
unsigned pidnet_debug_fastdiv(unsigned n, unsigned d) { return d ? fastdiv_debug(n, d) : 0u; }

int pidnet_create(const pidnet_cfg* cfg, pidnet_engine** out) {
  return guard([&] {
    if (!cfg || !out) fail("null argument");
    if (!((cfg->m == 2 && cfg->n >= 1) || (cfg->m == 3 && cfg->n >= 1))) fail("m must be 2 (S/M) or 3 (L)");
    if (cfg->planes % 8 || cfg->ppm_planes % 8 || cfg->head_planes % 8) fail("planes/ppm_planes/head_planes must be multiples of 8");
    if (cfg->num_classes < 1) fail("num_classes must be positive");
    pidnet_engine* h = new pidnet_engine();
    h->e.cfg = *cfg;
    *out = h;
  });
}

int pidnet_destroy(pidnet_engine* h) {
  return guard([&] { delete h; });
}

int pidnet_set_param(pidnet_engine* h, const char* key, const float* host_data, const int64_t* shape, int ndim) {
  return guard([&] {
    if (!h || !key || !host_data) fail("null argument");
    HostParam p;
    size_t n = 1;
    for (int i = 0; i < ndim; ++i) {
      p.shape.push_back(shape[i]);
      n *= static_cast<size_t>(shape[i]);
    }
    p.data.assign(host_data, host_data + n);
    h->e.params[key] = std::move(p);
    h->e.planned = false;
  });
}

int pidnet_set_option(pidnet_engine* h, const char* name, int value) {
  return guard([&] {
    if (!h || !name) fail("null argument");
    const std::string k = name;
    if (k == "conv_impl") h->e.conv_impl = value;
    else if (k == "lanes") h->e.lanes = value == 3 ? 3 : 1;
    else if (k == "use_ws") h->e.use_ws = value ? 1 : 0;
    else if (k == "use_pair") h->e.use_pair = value ? 1 : 0;
    else if (k == "use_stem2") h->e.use_stem2 = value < 0 ? 0 : (value > 2 ? 2 : value);
    else if (k == "use_pyramid") h->e.use_pyramid = value ? 1 : 0;
    else if (k == "ws_stages") h->e.ws_stages = value == 2 ? 2 : 3;
    else if (k == "fp32_head") h->e.fp32_head = value ? 1 : 0;
    else fail("unknown option '" + k + "'");
    h->e.planned = false;
  });
}

int pidnet_plan(pidnet_engine* h, int N, int H, int W, size_t* arena_bytes) {
  return guard([&] {
    if (!h) fail("null handle");
    if (N < 1 || H < 8 || W < 8) fail("bad input shape");
    h->e.plan(N, H, W);
    if (arena_bytes) *arena_bytes = h->e.b.act_cur + h->e.b.wt_cur;
  });
}

int pidnet_forward(pidnet_engine* h, void* stream, const float* x, float* out_main, float* out_p, float* out_d,
                   int use_graph) {
  return guard([&] {
    if (!h) fail("null handle");
    RunArgs a{x, {out_main, out_p, out_d}};
    h->e.forward(reinterpret_cast<cudaStream_t>(stream), a, use_graph != 0);
  });
}

int pidnet_forward_u8(pidnet_engine* h, void* stream, const unsigned char* bgr_hwc, const double* mean_rgb, const double* std_rgb,
                      float* out_main, float* out_p, float* out_d, int use_graph) {
  return guard([&] {
    if (!h || !bgr_hwc || !mean_rgb || !std_rgb) fail("null argument");
    RunArgs a{nullptr, {out_main, out_p, out_d}};
    a.x_u8 = bgr_hwc;
    a.lut = h->e.u8_lut(mean_rgb, std_rgb, reinterpret_cast<cudaStream_t>(stream));
    h->e.forward(reinterpret_cast<cudaStream_t>(stream), a, use_graph != 0);
  });
}

int pidnet_profile(pidnet_engine* h, void* stream, const float* x, float* out_main, float* out_p, float* out_d,
                   float* ms_per_op, int cap) {
  return guard([&] {
    if (!h || !ms_per_op) fail("null argument");
    RunArgs a{x, {out_main, out_p, out_d}};
    h->e.profile(reinterpret_cast<cudaStream_t>(stream), a, ms_per_op, cap);
  });
}

int pidnet_op_info(pidnet_engine* h, int i, char* name, int name_cap, char* kernel, int kernel_cap, double* flops,
                   double* bytes, int* lane) {
  return guard([&] {
    if (!h || i < 0 || i >= static_cast<int>(h->e.b.ops.size())) fail("bad op index");
    const Op& op = h->e.b.ops[i];
    if (name && name_cap > 0) std::snprintf(name, name_cap, "%s", op.name.c_str());
    if (kernel && kernel_cap > 0) std::snprintf(kernel, kernel_cap, "%s", op.kernel.c_str());
    if (flops) *flops = op.flops;
    if (bytes) *bytes = op.bytes;
    if (lane) *lane = op.lane;
  });
}

int pidnet_num_launches(pidnet_engine* h) { return h ? static_cast<int>(h->e.b.ops.size()) : -1; }
double pidnet_conv_flops(pidnet_engine* h) { return h ? h->e.b.flops : -1.0; }

int pidnet_debug_tensor(pidnet_engine* h, const char* name, float* host_out, int64_t* shape4) {
  return guard([&] {
    if (!h || !name) fail("null argument");
    auto it = h->e.b.named.find(name);
    if (it == h->e.b.named.end()) fail(std::string("no tensor named '") + name + "'");
    const T& t = it->second;
    if (shape4) { shape4[0] = t.N; shape4[1] = t.C; shape4[2] = t.H; shape4[3] = t.W; }
    if (!host_out) return;
    CK(cudaDeviceSynchronize());
    const size_t npix = static_cast<size_t>(t.N) * t.H * t.W;
    std::vector<uint16_t> tmp(npix * t.ps);
    CK(cudaMemcpy(tmp.data(), t.ptr, (npix - 1) * t.ps * 2 + static_cast<size_t>(t.C) * 2, cudaMemcpyDeviceToHost));
    for (int n = 0; n < t.N; ++n)
      for (int c = 0; c < t.C; ++c)
        for (int y = 0; y < t.H; ++y)
          for (int x = 0; x < t.W; ++x)
            host_out[((static_cast<size_t>(n) * t.C + c) * t.H + y) * t.W + x] =
                bf2f(tmp[((static_cast<size_t>(n) * t.H + y) * t.W + x) * t.ps + c]);
  });
}

// ---- single-op entry points
int pidnet_op_conv2d(void* stream, const void* x, int N, int H, int W, int Cin, const float* w, const float* bias,
                     int Cout, int k, int stride, int groups, const void* res, int relu, void* out_nhwc,
                     float* out_nchw_f32, int impl) {
  return guard([&] {
    if (!x || !w || (!out_nhwc == !out_nchw_f32)) fail("bad arguments");
    if (groups < 1 || Cin % groups || Cout % groups) fail("bad groups");
    if (groups > 1 && out_nchw_f32) fail("grouped conv with NCHW output is not supported");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (impl != 1) {
      CK(conv_tc_init());
      CK(conv3_ws_init());
    }
    const int Ho = cdiv(H, stride), Wo = cdiv(W, stride);
    const int cig = Cin / groups, cog = Cout / groups;
    Builder b;
    cudaDeviceProp prop;
    int dev = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaGetDeviceProperties(&prop, dev));
    b.num_sms = prop.multiProcessorCount;
    b.conv_impl = impl == 1 ? 1 : 0;
    b.use_ws = impl == 2 ? 0 : 1;
    if (impl == 3) { b.use_pair = 0; b.ws_stages = 2; }   // weight-stationary kernels without CTA pairs / rotating stages
    if (impl == 4) b.use_pair = 0;
    T xin = ext_tensor(x, N, H, W, Cin);
    T rt, ot;
    if (res) rt = ext_tensor(res, N, Ho, Wo, Cout);
    if (out_nhwc) ot = ext_tensor(out_nhwc, N, Ho, Wo, Cout);
    auto build = [&] {
      for (int g = 0; g < groups; ++g) {
        ConvSrcSpec s;
        s.in = groups > 1 ? Builder::slice(xin, g * cig, cig) : xin;
        s.k = k; s.stride = stride;
        const size_t per = static_cast<size_t>(cig) * k * k;
        s.w.assign(w + static_cast<size_t>(g) * cog * per, w + static_cast<size_t>(g + 1) * cog * per);
        std::vector<float> bs(cog, 0.f);
        if (bias) bs.assign(bias + g * cog, bias + (g + 1) * cog);
        T rs, os;
        if (res) rs = groups > 1 ? Builder::slice(rt, g * cog, cog) : rt;
        if (out_nhwc) os = groups > 1 ? Builder::slice(ot, g * cog, cog) : ot;
        b.conv("op_conv2d", {s}, bs, cog, relu != 0, res ? &rs : nullptr, out_nhwc ? &os : nullptr,
               out_nchw_f32 ? 0 : -1);
      }
    };
    b.reset(true);
    build();
    const size_t wt_bytes = b.wt_cur;
    CK(cudaMalloc(&b.wt_base, std::max<size_t>(wt_bytes, 1024)));
    b.wt_host.assign(wt_bytes, 0);
    b.reset(false);
    try {
      build();
      CK(cudaMemcpyAsync(b.wt_base, b.wt_host.data(), wt_bytes, cudaMemcpyHostToDevice, st));
      RunArgs a{nullptr, {out_nchw_f32, nullptr, nullptr}};
      for (auto& op : b.ops) {
        cudaError_t e = op.fn(st, a);
        if (e != cudaSuccess) fail(std::string("conv launch failed: ") + cudaGetErrorString(e));
      }
      CK(cudaStreamSynchronize(st));
    } catch (...) {
      cudaFree(b.wt_base);
      throw;
    }
    CK(cudaFree(b.wt_base));
  });
}

int pidnet_op_stem(void* stream, const float* x, int N, int H, int W, const float* w, const float* bias, int Cout,
                   void* out) {
  return guard([&] {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    std::vector<float> wt(static_cast<size_t>(27) * Cout), bs(Cout, 0.f);
    for (int co = 0; co < Cout; ++co)
      for (int k = 0; k < 27; ++k) wt[static_cast<size_t>(k) * Cout + co] = w[static_cast<size_t>(co) * 27 + k];
    if (bias) bs.assign(bias, bias + Cout);
    float* d = nullptr;
    CK(cudaMalloc(&d, (wt.size() + bs.size()) * 4));
    CK(cudaMemcpyAsync(d, wt.data(), wt.size() * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d + wt.size(), bs.data(), bs.size() * 4, cudaMemcpyHostToDevice, st));
    T o = ext_tensor(out, N, cdiv(H, 2), cdiv(W, 2), Cout);
    cudaError_t e = stem_conv_launch(x, N, H, W, o.view(), d, d + wt.size(), st);
    cudaError_t e2 = cudaStreamSynchronize(st);
    cudaFree(d);
    CK(e);
    CK(e2);
  });
}

int pidnet_op_pag(void* stream, const void* x, const void* low, void* out, int N, int H, int W, int C, int h, int w,
                  int relu) {
  return guard([&] {
    CK(pag_fuse_launch(ext_tensor(x, N, H, W, C).view(), ext_tensor(low, N, h, w, 2 * C + 8).view(),
                       ext_tensor(out, N, H, W, C).view(), relu, reinterpret_cast<cudaStream_t>(stream)));
  });
}

int pidnet_op_upadd(void* stream, const void* a, const void* bb, void* out, int N, int H, int W, int C, int h, int w,
                    const float* s, const float* t, int relu) {
  return guard([&] {
    View av = a ? ext_tensor(a, N, H, W, C).view() : View{nullptr, 0, 0, 0, 0, 0};
    View bv = bb ? ext_tensor(bb, N, h, w, C).view() : View{nullptr, 0, 0, 0, 0, 0};
    CK(upadd_launch(av, bv, ext_tensor(out, N, H, W, C).view(), s, t, relu, reinterpret_cast<cudaStream_t>(stream)));
  });
}

int pidnet_op_pool(void* stream, const void* x, void* out, int N, int H, int W, int C, int k, int stride, int pad,
                   const float* s, const float* t, int relu) {
  return guard([&] {
    const int oh = k == 0 ? 1 : (H + 2 * pad - k) / stride + 1, ow = k == 0 ? 1 : (W + 2 * pad - k) / stride + 1;
    CK(pool_affine_launch(ext_tensor(x, N, H, W, C).view(), ext_tensor(out, N, oh, ow, C).view(), k, stride, pad, s, t,
                          relu, reinterpret_cast<cudaStream_t>(stream)));
  });
}

int pidnet_op_lightbag(void* stream, const void* p, const void* i_low, const void* d, void* out, int N, int H, int W,
                       int C, int h, int w) {
  return guard([&] {
    CK(lightbag_uv_launch(ext_tensor(p, N, H, W, C).view(), ext_tensor(i_low, N, h, w, C).view(),
                          ext_tensor(d, N, H, W, C).view(), ext_tensor(out, N, H, W, 2 * C).view(),
                          reinterpret_cast<cudaStream_t>(stream)));
  });
}

int pidnet_op_bag(void* stream, const void* p, const void* i_low, const void* d, void* out, int N, int H, int W, int C,
                  int h, int w, const float* s, const float* t) {
  return guard([&] {
    CK(bag_blend_launch(ext_tensor(p, N, H, W, C).view(), ext_tensor(i_low, N, h, w, C).view(),
                        ext_tensor(d, N, H, W, C).view(), ext_tensor(out, N, H, W, C).view(), s, t,
                        reinterpret_cast<cudaStream_t>(stream)));
  });
}


// ---- training (train.inc)
struct pidnet_trainer_ { TrainNet t; };

int pidnet_train_create(const pidnet_cfg* cfg, pidnet_trainer** out) {
  return guard([&] {
    if (!cfg || !out) fail("null argument");
    pidnet_trainer_* h = new pidnet_trainer_();
    h->t.cfg = *cfg;
    *out = reinterpret_cast<pidnet_trainer*>(h);
  });
}
int pidnet_train_destroy(pidnet_trainer* h) {
  return guard([&] { delete reinterpret_cast<pidnet_trainer_*>(h); });
}
int pidnet_train_bind(pidnet_trainer* h, const char* key, float* dev_param, float* dev_grad, const int64_t* shape, int ndim) {
  return guard([&] {
    if (!h || !key || !dev_param) fail("null argument");
    TrainParam p;
    p.w = dev_param; p.g = dev_grad;
    for (int i = 0; i < ndim; ++i) p.shape.push_back(shape[i]);
    TrainNet& t = reinterpret_cast<pidnet_trainer_*>(h)->t;
    t.P[key] = p;
    t.planned = false;
  });
}
int pidnet_train_plan(pidnet_trainer* h, int N, int H, int W, size_t* arena_bytes) {
  return guard([&] {
    if (!h) fail("null handle");
    TrainNet& t = reinterpret_cast<pidnet_trainer_*>(h)->t;
    t.plan(N, H, W);
    if (arena_bytes) *arena_bytes = t.b.act_cur + t.b.wt_cur;
  });
}
int pidnet_train_step(pidnet_trainer* h, void* stream, const float* x, const int64_t* labels, const float* bd_gt,
                      const float* class_weights, const pidnet_criterion_cfg* cfg, int backward, float* out12,
                      float* out_main, float* out_p, float* out_d, float* aux_ce_map) {
  return guard([&] {
    if (!h || !x || !labels || !bd_gt || !cfg) fail("null argument");
    TrainNet& t = reinterpret_cast<pidnet_trainer_*>(h)->t;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    t.aux_ce_map = aux_ce_map;
    if (backward < 0 || backward > 2) fail("pidnet_train_step: backward must be 0, 1 or 2");
    t.step(st, x, labels, bd_gt, class_weights, *cfg, backward);
    const size_t lp = static_cast<size_t>(t.N) * t.h8 * t.w8;
    if (out12) CK(cudaMemcpyAsync(out12, t.out12, 16 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (out_main) CK(cudaMemcpyAsync(out_main, t.logits[0], lp * t.cfg.num_classes * 4, cudaMemcpyDeviceToDevice, st));
    if (out_p) CK(cudaMemcpyAsync(out_p, t.logits[1], lp * t.cfg.num_classes * 4, cudaMemcpyDeviceToDevice, st));
    if (out_d) CK(cudaMemcpyAsync(out_d, t.logits[2], lp * 4, cudaMemcpyDeviceToDevice, st));
  });
}
/* train-mode forward only (batch statistics, running-stat update), no criterion / backward */
int pidnet_train_forward(pidnet_trainer* h, void* stream, const float* x, float* out_main, float* out_p, float* out_d) {
  return guard([&] {
    if (!h || !x) fail("null argument");
    TrainNet& t = reinterpret_cast<pidnet_trainer_*>(h)->t;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    pidnet_criterion_cfg none{};
    t.step(st, x, nullptr, nullptr, nullptr, none, 0);
    const size_t lp = static_cast<size_t>(t.N) * t.h8 * t.w8;
    if (out_main) CK(cudaMemcpyAsync(out_main, t.logits[0], lp * t.cfg.num_classes * 4, cudaMemcpyDeviceToDevice, st));
    if (out_p) CK(cudaMemcpyAsync(out_p, t.logits[1], lp * t.cfg.num_classes * 4, cudaMemcpyDeviceToDevice, st));
    if (out_d) CK(cudaMemcpyAsync(out_d, t.logits[2], lp * 4, cudaMemcpyDeviceToDevice, st));
  });
}
/* network backward of the last train-mode forward (pidnet_train_forward, or pidnet_train_step with backward = 0 / 2) */
int pidnet_train_backward(pidnet_trainer* h, void* stream, const float* x, const float* g_main, const float* g_p, const float* g_d,
                          int segment) {
  return guard([&] {
    if (!h) fail("null handle");
    TrainNet& t = reinterpret_cast<pidnet_trainer_*>(h)->t;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool any = g_main || g_p || g_d;
    if (any && !(g_main && g_p && g_d)) fail("pidnet_train_backward: pass all three logit gradients or none");
    if (any && segment <= 0) {
      if (!t.planned) fail("pidnet_train_backward called before pidnet_train_plan");
      const size_t lp = static_cast<size_t>(t.N) * t.h8 * t.w8;
      CK(cudaMemcpyAsync(t.dlogits[0], g_main, lp * t.cfg.num_classes * 4, cudaMemcpyDeviceToDevice, st));
      CK(cudaMemcpyAsync(t.dlogits[1], g_p, lp * t.cfg.num_classes * 4, cudaMemcpyDeviceToDevice, st));
      CK(cudaMemcpyAsync(t.dlogits[2], g_d, lp * 4, cudaMemcpyDeviceToDevice, st));
    }
    t.backward_from_logits(st, x, segment);
  });
}
int pidnet_train_num_segments(pidnet_trainer* h) {
  return h ? reinterpret_cast<pidnet_trainer_*>(h)->t.nseg : -1;
}
int pidnet_train_segment_ranges(pidnet_trainer* h, int segment, const float* grad_base, int64_t* begin_end, int cap_pairs,
                                int* n_pairs) {
  return guard([&] {
    if (!h || !grad_base || !n_pairs) fail("null argument");
    TrainNet& t = reinterpret_cast<pidnet_trainer_*>(h)->t;
    if (!t.planned) fail("pidnet_train_segment_ranges called before pidnet_train_plan");
    if (segment < 0 || segment >= t.nseg) fail("segment index out of range");
    const auto rs = t.segment_ranges(segment);
    *n_pairs = static_cast<int>(rs.size());
    if (!begin_end) return;
    if (static_cast<int>(rs.size()) > cap_pairs) fail("pidnet_train_segment_ranges: buffer too small");
    for (size_t i = 0; i < rs.size(); ++i) {
      begin_end[2 * i] = static_cast<int64_t>(rs[i].first - grad_base);
      begin_end[2 * i + 1] = begin_end[2 * i] + static_cast<int64_t>(rs[i].second);
    }
  });
}
/* measurement: per-launch device times of one training step.  Writes a text table (one line per launch:
 * "F|B <ms> <kernel> <name>") into buf; returns the criterion time separately. */
int pidnet_train_profile(pidnet_trainer* h, void* stream, const float* x, const int64_t* labels, const float* bd_gt,
                         const float* class_weights, const pidnet_criterion_cfg* cfg, char* buf, size_t cap, float* crit_ms) {
  return guard([&] {
    TrainNet& t = reinterpret_cast<pidnet_trainer_*>(h)->t;
    std::vector<float> f, bw;
    float cm = 0;
    t.profile(reinterpret_cast<cudaStream_t>(stream), x, labels, bd_gt, class_weights, *cfg, f, bw, cm);
    if (crit_ms) *crit_ms = cm;
    std::string out;
    char line[512];
    for (size_t i = 0; i < f.size(); ++i) {
      std::snprintf(line, sizeof(line), "F %.5f %s %s\n", f[i], t.b.ops[i].kernel.empty() ? "-" : t.b.ops[i].kernel.c_str(), t.b.ops[i].name.c_str());
      out += line;
    }
    for (size_t i = 0; i < bw.size(); ++i) {
      const Op& op = i < t.bops.size() ? t.bops[i] : t.post_ops[i - t.bops.size()];
      std::snprintf(line, sizeof(line), "B %.5f %s %s\n", bw[i], op.kernel.empty() ? "-" : op.kernel.c_str(), op.name.c_str());
      out += line;
    }
    if (buf && cap) std::snprintf(buf, cap, "%s", out.c_str());
  });
}
static void dump_nhwc(const T& t, float* host_out, int64_t* shape4) {
  if (shape4) { shape4[0] = t.N; shape4[1] = t.C; shape4[2] = t.H; shape4[3] = t.W; }
  if (!host_out) return;
  CK(cudaDeviceSynchronize());
  const size_t npix = static_cast<size_t>(t.N) * t.H * t.W;
  std::vector<uint16_t> tmp(npix * t.ps);
  CK(cudaMemcpy(tmp.data(), t.ptr, (npix - 1) * t.ps * 2 + static_cast<size_t>(t.C) * 2, cudaMemcpyDeviceToHost));
  for (int n = 0; n < t.N; ++n)
    for (int c = 0; c < t.C; ++c)
      for (int y = 0; y < t.H; ++y)
        for (int x = 0; x < t.W; ++x)
          host_out[((static_cast<size_t>(n) * t.C + c) * t.H + y) * t.W + x] =
              bf2f(tmp[((static_cast<size_t>(n) * t.H + y) * t.W + x) * t.ps + c]);
}
/* debug: named stage tensor (grad != 0: its gradient) of the last training step as fp32 NCHW on the host */
int pidnet_train_debug_tensor(pidnet_trainer* h, const char* name, int grad, float* host_out, int64_t* shape4) {
  return guard([&] {
    TrainNet& t = reinterpret_cast<pidnet_trainer_*>(h)->t;
    auto it = t.named.find(name);
    if (it == t.named.end()) fail(std::string("no training tensor named '") + name + "'");
    dump_nhwc(grad ? t.tt[it->second].g : t.tt[it->second].v, host_out, shape4);
  });
}
int pidnet_train_set_option(pidnet_trainer* h, const char* name, int value) {
  return guard([&] {
    if (!h || !name) fail("null argument");
    TrainNet& t = reinterpret_cast<pidnet_trainer_*>(h)->t;
    if (std::string(name) == "use_graph") {
      t.use_graph = value != 0;
      t.drop_graphs();
    } else if (std::string(name) == "wgrad_halo") {
      t.wgrad_halo = value != 0;
      t.planned = false;   // needs a re-plan
      t.drop_graphs();
    } else if (std::string(name) == "wgrad_stack") {
      t.wgrad_stack = value != 0;
      t.planned = false;   // needs a re-plan
      t.drop_graphs();
    } else if (std::string(name) == "fused_bn") {
      t.fused_bn = value != 0;
      t.planned = false;   // needs a re-plan
      t.drop_graphs();
    } else if (std::string(name) == "conv_stats") {
      t.conv_stats = value;
      t.planned = false;   // needs a re-plan
      t.drop_graphs();
    } else if (std::string(name) == "overlap_wgrad") {
      t.overlap_wgrad = value != 0;
      t.drop_graphs();
    } else if (std::string(name) == "bwd_segments") {
      if (value < 1 || value > 16) fail("bwd_segments must be in 1..16");
      t.nseg = value;
      t.drop_graphs();
    } else {
      fail(std::string("unknown training option '") + name + "'");
    }
  });
}
int pidnet_train_num_launches(pidnet_trainer* h, int* fwd, int* bwd) {
  return guard([&] {
    TrainNet& t = reinterpret_cast<pidnet_trainer_*>(h)->t;
    if (fwd) *fwd = static_cast<int>(t.b.ops.size() + 1);   // + the batched weight packing
    if (bwd) *bwd = static_cast<int>(t.bops.size() + t.post_ops.size());
  });
}

// ---- post-processing
int pidnet_postprocess(void* stream, const float* logits, int N, int C, int h, int w, int H, int W, unsigned char* pred,
                       const int64_t* labels, int64_t ignore_label, unsigned long long* confusion, unsigned* cell_mask_ws) {
  return guard([&] {
    if (!logits || (!pred && !confusion)) fail("pidnet_postprocess: nothing to compute (null logits, or neither pred nor confusion)");
    if (confusion && !labels) fail("pidnet_postprocess: the confusion matrix needs labels");
    if (N < 1 || h < 1 || w < 1 || H < 1 || W < 1) fail("pidnet_postprocess: empty tensor");
    if (C < 1 || C > 32) fail("pidnet_postprocess supports 1..32 classes");
    cudaError_t e = postprocess_launch(logits, N, C, h, w, H, W, pred, labels, static_cast<long>(ignore_label), confusion,
                                       cell_mask_ws, reinterpret_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) fail(std::string("pidnet_postprocess: ") + cudaGetErrorString(e));
  });
}

// ---- optimizer step
int pidnet_sgd_step(void* stream, float* param, const float* grad, float* momentum_buf, int64_t n, float lr, float momentum,
                    float dampening, float weight_decay, int nesterov, int first_step, float grad_scale) {
  return guard([&] {
    if (!param || !grad || (momentum != 0.f && !momentum_buf)) fail("null argument");
    if (n < 0 || n % 4) fail("pidnet_sgd_step: n must be a non-negative multiple of 4 (pad the flat buffers)");
    if (nesterov && (momentum <= 0.f || dampening != 0.f)) fail("Nesterov momentum requires a momentum and zero dampening");
    cudaError_t e = sgd_step_launch(param, grad, momentum_buf ? momentum_buf : param, n, lr, momentum, dampening, weight_decay,
                                    nesterov, first_step, grad_scale, reinterpret_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) fail(std::string("pidnet_sgd_step: ") + cudaGetErrorString(e));
  });
}

// ---- criterion (FullModel / OhemCrossEntropy / BondaryLoss), see criterion.cu
size_t pidnet_criterion_workspace_bytes(int N, int H, int W) { return criterion_workspace_bytes(N, H, W); }

int pidnet_criterion(void* stream, const float* x_p, const float* x_m, const float* x_d, int N, int C, int h, int w,
                     const int64_t* labels, const float* bd_gt, int H, int W, const float* class_weights,
                     const pidnet_criterion_cfg* cfg, void* workspace, size_t workspace_bytes, float* out12,
                     float* grad_p, float* grad_m, float* grad_d, float* aux_ce_map) {
  return guard([&] {
    if (!x_p || !x_m || !x_d || !labels || !bd_gt || !cfg || !workspace || !out12) fail("null argument");
    if (workspace_bytes < criterion_workspace_bytes(N, H, W)) fail("criterion workspace too small");
    if (C < 1 || C > 32) fail("criterion supports 1..32 classes");
    const bool bwd = grad_p || grad_m || grad_d;
    if (bwd && !(grad_p && grad_m && grad_d)) fail("pass all three gradient buffers or none");
    CritParams p;
    std::memset(&p, 0, sizeof(p));
    p.x_p = x_p; p.x_m = x_m; p.x_d = x_d; p.labels = labels; p.bd_gt = bd_gt; p.class_w = class_weights;
    p.N = N; p.C = C; p.h = h; p.w = w; p.H = H; p.W = W;
    p.ignore_label = cfg->ignore_label;
    p.ohem_thres = cfg->ohem_thres; p.bd_threshold = cfg->bd_threshold;
    p.min_kept = cfg->ohem_keep < 1 ? 1 : cfg->ohem_keep;
    p.bw0 = cfg->balance_weight_aux; p.bw1 = cfg->balance_weight_main; p.sb = cfg->sb_weight;
    p.coeff_bce = cfg->coeff_bce;
    p.out = out12; p.g_p = grad_p; p.g_m = grad_m; p.g_d = grad_d; p.aux_ce = aux_ce_map;
    CK(criterion_launch(p, workspace, bwd, reinterpret_cast<cudaStream_t>(stream)));
  });
}

int pidnet_upsample_align_corners(void* stream, const float* x, int NC, int h, int w, float* out, int H, int W) {
  return guard([&] { CK(upsample_ac_launch(x, NC, h, w, out, H, W, reinterpret_cast<cudaStream_t>(stream))); });
}

#ifdef PIDNET_PROBES
// hardware probe: MN-major operands. a: [64][128] bf16, b: [64][64] bf16 (device), out [128][64] fp32
int pidnet_probe_mn(void* stream, const void* a, const void* b, int lbo, int sbo, float* out) {
  return guard([&] {
    MnProbeParams p;
    std::memset(&p, 0, sizeof(p));
    uint64_t da[2] = {128, 64}, sa[1] = {256};
    uint64_t db[2] = {64, 64}, sb[1] = {128};
    uint32_t box[2] = {64, 64};
    p.tmA = encode_map(a, 2, da, sa, box, 128);
    p.tmB = encode_map(b, 2, db, sb, box, 128);
    p.lbo_a = lbo; p.sbo = sbo; p.out = out;
    CK(mn_probe_launch(p, reinterpret_cast<cudaStream_t>(stream)));
    CK(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)));
  });
}

// hardware probe: cycles for `iters` x 4 back-to-back M128 x N x K16 SS MMAs on `blocks` CTAs (out: device int64[blocks])
int pidnet_probe_mma_rate(void* stream, int N, int iters, int distinct, int blocks, long long* out) {
  return guard([&] {
    CK(mma_rate_launch(N, iters, distinct, blocks, out, reinterpret_cast<cudaStream_t>(stream)));
    CK(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)));
  });
}

// hardware probes for CTA pairs (cta_group::2): a [256][64] bf16, b [64][64] bf16 (device), out [256][64] fp32; rate as above
int pidnet_probe_pair(void* stream, const void* a, const void* b, int swap_b, float* out) {
  return guard([&] {
    PairProbeParams p;
    std::memset(&p, 0, sizeof(p));
    uint64_t da[2] = {64, 256}, sa[1] = {128};
    uint64_t db[2] = {64, 64}, sb[1] = {128};
    uint32_t boxa[2] = {64, 128}, boxb[2] = {64, 32};
    p.tmA = encode_map(a, 2, da, sa, boxa, 128);
    p.tmB = encode_map(b, 2, db, sb, boxb, 128);
    p.swap_b = swap_b; p.out = out;
    CK(pair_probe_launch(p, reinterpret_cast<cudaStream_t>(stream)));
    CK(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)));
  });
}
int pidnet_probe_mma_rate_pair(void* stream, int N, int iters, int distinct, int pairs, long long* out) {
  return guard([&] {
    CK(mma_rate_pair_launch(N, iters, distinct, pairs, out, reinterpret_cast<cudaStream_t>(stream)));
    CK(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)));
  });
}

// hardware probe (see probe.cu): x [18][10][64] bf16, w [64][64] bf16 (device), out [128][64] fp32 (device)
int pidnet_probe_halo(void* stream, const void* x, const void* w, int r, int s, int mode, float* out) {
  return guard([&] {
    ProbeParams p;
    std::memset(&p, 0, sizeof(p));
    uint64_t dims[4] = {64, 10, 18, 1}, str[3] = {128, 1280, 23040};
    uint32_t box[4] = {64, 10, 18, 1};
    p.tmA = encode_map(x, 4, dims, str, box, 128);
    uint64_t d2[2] = {64, 64}, s2[1] = {128};
    uint32_t b2[2] = {64, 64};
    p.tmB = encode_map(w, 2, d2, s2, b2, 128);
    p.r = r; p.s = s; p.mode = mode; p.out = out;
    CK(halo_probe_launch(p, reinterpret_cast<cudaStream_t>(stream)));
    CK(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)));
  });
}

#endif   // PIDNET_PROBES

}  // extern "C"
#pragma GCC visibility pop
