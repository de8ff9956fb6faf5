// HBM-bound fused kernels: 128-bit vectorised NHWC bf16 access, fp32 math.
#include "kernels.cuh"
#include "ptx.cuh"
#include <cstdlib>
#include <cstring>

namespace pidnet {

namespace {

struct F8 {
  float v[8];
};

__device__ __forceinline__ F8 ld8(const bf16* p) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
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

// torch upsample_bilinear2d, align_corners=False: src = max(0, scale*(dst+0.5)-0.5), scale = in/out (fp32)
struct Lerp {
  int i0, i1;
  float l;
};
// Exact unsigned division by a launch-time constant (Granlund-Montgomery round-up): umulhi + 4 ALU ops instead of the
// ~20-instruction runtime division; the elementwise kernels are instruction-issue bound and decode three of them per thread.
struct FastDiv {
  uint32_t mul, sh1, sh2;
};
inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f{0u, 0u, 0u};   // d == 1: q = n
  if (d > 1) {
    const uint32_t l = 32u - static_cast<uint32_t>(__builtin_clz(d - 1));   // ceil(log2 d)
    f.mul = static_cast<uint32_t>((((1ull << l) - d) << 32) / d + 1);
    f.sh1 = 1; f.sh2 = l - 1;
  }
  return f;
}
// host restatement of fdiv() for the CPU test suite (pidnet_debug_fastdiv): exactness of the magic numbers is what every
// elementwise kernel's index decode rests on
uint32_t fastdiv_host(uint32_t n, uint32_t d) {
  const FastDiv f = make_fastdiv(d);
  const uint32_t t = static_cast<uint32_t>((static_cast<uint64_t>(f.mul) * n) >> 32);
  return (t + ((n - t) >> f.sh1)) >> f.sh2;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) {
  const uint32_t t = __umulhi(f.mul, n);
  return (t + ((n - t) >> f.sh1)) >> f.sh2;
}
// launch-time decode constants of an elementwise kernel: divisors (channel groups, W, H or strips) and the two bilinear scales
struct Dec {
  FastDiv g, w, h;
  float sh, sw;   // low-res / hi-res size ratios (torch: scale = in / out in fp32)
};
inline Dec make_dec(int groups, int W, int Hdiv, int in_h, int out_h, int in_w, int out_w) {
  Dec d;
  d.g = make_fastdiv(groups); d.w = make_fastdiv(W); d.h = make_fastdiv(Hdiv);
  d.sh = out_h ? static_cast<float>(in_h) / static_cast<float>(out_h) : 1.f;
  d.sw = out_w ? static_cast<float>(in_w) / static_cast<float>(out_w) : 1.f;
  return d;
}
__device__ __forceinline__ Lerp lerp_of(int dst, int in, float scale) {
  float src = scale * (static_cast<float>(dst) + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  Lerp r;
  r.i0 = static_cast<int>(src);
  if (r.i0 > in - 1) r.i0 = in - 1;
  r.i1 = r.i0 + (r.i0 < in - 1 ? 1 : 0);
  r.l = src - static_cast<float>(r.i0);
  return r;
}

// bilinear sample of 8 channels of a low-res view at hi-res pixel (h, w) of an (H, W) grid
// (32-bit element offsets inside one image: the launchers reject low-res images of 2^31 elements or more)
__device__ __forceinline__ F8 sample8(const View& b, int n, const Lerp& lh, const Lerp& lw, int c) {
  const int ps = static_cast<int>(b.ps);
  const bf16* base = b.ptr + static_cast<long>(n) * (b.H * b.W * ps) + c;
  const int r0 = lh.i0 * b.W, r1 = lh.i1 * b.W;
  const F8 v00 = ld8(base + (r0 + lw.i0) * ps);
  const F8 v01 = ld8(base + (r0 + lw.i1) * ps);
  const F8 v10 = ld8(base + (r1 + lw.i0) * ps);
  const F8 v11 = ld8(base + (r1 + lw.i1) * ps);
  const float w00 = (1.f - lh.l) * (1.f - lw.l), w01 = (1.f - lh.l) * lw.l;
  const float w10 = lh.l * (1.f - lw.l), w11 = lh.l * lw.l;
  F8 r;
#pragma unroll
  for (int e = 0; e < 8; ++e) r.v[e] = w00 * v00.v[e] + w01 * v01.v[e] + w10 * v10.v[e] + w11 * v11.v[e];
  return r;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }
// thread index -> (channel group, w, h, n, pixel) with 32-bit unsigned divisions (three instead of the six 64-bit div/mod
// of the naive form: these elementwise kernels are instruction-issue bound, and 64-bit division was a third of their
// instructions).  Launchers reject tensors with >= 2^32 (pixel, channel-group) items.
__device__ __forceinline__ void decode_idx(unsigned idx, const Dec& d, unsigned groups, unsigned W, unsigned H, int& cg,
                                           int& w, int& h, int& n, unsigned& pix) {
  pix = fdiv(idx, d.g);
  cg = static_cast<int>(idx - pix * groups);
  const unsigned t1 = fdiv(pix, d.w);
  w = static_cast<int>(pix - t1 * W);
  const unsigned nn = fdiv(t1, d.h);
  h = static_cast<int>(t1 - nn * H);
  n = static_cast<int>(nn);
}

// --------------------------------------------------------------------------- vertical strip walk
// The upsampling kernels below give one thread a STRIP of kStrip vertically adjacent hi-res pixels of one (n, w, 8-channel
// group).  The horizontal interpolation weights are fixed for the strip and consecutive rows share their two low-res source
// rows (or advance by one), so the thread keeps the two horizontally-interpolated low-res rows in registers and gathers a
// new one only when the source row changes: 2/scale row gathers per pixel instead of 4 corner gathers (x2: 10 instead of 32
// 16-byte gathers per 8 pixels; x8: 4 instead of 32), and the index arithmetic is paid once per strip.  Consecutive threads
// still map to consecutive (w, channel group), so every hi-res access stays a fully coalesced 128-bit access.
constexpr int kStrip = 8;

struct StripIdx {
  int cg, w, n, h0;
  bool valid;
};
// item index -> (channel group, w, strip, n); invalid tail threads are redirected to the last item (they still take part
// in warp shuffles) and must not store
// (d.h divides by the number of strips per image)
__device__ __forceinline__ StripIdx strip_decode(unsigned gid, const Dec& d, unsigned groups, unsigned W, unsigned H, unsigned N) {
  const unsigned strips = (H + kStrip - 1) / kStrip;
  const unsigned total = N * strips * W * groups;
  StripIdx r;
  r.valid = gid < total;
  if (!r.valid) gid = total - groups + gid % groups;
  const unsigned t0 = fdiv(gid, d.g);
  r.cg = static_cast<int>(gid - t0 * groups);
  const unsigned t1 = fdiv(t0, d.w);
  r.w = static_cast<int>(t0 - t1 * W);
  const unsigned nn = fdiv(t1, d.h);
  r.h0 = static_cast<int>(t1 - nn * strips) * kStrip;
  r.n = static_cast<int>(nn);
  return r;
}
// horizontally interpolated 8 channels of low-res row `row` (o0 / o1: element offsets of the two source columns)
__device__ __forceinline__ F8 hsample8(const bf16* img, int row, int Wl, long ps, long o0, long o1, float l) {
  const bf16* r = img + row * (Wl * static_cast<int>(ps));
  const F8 a = ld8(r + o0), b = ld8(r + o1);
  F8 o;
#pragma unroll
  for (int e = 0; e < 8; ++e) o.v[e] = (1.f - l) * a.v[e] + l * b.v[e];
  return o;
}
__device__ __forceinline__ F8 vlerp8(const F8& a, const F8& b, float l) {
  F8 o;
#pragma unroll
  for (int e = 0; e < 8; ++e) o.v[e] = (1.f - l) * a.v[e] + l * b.v[e];
  return o;
}

// --------------------------------------------------------------------------- stem
// Register-tiled direct conv: one thread = 2 horizontally adjacent output pixels x 32 output channels
// (1728 FMAs per 45 input loads + 216 broadcast LDS.128 of the weights), fp32 math on the fp32 image.
template <int CH>  // output channels per thread (32; 8 for tiny test configs)
__global__ void __launch_bounds__(128) stem_conv_kernel(const float* __restrict__ x, int N, int H, int W, View out,
                                                        const float* __restrict__ w, const float* __restrict__ bias) {
  __shared__ __align__(16) float ws[27 * CH + CH];
  const int Cout = out.C;
  const int c0 = blockIdx.y * CH;  // this block's channel slice
  for (int i = threadIdx.x; i < 27 * CH; i += blockDim.x) ws[i] = w[(i / CH) * Cout + c0 + (i % CH)];
  if (threadIdx.x < CH) ws[27 * CH + threadIdx.x] = bias[c0 + threadIdx.x];
  __syncthreads();
  const int Wp = (out.W + 1) >> 1;  // pixel pairs per output row
  const long total = static_cast<long>(N) * out.H * Wp;
  const long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int pw = static_cast<int>(idx % Wp);
  const long t1 = idx / Wp;
  const int oh = static_cast<int>(t1 % out.H);
  const int n = static_cast<int>(t1 / out.H);
  const int ow = pw * 2;
  float acc0[CH], acc1[CH];
#pragma unroll
  for (int e = 0; e < CH; ++e) acc0[e] = acc1[e] = ws[27 * CH + e];
  const float* xn = x + static_cast<long>(n) * 3 * H * W;
  const int iw0 = ow * 2 - 1;
#pragma unroll
  for (int ci = 0; ci < 3; ++ci) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int ih = oh * 2 - 1 + r;
      float v[5];
      const bool rowok = ih >= 0 && ih < H;
      const float* row = xn + (static_cast<long>(ci) * H + (rowok ? ih : 0)) * W;
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const int iw = iw0 + j;
        v[j] = (rowok && iw >= 0 && iw < W) ? __ldg(row + iw) : 0.f;
      }
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const float4* wp = reinterpret_cast<const float4*>(ws + ((ci * 3 + r) * 3 + q) * CH);
        const float a0 = v[q], a1 = v[q + 2];
#pragma unroll
        for (int e4 = 0; e4 < CH / 4; ++e4) {
          const float4 wv = wp[e4];
          acc0[e4 * 4 + 0] += a0 * wv.x; acc0[e4 * 4 + 1] += a0 * wv.y;
          acc0[e4 * 4 + 2] += a0 * wv.z; acc0[e4 * 4 + 3] += a0 * wv.w;
          acc1[e4 * 4 + 0] += a1 * wv.x; acc1[e4 * 4 + 1] += a1 * wv.y;
          acc1[e4 * 4 + 2] += a1 * wv.z; acc1[e4 * 4 + 3] += a1 * wv.w;
        }
      }
    }
  }
  bf16* o0 = out.ptr + ((static_cast<long>(n) * out.H + oh) * out.W + ow) * out.ps + c0;
#pragma unroll
  for (int g = 0; g < CH / 8; ++g) {
    F8 f;
#pragma unroll
    for (int e = 0; e < 8; ++e) f.v[e] = fmaxf(acc0[g * 8 + e], 0.f);
    st8(o0 + g * 8, f);
  }
  if (ow + 1 < out.W) {
#pragma unroll
    for (int g = 0; g < CH / 8; ++g) {
      F8 f;
#pragma unroll
      for (int e = 0; e < 8; ++e) f.v[e] = fmaxf(acc1[g * 8 + e], 0.f);
      st8(o0 + out.ps + g * 8, f);
    }
  }
}

// --------------------------------------------------------------------------- upadd / affine
__global__ void __launch_bounds__(256) upadd_kernel(View a, View b, View r, View out, const float* __restrict__ s,
                                                    const float* __restrict__ t, int relu, Dec dec) {
  pdl_wait();
  pdl_launch_dependents();
  const int groups = out.C >> 3;
  const unsigned total = static_cast<unsigned>(out.N) * out.H * out.W * groups;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int cg, w, h, n;
  unsigned pixu;
  decode_idx(idx, dec, groups, out.W, out.H, cg, w, h, n, pixu);
  const long pix = pixu;
  F8 v;
#pragma unroll
  for (int e = 0; e < 8; ++e) v.v[e] = 0.f;
  if (a.ptr) v = ld8(a.ptr + pix * a.ps + cg * 8);
  if (b.ptr) {
    const Lerp lh = lerp_of(h, b.H, dec.sh), lw = lerp_of(w, b.W, dec.sw);
    const F8 u = sample8(b, n, lh, lw, cg * 8);
#pragma unroll
    for (int e = 0; e < 8; ++e) v.v[e] += u.v[e];
  }
  if (s) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v.v[e] = v.v[e] * __ldg(s + cg * 8 + e) + __ldg(t + cg * 8 + e);
  }
  if (r.ptr) {   // residual added after the affine map (BatchNorm -> + identity -> ReLU)
    const F8 rv = ld8(r.ptr + pix * r.ps + cg * 8);
#pragma unroll
    for (int e = 0; e < 8; ++e) v.v[e] += rv.v[e];
  }
  if (relu) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v.v[e] = fmaxf(v.v[e], 0.f);
  }
  st8(out.ptr + pix * out.ps + cg * 8, v);
}

// strip form of upadd_kernel for the upsampling case (b != null): out = act(s * (a + U(b)) + t (+ r))
__device__ __forceinline__ F8 cvt8(const uint4& u) {
  F8 r;
  r.v[0] = __uint_as_float(u.x << 16); r.v[1] = __uint_as_float(u.x & 0xFFFF0000u);
  r.v[2] = __uint_as_float(u.y << 16); r.v[3] = __uint_as_float(u.y & 0xFFFF0000u);
  r.v[4] = __uint_as_float(u.z << 16); r.v[5] = __uint_as_float(u.z & 0xFFFF0000u);
  r.v[6] = __uint_as_float(u.w << 16); r.v[7] = __uint_as_float(u.w & 0xFFFF0000u);
  return r;
}
// the two low-res source rows of hi-res row h, kept across the strip (see "vertical strip walk")
struct RowPair {
  F8 a, b;
  int ra = -1, rb = -1;
  template <class Load>
  __device__ __forceinline__ void advance(const Lerp& lh, Load&& load) {
    if (lh.i0 != ra) {
      if (lh.i0 == rb) a = b; else a = load(lh.i0);
      ra = lh.i0;
    }
    if (lh.i1 != rb) {
      if (lh.i1 == ra) b = a; else b = load(lh.i1);
      rb = lh.i1;
    }
  }
};
__device__ __forceinline__ void upadd_strip_body(const View& a, const View& b, const View& r, const View& out,
                                                 const float* __restrict__ s, const float* __restrict__ t, int relu,
                                                 const Dec& dec) {
  const StripIdx ix = strip_decode(blockIdx.x * blockDim.x + threadIdx.x, dec, out.C >> 3, out.W, out.H, out.N);
  if (!ix.valid) return;
  const int cg = ix.cg;
  const Lerp lw = lerp_of(ix.w, b.W, dec.sw);
  const bf16* limg = b.ptr + static_cast<long>(ix.n) * b.H * b.W * b.ps;
  const long o0 = static_cast<long>(lw.i0) * b.ps + cg * 8, o1 = static_cast<long>(lw.i1) * b.ps + cg * 8;
  const long pixn = static_cast<long>(ix.n) * out.H * out.W + ix.w;
  const int rows = min(kStrip, out.H - ix.h0);
  uint4 ar[kStrip], rr[kStrip];
#pragma unroll
  for (int k = 0; k < kStrip; ++k) {
    if (k < rows) {
      const long pix = pixn + static_cast<long>(ix.h0 + k) * out.W;
      if (a.ptr) ar[k] = __ldg(reinterpret_cast<const uint4*>(a.ptr + pix * a.ps + cg * 8));
      if (r.ptr) rr[k] = __ldg(reinterpret_cast<const uint4*>(r.ptr + pix * r.ps + cg * 8));
    }
  }
  F8 sc, sh;
  if (s) {
#pragma unroll
    for (int e = 0; e < 8; ++e) { sc.v[e] = __ldg(s + cg * 8 + e); sh.v[e] = __ldg(t + cg * 8 + e); }
  }
  RowPair rp;
#pragma unroll
  for (int k = 0; k < kStrip; ++k) {
    if (k < rows) {
      const int h = ix.h0 + k;
      const Lerp lh = lerp_of(h, b.H, dec.sh);
      rp.advance(lh, [&](int row) { return hsample8(limg, row, b.W, b.ps, o0, o1, lw.l); });
      F8 v = vlerp8(rp.a, rp.b, lh.l);
      if (a.ptr) {
        const F8 av = cvt8(ar[k]);
#pragma unroll
        for (int e = 0; e < 8; ++e) v.v[e] += av.v[e];
      }
      if (s) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v.v[e] = v.v[e] * sc.v[e] + sh.v[e];
      }
      if (r.ptr) {
        const F8 rv = cvt8(rr[k]);
#pragma unroll
        for (int e = 0; e < 8; ++e) v.v[e] += rv.v[e];
      }
      if (relu) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v.v[e] = fmaxf(v.v[e], 0.f);
      }
      st8(out.ptr + (pixn + static_cast<long>(h) * out.W) * out.ps + cg * 8, v);
    }
  }
}
__global__ void __launch_bounds__(256) upadd_strip_kernel(View a, View b, View r, View out, const float* __restrict__ s,
                                                          const float* __restrict__ t, int relu, Dec dec) {
  pdl_wait();
  pdl_launch_dependents();
  upadd_strip_body(a, b, r, out, s, t, relu, dec);
}
// up to kUpaddJobs independent jobs with the same output geometry in one launch (blockIdx.y = job): the PAPPM branch adds
// (a job whose `b` has the output's own size is a plain affine + ReLU: identity interpolation is exact)
constexpr int kUpaddJobs = 5;
struct UpaddBatch {
  View a[kUpaddJobs], b[kUpaddJobs], out[kUpaddJobs];
  const float* s[kUpaddJobs];
  const float* t[kUpaddJobs];
  Dec dec[kUpaddJobs];
  int relu;
};
__global__ void __launch_bounds__(256) upadd_strip_batch_kernel(const __grid_constant__ UpaddBatch p) {
  pdl_wait();
  pdl_launch_dependents();
  const int j = blockIdx.y;
  upadd_strip_body(p.a[j], p.b[j], View{nullptr, 0, 0, 0, 0, 0}, p.out[j], p.s[j], p.t[j], p.relu, p.dec[j]);
}

// --------------------------------------------------------------------------- avg pool + affine
// one block per output pixel; thread = (channel group, window split); smem reduce over the splits
__global__ void __launch_bounds__(256) pool_affine_kernel(View x, View out, int k, int stride, int pad,
                                                          const float* __restrict__ s, const float* __restrict__ t,
                                                          int relu) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ float red[];  // [splits][groups*8]
  const int groups = x.C >> 3;
  const int splits = blockDim.x / groups;
  const int cg = threadIdx.x % groups;
  const int sp = threadIdx.x / groups;
  const int opix = blockIdx.x;
  const int ow = opix % out.W;
  const int oh = (opix / out.W) % out.H;
  const int n = opix / (out.W * out.H);
  int h0, h1, w0, w1;
  float div;
  if (k == 0) {
    h0 = 0; h1 = x.H; w0 = 0; w1 = x.W;
    div = static_cast<float>(x.H * x.W);
  } else {
    h0 = oh * stride - pad; w0 = ow * stride - pad;
    h1 = min(h0 + k, x.H); w1 = min(w0 + k, x.W);
    h0 = max(h0, 0); w0 = max(w0, 0);
    div = static_cast<float>(k * k);  // count_include_pad=True: windows never leave the padded image here
  }
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (sp < splits) {
    const int ww = w1 - w0;
    const int cnt = (h1 - h0) * ww;
    const bf16* base = x.ptr + static_cast<long>(n) * x.H * x.W * x.ps + cg * 8;
    for (int i = sp; i < cnt; i += splits) {
      const int ih = h0 + i / ww, iw = w0 + i % ww;
      const F8 v = ld8(base + (static_cast<long>(ih) * x.W + iw) * x.ps);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += v.v[e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) red[(sp * groups + cg) * 8 + e] = acc[e];
  }
  __syncthreads();
  if (sp == 0) {
    F8 o;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = 0.f;
      for (int q = 0; q < splits; ++q) v += red[(q * groups + cg) * 8 + e];
      v = v / div;
      if (s) v = v * __ldg(s + cg * 8 + e) + __ldg(t + cg * 8 + e);
      o.v[e] = relu ? fmaxf(v, 0.f) : v;
    }
    st8(out.ptr + static_cast<long>(opix) * out.ps + cg * 8, o);
  }
}

// --------------------------------------------------------------------------- pooled pyramid (see kernels.cuh)
struct PyramidParams {
  View x;
  View out[4];
  const float* s;   // [4][C]
  const float* t;
  View ew[2];       // optional full-resolution by-products relu(se[j] * x + te[j]) (PAPPM scale0 / shortcut operands)
  const float* se;  // [2][C]
  const float* te;
};
__global__ void __launch_bounds__(256) pool_pyramid_kernel(PyramidParams p) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ float sat[];   // [(H+1)][(W+1)][32]: sat[h][w][c] = sum of x[<h][<w][c]
  const int H = p.x.H, W = p.x.W, W1 = W + 1;
  const int groups = p.x.C >> 5;              // 32-channel slices
  const int n = blockIdx.x / groups, c0 = (blockIdx.x - n * groups) * 32;
  const bf16* xin = p.x.ptr + static_cast<long>(n) * H * W * p.x.ps + c0;
  // zero border, then the pixels (thread = pixel x 8-channel group)
  for (int i = threadIdx.x; i < (H + 1 + W) * 32; i += 256) {
    const int c = i & 31, j = i >> 5;          // j < W1: row 0; else column 0 of row j - W
    if (j < W1) sat[j * 32 + c] = 0.f;
    else sat[static_cast<long>(j - W) * W1 * 32 + c] = 0.f;
  }
  for (int i = threadIdx.x; i < H * W * 4; i += 256) {
    const int g = i & 3, pix = i >> 2;
    const int h = pix / W, w = pix - h * W;
    const F8 v = ld8(xin + static_cast<long>(pix) * p.x.ps + g * 8);
    float* d = sat + (static_cast<long>(h + 1) * W1 + w + 1) * 32 + g * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) d[e] = v.v[e];
#pragma unroll
    for (int j = 0; j < 2; ++j)
      if (p.ew[j].ptr) {
        const float* sc = p.se + j * p.x.C + c0 + g * 8;
        const float* sh = p.te + j * p.x.C + c0 + g * 8;
        F8 o;
#pragma unroll
        for (int e = 0; e < 8; ++e) o.v[e] = fmaxf(v.v[e] * __ldg(sc + e) + __ldg(sh + e), 0.f);
        st8(p.ew[j].ptr + (static_cast<long>(n) * H * W + pix) * p.ew[j].ps + c0 + g * 8, o);
      }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H * 32; i += 256) {   // prefix along w
    const int c = i & 31, h = i >> 5;
    float* r = sat + static_cast<long>(h + 1) * W1 * 32 + c;
    float a = 0.f;
    for (int w = 1; w <= W; ++w) { a += r[w * 32]; r[w * 32] = a; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < W * 32; i += 256) {   // prefix along h
    const int c = i & 31, w = i >> 5;
    float* r = sat + static_cast<long>(w + 1) * 32 + c;
    float a = 0.f;
    for (int h = 1; h <= H; ++h) { a += r[static_cast<long>(h) * W1 * 32]; r[static_cast<long>(h) * W1 * 32] = a; }
  }
  __syncthreads();
  auto box = [&](int h0, int h1, int w0, int w1, int c) {
    return sat[(static_cast<long>(h1) * W1 + w1) * 32 + c] - sat[(static_cast<long>(h0) * W1 + w1) * 32 + c] -
           sat[(static_cast<long>(h1) * W1 + w0) * 32 + c] + sat[(static_cast<long>(h0) * W1 + w0) * 32 + c];
  };
#pragma unroll
  for (int lv = 0; lv < 4; ++lv) {
    const View& o = p.out[lv];
    const int k = lv == 0 ? 5 : (lv == 1 ? 9 : 17), stride = lv == 0 ? 2 : (lv == 1 ? 4 : 8), pad = stride;
    const float* sc = p.s + lv * p.x.C + c0;
    const float* sh = p.t + lv * p.x.C + c0;
    bf16* ob = o.ptr + static_cast<long>(n) * o.H * o.W * o.ps + c0;
    for (int i = threadIdx.x; i < o.H * o.W * 32; i += 256) {
      const int c = i & 31, op = i >> 5;
      const int oh = op / o.W, ow = op - oh * o.W;
      float v;
      if (lv == 3) {
        v = box(0, H, 0, W, c) / static_cast<float>(H * W);
      } else {   // count_include_pad=True: the divisor is k*k, the sum runs over the in-image part of the window
        const int h0 = max(oh * stride - pad, 0), w0 = max(ow * stride - pad, 0);
        const int h1 = min(oh * stride - pad + k, H), w1 = min(ow * stride - pad + k, W);
        v = box(h0, h1, w0, w1, c) / static_cast<float>(k * k);
      }
      v = fmaxf(v * __ldg(sc + c) + __ldg(sh + c), 0.f);
      ob[static_cast<long>(op) * o.ps + c] = __float2bfloat16_rn(v);
    }
  }
}

// --------------------------------------------------------------------------- PagFM fuse, Light_Bag / Bag
// One thread per pixel x 8 channels.  (Strip-walk forms of these three kernels -- as upadd_strip_kernel above -- were
// measured and dropped: fewer instructions, but 152 registers left 12 % occupancy: PagFM 0.20 vs 0.15 ms, Light_Bag 0.31
// vs 0.25 ms, Bag 0.68 ms on PIDNet-L.)
template <int LP>  // lanes per pixel = C/8
__global__ void __launch_bounds__(256) pag_fuse_flat_kernel(View x, View low, View out, int relu, Dec dec) {
  pdl_wait();
  pdl_launch_dependents();
  const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned npix = static_cast<unsigned>(x.N) * x.H * x.W;
  const bool valid = gid / LP < npix;
  int cg, w, h, n;
  unsigned pixu;
  decode_idx(valid ? gid : (npix - 1) * LP + gid % LP, dec, LP, x.W, x.H, cg, w, h, n, pixu);   // LP is a compile-time power of two
  const long pix = pixu;
  const int C = x.C;
  const Lerp lh = lerp_of(h, low.H, dec.sh), lw = lerp_of(w, low.W, dec.sw);
  const F8 xv = ld8(x.ptr + pix * x.ps + cg * 8);
  const F8 yv = sample8(low, n, lh, lw, cg * 8);
  const F8 zv = sample8(low, n, lh, lw, C + cg * 8);
  float dot = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) dot += xv.v[e] * zv.v[e];
#pragma unroll
  for (int o = LP / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  // scalar term t (channel 2C of `low`)
  const bf16* tb = low.ptr + static_cast<long>(n) * low.H * low.W * low.ps + 2 * C;
  const float t00 = __bfloat162float(tb[(static_cast<long>(lh.i0) * low.W + lw.i0) * low.ps]);
  const float t01 = __bfloat162float(tb[(static_cast<long>(lh.i0) * low.W + lw.i1) * low.ps]);
  const float t10 = __bfloat162float(tb[(static_cast<long>(lh.i1) * low.W + lw.i0) * low.ps]);
  const float t11 = __bfloat162float(tb[(static_cast<long>(lh.i1) * low.W + lw.i1) * low.ps]);
  const float tt = (1.f - lh.l) * ((1.f - lw.l) * t00 + lw.l * t01) + lh.l * ((1.f - lw.l) * t10 + lw.l * t11);
  const float g = sigmoidf_(dot + tt);
  F8 o;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    float v = (1.f - g) * xv.v[e] + g * yv.v[e];
    o.v[e] = relu ? fmaxf(v, 0.f) : v;
  }
  if (valid) st8(out.ptr + pix * out.ps + cg * 8, o);
}

// ---- wide form: 16 channels per thread, mixed-precision FMAs
// fma.rn.f32.bf16 (sm_100: FHFMA.BF16) multiplies two bf16 values and accumulates in fp32, reading either half of a 32-bit
// register directly: the loaded vectors are never unpacked.  The bilinear weights must then be bf16 values -- exact when the
// low-res field is upsampled by 2, 4 or 8 (the weights and their pairwise products are dyadic rationals of at most 8 significant
// bits), which is every PagFM of the network at power-of-two input sizes; other geometries keep the fp32-weight kernel above.
// Products bf16 x bf16 are exact in fp32, so the result differs from the unpacked form only by summation order.
// The flat form was instruction-issue bound (74 % issue-active at 35 % of HBM): ~270 instructions per 8 channels, a quarter of
// them unpacking, another fifth the per-thread index decode that 16 channels per thread now amortise twice as far.
__device__ __forceinline__ float fhfma_lo(uint32_t a, uint32_t b_lo, float c) {   // a.lo * b.lo + c
  float d;
  asm("{\n\t.reg .b16 al, ah, bl, bh;\n\tmov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\tfma.rn.f32.bf16 %0, al, bl, %3;\n\t}"
      : "=f"(d) : "r"(a), "r"(b_lo), "f"(c));
  return d;
}
__device__ __forceinline__ float fhfma_hi(uint32_t a, uint32_t b_lo, float c) {   // a.hi * b.lo + c
  float d;
  asm("{\n\t.reg .b16 al, ah, bl, bh;\n\tmov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\tfma.rn.f32.bf16 %0, ah, bl, %3;\n\t}"
      : "=f"(d) : "r"(a), "r"(b_lo), "f"(c));
  return d;
}
__device__ __forceinline__ float fhfma_ll(uint32_t a, uint32_t b, float c) { return fhfma_lo(a, b, c); }
__device__ __forceinline__ float fhfma_hh(uint32_t a, uint32_t b, float c) {      // a.hi * b.hi + c
  float d;
  asm("{\n\t.reg .b16 al, ah, bl, bh;\n\tmov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\tfma.rn.f32.bf16 %0, ah, bh, %3;\n\t}"
      : "=f"(d) : "r"(a), "r"(b), "f"(c));
  return d;
}
__device__ __forceinline__ uint32_t bf16_bits(float w) {   // bf16(w) in the low half (w is exactly representable)
  return __float_as_uint(w) >> 16;
}
// acc[e] += w * v[e] for the 8 bf16 of v (w: bf16 bits in the low half)
__device__ __forceinline__ void axpy8(float (&acc)[8], const uint4& v, uint32_t w) {
  acc[0] = fhfma_lo(v.x, w, acc[0]); acc[1] = fhfma_hi(v.x, w, acc[1]);
  acc[2] = fhfma_lo(v.y, w, acc[2]); acc[3] = fhfma_hi(v.y, w, acc[3]);
  acc[4] = fhfma_lo(v.z, w, acc[4]); acc[5] = fhfma_hi(v.z, w, acc[5]);
  acc[6] = fhfma_lo(v.w, w, acc[6]); acc[7] = fhfma_hi(v.w, w, acc[7]);
}
// d += sum_e a[e] * b[e] over the 8 bf16 pairs (two partial sums to shorten the dependency chain)
__device__ __forceinline__ float dot8(const uint4& a, const uint4& b, float d) {
  float d0 = fhfma_ll(a.x, b.x, d), d1 = fhfma_hh(a.x, b.x, 0.f);
  d0 = fhfma_ll(a.y, b.y, d0); d1 = fhfma_hh(a.y, b.y, d1);
  d0 = fhfma_ll(a.z, b.z, d0); d1 = fhfma_hh(a.z, b.z, d1);
  d0 = fhfma_ll(a.w, b.w, d0); d1 = fhfma_hh(a.w, b.w, d1);
  return d0 + d1;
}
__device__ __forceinline__ uint32_t blend_pack(uint32_t xw, float y0, float y1, float g, int relu) {
  const float x0 = __uint_as_float(xw << 16), x1 = __uint_as_float(xw & 0xFFFF0000u);
  const float o0 = fmaf(g, y0 - x0, x0), o1 = fmaf(g, y1 - x1, x1);
  uint32_t r;
  if (relu) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(o1), "f"(o0));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(o1), "f"(o0));
  return r;
}
template <int LP>  // lanes per pixel = C/16
__global__ void __launch_bounds__(256) pag_fuse_wide_kernel(View x, View low, View out, int relu, Dec dec) {
  pdl_wait();
  pdl_launch_dependents();
  const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned npix = static_cast<unsigned>(x.N) * x.H * x.W;
  const bool valid = gid / LP < npix;
  int cg, w, h, n;
  unsigned pixu;
  decode_idx(valid ? gid : (npix - 1) * LP + gid % LP, dec, LP, x.W, x.H, cg, w, h, n, pixu);
  const int C = x.C, c0 = cg * 16;
  const Lerp lh = lerp_of(h, low.H, dec.sh), lw = lerp_of(w, low.W, dec.sw);
  const float w00 = (1.f - lh.l) * (1.f - lw.l), w01 = (1.f - lh.l) * lw.l, w10 = lh.l * (1.f - lw.l), w11 = lh.l * lw.l;
  const uint32_t b00 = bf16_bits(w00), b01 = bf16_bits(w01), b10 = bf16_bits(w10), b11 = bf16_bits(w11);
  const int ps = static_cast<int>(low.ps);
  const bf16* lb = low.ptr + static_cast<long>(n) * (low.H * low.W * ps);
  const int o00 = (lh.i0 * low.W + lw.i0) * ps, o01 = (lh.i0 * low.W + lw.i1) * ps;
  const int o10 = (lh.i1 * low.W + lw.i0) * ps, o11 = (lh.i1 * low.W + lw.i1) * ps;
  const bf16* xp = x.ptr + static_cast<long>(pixu) * x.ps + c0;
  uint4 xv[2];
  float y[2][8];
  float d00 = 0.f, d01 = 0.f, d10 = 0.f, d11 = 0.f;
#pragma unroll
  for (int v = 0; v < 2; ++v) {
    const int c = c0 + v * 8;
    xv[v] = __ldg(reinterpret_cast<const uint4*>(xp + v * 8));
    const uint4 y00 = __ldg(reinterpret_cast<const uint4*>(lb + o00 + c)), y01 = __ldg(reinterpret_cast<const uint4*>(lb + o01 + c));
    const uint4 y10 = __ldg(reinterpret_cast<const uint4*>(lb + o10 + c)), y11 = __ldg(reinterpret_cast<const uint4*>(lb + o11 + c));
    const uint4 z00 = __ldg(reinterpret_cast<const uint4*>(lb + o00 + C + c)), z01 = __ldg(reinterpret_cast<const uint4*>(lb + o01 + C + c));
    const uint4 z10 = __ldg(reinterpret_cast<const uint4*>(lb + o10 + C + c)), z11 = __ldg(reinterpret_cast<const uint4*>(lb + o11 + C + c));
#pragma unroll
    for (int e = 0; e < 8; ++e) y[v][e] = 0.f;
    axpy8(y[v], y00, b00); axpy8(y[v], y01, b01); axpy8(y[v], y10, b10); axpy8(y[v], y11, b11);
    d00 = dot8(xv[v], z00, d00); d01 = dot8(xv[v], z01, d01); d10 = dot8(xv[v], z10, d10); d11 = dot8(xv[v], z11, d11);
  }
  float dot = w00 * d00 + w01 * d01 + w10 * d10 + w11 * d11;
#pragma unroll
  for (int o = LP / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  // scalar term t (channel 2C of `low`)
  const bf16* tb = lb + 2 * C;
  const float tt = w00 * __bfloat162float(tb[o00]) + w01 * __bfloat162float(tb[o01]) + w10 * __bfloat162float(tb[o10]) +
                   w11 * __bfloat162float(tb[o11]);
  const float g = sigmoidf_(dot + tt);
  if (valid) {
    bf16* op = out.ptr + static_cast<long>(pixu) * out.ps + c0;
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      uint4 o;
      o.x = blend_pack(xv[v].x, y[v][0], y[v][1], g, relu); o.y = blend_pack(xv[v].y, y[v][2], y[v][3], g, relu);
      o.z = blend_pack(xv[v].z, y[v][4], y[v][5], g, relu); o.w = blend_pack(xv[v].w, y[v][6], y[v][7], g, relu);
      *reinterpret_cast<uint4*>(op + v * 8) = o;
    }
  }
}

__global__ void __launch_bounds__(256) lightbag_uv_flat_kernel(View p, View il, View d, View out, Dec dec) {
  pdl_wait();
  pdl_launch_dependents();
  const int groups = p.C >> 3;
  const long total = static_cast<long>(p.N) * p.H * p.W * groups;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int cg, w, h, n;
  unsigned pixu;
  decode_idx(idx, dec, groups, p.W, p.H, cg, w, h, n, pixu);
  const long pix = pixu;
  const Lerp lh = lerp_of(h, il.H, dec.sh), lw = lerp_of(w, il.W, dec.sw);
  const F8 iv = sample8(il, n, lh, lw, cg * 8);
  const F8 pv = ld8(p.ptr + pix * p.ps + cg * 8);
  const F8 dv = ld8(d.ptr + pix * d.ps + cg * 8);
  F8 u, v;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float g = sigmoidf_(dv.v[e]);
    u.v[e] = (1.f - g) * iv.v[e] + pv.v[e];
    v.v[e] = iv.v[e] + g * pv.v[e];
  }
  st8(out.ptr + pix * out.ps + cg * 8, u);
  st8(out.ptr + pix * out.ps + p.C + cg * 8, v);
}


__global__ void __launch_bounds__(256) bag_blend_flat_kernel(View p, View il, View d, View out, const float* __restrict__ s,
                                                             const float* __restrict__ t, Dec dec) {
  pdl_wait();
  pdl_launch_dependents();
  const int groups = p.C >> 3;
  const long total = static_cast<long>(p.N) * p.H * p.W * groups;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int cg, w, h, n;
  unsigned pixu;
  decode_idx(idx, dec, groups, p.W, p.H, cg, w, h, n, pixu);
  const long pix = pixu;
  const Lerp lh = lerp_of(h, il.H, dec.sh), lw = lerp_of(w, il.W, dec.sw);
  const F8 iv = sample8(il, n, lh, lw, cg * 8);
  const F8 pv = ld8(p.ptr + pix * p.ps + cg * 8);
  const F8 dv = ld8(d.ptr + pix * d.ps + cg * 8);
  F8 o;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float g = sigmoidf_(dv.v[e]);
    const float a = g * pv.v[e] + (1.f - g) * iv.v[e];
    o.v[e] = fmaxf(a * __ldg(s + cg * 8 + e) + __ldg(t + cg * 8 + e), 0.f);
  }
  st8(out.ptr + pix * out.ps + cg * 8, o);
}

// --------------------------------------------------------------------------- SIMT reference conv
__global__ void __launch_bounds__(128) conv_ref_kernel(const ConvRefParams p) {
  const long total = static_cast<long>(p.N) * p.Ho * p.Wo * p.Cout;
  const long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int co = static_cast<int>(idx % p.Cout);
  long pix = idx / p.Cout;
  const int ow = static_cast<int>(pix % p.Wo);
  pix /= p.Wo;
  const int oh = static_cast<int>(pix % p.Ho);
  const int n = static_cast<int>(pix / p.Ho);
  float acc = 0.f;
  const bf16* wrow = p.wpk + static_cast<long>(co) * p.Ktot;
  long kbase = 0;
  for (int s = 0; s < p.nsrc; ++s) {
    const ConvSrc& src = p.src[s];
    for (int t = 0; t < src.ntaps; ++t) {
      const uint32_t tap = src.taps[t];
      const RefMap& m = p.maps[tap & 0xFF];
      const int ih = oh + static_cast<int>((tap >> 8) & 0xFF) - 8;
      const int iw = ow + static_cast<int>((tap >> 16) & 0xFF) - 8;
      const bool in = ih >= 0 && ih < m.H && iw >= 0 && iw < m.W;
      if (in) {
        const bf16* a = m.ptr + n * m.sN + ih * m.sH + iw * m.sW;
        for (int c = 0; c < m.C; ++c) acc += __bfloat162float(a[c]) * __bfloat162float(wrow[kbase + c]);
      }
      kbase += static_cast<long>(src.chunks) * p.BK;
    }
  }
  acc += p.bias[co];
  if (p.res.ptr) acc += __bfloat162float(p.res.ptr[n * p.res.sN + oh * p.res.sH + ow * p.res.sW + co]);
  if (p.relu) acc = fmaxf(acc, 0.f);
  if (p.out_mode == kOutNHWCbf16)
    p.out[n * p.o_sN + oh * p.o_sH + ow * p.o_sW + co] = __float2bfloat16_rn(acc);
  else
    p.out_f32[((static_cast<long>(n) * p.Cout + co) * p.Ho + oh) * p.Wo + ow] = acc;
}

inline unsigned blocks_for(long total, int threads) { return static_cast<unsigned>((total + threads - 1) / threads); }

}  // namespace

uint32_t fastdiv_debug(uint32_t n, uint32_t d) { return fastdiv_host(n, d); }

cudaError_t stem_conv_launch(const float* x, int N, int H, int W, View out, const float* w, const float* bias,
                             cudaStream_t st) {
  if (out.C % 8 != 0) return cudaErrorInvalidValue;
  const long total = static_cast<long>(N) * out.H * ((out.W + 1) / 2);
  if (out.C % 32 == 0) {
    dim3 grid(blocks_for(total, 128), out.C / 32, 1);
    stem_conv_kernel<32><<<grid, 128, 0, st>>>(x, N, H, W, out, w, bias);
  } else {
    dim3 grid(blocks_for(total, 128), out.C / 8, 1);
    stem_conv_kernel<8><<<grid, 128, 0, st>>>(x, N, H, W, out, w, bias);
  }
  return cudaGetLastError();
}

// up-sampling factor 2, 4 or 8 in both directions: every bilinear weight (and product of two) is a bf16 value
static bool pow2_upsample(int in, int out) { return out == 2 * in || out == 4 * in || out == 8 * in; }
cudaError_t pag_fuse_launch(View x, View low, View out, int relu, cudaStream_t st) {
  if (static_cast<long>(x.N) * x.H * x.W * (x.C / 8) + 256 >= (1L << 32)) return cudaErrorInvalidValue;   // 32-bit index arithmetic
  if (static_cast<long>(low.H) * low.W * low.ps >= (1L << 31)) return cudaErrorInvalidValue;   // 32-bit offsets inside an image
  static const bool wide_ok = [] { const char* v = std::getenv("PIDNET_PAG_WIDE"); return !(v && v[0] == '0'); }();
  if (wide_ok && x.C % 16 == 0 && pow2_upsample(low.H, x.H) && pow2_upsample(low.W, x.W)) {
    const int LP = x.C / 16;
    const Dec dec = make_dec(LP, x.W, x.H, low.H, x.H, low.W, x.W);
    const dim3 grid(blocks_for(static_cast<long>(x.N) * x.H * x.W * LP, 256), 1, 1), block(256, 1, 1);
    switch (LP) {
      case 1: return launch_pdl(pag_fuse_wide_kernel<1>, grid, block, 0, st, x, low, out, relu, dec);
      case 2: return launch_pdl(pag_fuse_wide_kernel<2>, grid, block, 0, st, x, low, out, relu, dec);
      case 4: return launch_pdl(pag_fuse_wide_kernel<4>, grid, block, 0, st, x, low, out, relu, dec);
      case 8: return launch_pdl(pag_fuse_wide_kernel<8>, grid, block, 0, st, x, low, out, relu, dec);
      case 16: return launch_pdl(pag_fuse_wide_kernel<16>, grid, block, 0, st, x, low, out, relu, dec);
      default: break;   // other widths: the flat kernel below
    }
  }
  const int LP = x.C / 8;
  const Dec dec = make_dec(LP, x.W, x.H, low.H, x.H, low.W, x.W);
  const dim3 grid(blocks_for(static_cast<long>(x.N) * x.H * x.W * LP, 256), 1, 1), block(256, 1, 1);
  switch (LP) {
    case 1: launch_pdl(pag_fuse_flat_kernel<1>, grid, block, 0, st, x, low, out, relu, dec); break;
    case 2: launch_pdl(pag_fuse_flat_kernel<2>, grid, block, 0, st, x, low, out, relu, dec); break;
    case 4: launch_pdl(pag_fuse_flat_kernel<4>, grid, block, 0, st, x, low, out, relu, dec); break;
    case 8: launch_pdl(pag_fuse_flat_kernel<8>, grid, block, 0, st, x, low, out, relu, dec); break;
    case 16: launch_pdl(pag_fuse_flat_kernel<16>, grid, block, 0, st, x, low, out, relu, dec); break;
    case 32: launch_pdl(pag_fuse_flat_kernel<32>, grid, block, 0, st, x, low, out, relu, dec); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t upadd_launch(View a, View b, View out, const float* s, const float* t, int relu, cudaStream_t st) {
  return upadd_res_launch(a, b, View{nullptr, 0, 0, 0, 0, 0}, out, s, t, relu, st);
}

cudaError_t upadd_res_launch(View a, View b, View r, View out, const float* s, const float* t, int relu, cudaStream_t st) {
  const long total = static_cast<long>(out.N) * out.H * out.W * (out.C / 8);
  if (total + 256 >= (1L << 32)) return cudaErrorInvalidValue;   // 32-bit index arithmetic in the kernel
  if (b.ptr) {
    const long strips = static_cast<long>(out.N) * ((out.H + kStrip - 1) / kStrip) * out.W * (out.C / 8);
    if (static_cast<long>(b.H) * b.W * b.ps >= (1L << 31)) return cudaErrorInvalidValue;   // 32-bit offsets inside an image
    launch_pdl(upadd_strip_kernel, dim3(blocks_for(strips, 256), 1, 1), dim3(256, 1, 1), 0, st, a, b, r, out, s, t, relu, make_dec(out.C / 8, out.W, (out.H + kStrip - 1) / kStrip, b.H, out.H, b.W, out.W));
  } else {
    launch_pdl(upadd_kernel, dim3(blocks_for(total, 256), 1, 1), dim3(256, 1, 1), 0, st, a, b, r, out, s, t, relu,
                                                          make_dec(out.C / 8, out.W, out.H, 0, 0, 0, 0));
  }
  return cudaGetLastError();
}

cudaError_t upadd_batch_launch(int njobs, const View* a, const View* b, const View* out, const float* const* s,
                               const float* const* t, int relu, cudaStream_t st) {
  if (njobs < 1 || njobs > kUpaddJobs) return cudaErrorInvalidValue;
  UpaddBatch p;
  std::memset(&p, 0, sizeof(p));
  p.relu = relu;
  for (int j = 0; j < njobs; ++j) {
    if (!b[j].ptr || out[j].N != out[0].N || out[j].H != out[0].H || out[j].W != out[0].W || out[j].C != out[0].C)
      return cudaErrorInvalidValue;
    if (static_cast<long>(b[j].H) * b[j].W * b[j].ps >= (1L << 31)) return cudaErrorInvalidValue;
    p.a[j] = a[j]; p.b[j] = b[j]; p.out[j] = out[j]; p.s[j] = s[j]; p.t[j] = t[j];
    p.dec[j] = make_dec(out[j].C / 8, out[j].W, (out[j].H + kStrip - 1) / kStrip, b[j].H, out[j].H, b[j].W, out[j].W);
  }
  const long strips = static_cast<long>(out[0].N) * ((out[0].H + kStrip - 1) / kStrip) * out[0].W * (out[0].C / 8);
  if (strips * kStrip + 256 >= (1L << 32)) return cudaErrorInvalidValue;
  launch_pdl(upadd_strip_batch_kernel, dim3(blocks_for(strips, 256), njobs, 1), dim3(256, 1, 1), 0, st, p);
  return cudaGetLastError();
}

cudaError_t pool_pyramid_launch(View x, const View out[4], const float* s, const float* t, const View ew[2], const float* se,
                                const float* te, cudaStream_t st) {
  if (x.C % 32 != 0) return cudaErrorNotSupported;
  const size_t smem = static_cast<size_t>(x.H + 1) * (x.W + 1) * 32 * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorNotSupported;
  static size_t opted = 0;
  if (smem > 48 * 1024 && smem > opted) {
    cudaError_t e = cudaFuncSetAttribute(pool_pyramid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    opted = 200 * 1024;
  }
  PyramidParams p;
  p.x = x; p.s = s; p.t = t; p.se = se; p.te = te;
  for (int i = 0; i < 4; ++i) p.out[i] = out[i];
  for (int i = 0; i < 2; ++i) p.ew[i] = ew ? ew[i] : View{nullptr, 0, 0, 0, 0, 0};
  launch_pdl(pool_pyramid_kernel, dim3(x.N * (x.C / 32), 1, 1), dim3(256, 1, 1), smem, st, p);
  return cudaGetLastError();
}

cudaError_t pool_affine_launch(View x, View out, int k, int stride, int pad, const float* s, const float* t, int relu,
                               cudaStream_t st) {
  const int groups = x.C / 8;
  if (groups > 256 || groups < 1) return cudaErrorInvalidValue;
  const int splits = 256 / groups;
  const int threads = splits * groups;
  const size_t smem = static_cast<size_t>(threads) * 8 * sizeof(float);
  launch_pdl(pool_affine_kernel, dim3(out.N * out.H * out.W, 1, 1), dim3(threads, 1, 1), smem, st, x, out, k, stride, pad, s, t, relu);
  return cudaGetLastError();
}

cudaError_t lightbag_uv_launch(View p, View i_low, View d, View out, cudaStream_t st) {
  const long total = static_cast<long>(p.N) * p.H * p.W * (p.C / 8);
  if (total + 256 >= (1L << 32)) return cudaErrorInvalidValue;   // 32-bit index arithmetic in the kernel
  launch_pdl(lightbag_uv_flat_kernel, dim3(blocks_for(total, 256), 1, 1), dim3(256, 1, 1), 0, st, p, i_low, d, out,
             make_dec(p.C / 8, p.W, p.H, i_low.H, p.H, i_low.W, p.W));
  return cudaGetLastError();
}

cudaError_t bag_blend_launch(View p, View i_low, View d, View out, const float* s, const float* t, cudaStream_t st) {
  const long total = static_cast<long>(p.N) * p.H * p.W * (p.C / 8);
  if (total + 256 >= (1L << 32)) return cudaErrorInvalidValue;   // 32-bit index arithmetic in the kernel
  launch_pdl(bag_blend_flat_kernel, dim3(blocks_for(total, 256), 1, 1), dim3(256, 1, 1), 0, st, p, i_low, d, out, s, t,
             make_dec(p.C / 8, p.W, p.H, i_low.H, p.H, i_low.W, p.W));
  return cudaGetLastError();
}

cudaError_t conv_ref_launch(const ConvRefParams& p, cudaStream_t st) {
  const long total = static_cast<long>(p.N) * p.Ho * p.Wo * p.Cout;
  conv_ref_kernel<<<blocks_for(total, 128), 128, 0, st>>>(p);
  return cudaGetLastError();
}

}  // namespace pidnet
