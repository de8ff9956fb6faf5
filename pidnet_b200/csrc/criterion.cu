// Fused training/validation criterion of the PIDNet path (reference: FullModel.forward utils/utils.py:37-57,
// OhemCrossEntropy utils/criterion.py:43-99, BondaryLoss :102-132; scalar restatement: SURVEY.md Appendix G).
//
// The reference upsamples the three low-res logit maps x8 (bilinear, align_corners=True) to label size,
// materialising ~2 GB per step at batch 12, then runs softmax / gather / a full sort twice.  Here one pass
// over the label pixels interpolates the 19+19+1 logits on the fly and emits, per pixel, only the target
// probability p, the weighted CE and two validity bits (9 B/pixel); the OHEM k-th order statistic is an
// exact 3-pass radix select on the bit pattern of p (skipped when #{p < thres} > k); a second pass forms
// the selected-pixel means; the backward pass recomputes the softmaxes and scatters the gradients through
// the transposed interpolation into shared-memory tiles before touching global memory.
// Everything stays on the device (no host sync); sums are accumulated in fp64.
#include <cuda_runtime.h>
#include <cstdint>
#include <cfloat>
#include <algorithm>
#include "criterion.cuh"

namespace pidnet {
namespace {

constexpr int kMaxC = 32;

// accumulator slots (double)
// [0, A_PIXEND) are filled by the per-pixel pass: A_NGE0 counts labels >= 0 (the denominator of pixel_acc, utils/utils.py:31-34),
// A_NBAD labels that are neither ignore_label nor a class index (the reference's gather would fault on them)
enum { A_CE_P = 0, A_ACC, A_BCE_POS, A_BCE_NEG, A_NPOS, A_NNEG, A_NV1, A_NV2, A_NLT1, A_NLT2, A_NGE0, A_NBAD, A_PIXEND,
       A_S1 = A_PIXEND, A_S2, A_K1, A_K2, A_COUNT };
// select state (per OHEM set s): u32 [need, prefix, rank, thr_bits]
struct SelState {
  unsigned need[2], prefix[2], rank[2];
  float thr[2];
};

struct Lerp {
  int i0, i1;
  float l;
};
// torch upsample_bilinear2d, align_corners=True: src = dst * (in-1)/(out-1)
__device__ __forceinline__ Lerp lerp_ac(int dst, int in, int out) {
  const float scale = out > 1 ? static_cast<float>(in - 1) / static_cast<float>(out - 1) : 0.f;
  const float src = scale * static_cast<float>(dst);
  Lerp r;
  r.i0 = static_cast<int>(src);
  if (r.i0 > in - 1) r.i0 = in - 1;
  r.i1 = r.i0 + (r.i0 < in - 1 ? 1 : 0);
  r.l = src - static_cast<float>(r.i0);
  return r;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct PixelCtx {
  int n, y, x;
  Lerp ly, lx;
  float w00, w01, w10, w11;
};
__device__ __forceinline__ PixelCtx make_ctx(long pix, int H, int W, int h, int w) {
  PixelCtx c;
  c.x = static_cast<int>(pix % W);
  const long t = pix / W;
  c.y = static_cast<int>(t % H);
  c.n = static_cast<int>(t / H);
  c.ly = lerp_ac(c.y, h, H);
  c.lx = lerp_ac(c.x, w, W);
  c.w00 = (1.f - c.ly.l) * (1.f - c.lx.l);
  c.w01 = (1.f - c.ly.l) * c.lx.l;
  c.w10 = c.ly.l * (1.f - c.lx.l);
  c.w11 = c.ly.l * c.lx.l;
  return c;
}
// interpolated value of plane `p` ([h][w]) -- torch's evaluation order: rows first, then columns
__device__ __forceinline__ float interp(const float* __restrict__ p, const PixelCtx& c, int w) {
  const float a = __ldg(p + c.ly.i0 * w + c.lx.i0), b = __ldg(p + c.ly.i0 * w + c.lx.i1);
  const float d = __ldg(p + c.ly.i1 * w + c.lx.i0), e = __ldg(p + c.ly.i1 * w + c.lx.i1);
  return (1.f - c.ly.l) * ((1.f - c.lx.l) * a + c.lx.l * b) + c.ly.l * ((1.f - c.lx.l) * d + c.lx.l * e);
}

// --------------------------------------------------------------------------------------- pass 1
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// CMAX: compile-time bound of the class loop (logits stay in registers; no dynamically indexed arrays)
template <int CMAX>
__global__ void __launch_bounds__(256) crit_pixel_kernel(CritParams p) {
  __shared__ double red[A_PIXEND][8];
  const long npix = static_cast<long>(p.N) * p.H * p.W;
  const long pix = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  // every slot receives at most one value per thread: the warp reduction runs in fp32 (exact for the
  // counters, one rounding for the losses), the cross-warp / cross-block accumulation in fp64
  float acc[A_PIXEND];
#pragma unroll
  for (int i = 0; i < A_PIXEND; ++i) acc[i] = 0.f;
  if (pix < npix) {
    const PixelCtx c = make_ctx(pix, p.H, p.W, p.h, p.w);
    const long t64 = p.labels[pix];
    // a label that is neither ignore_label nor a class index is never used as an index: the pixel is dropped and counted
    const bool inrange = t64 >= 0 && t64 < p.C;
    const bool valid1 = t64 != p.ignore_label && inrange;
    if (t64 != p.ignore_label && !inrange) acc[A_NBAD] = 1.f;
    if (t64 >= 0) acc[A_NGE0] = 1.f;
    const int t = valid1 ? static_cast<int>(t64) : 0;
    const size_t plane = static_cast<size_t>(p.h) * p.w;
    const float* xm = p.x_m + static_cast<size_t>(c.n) * p.C * plane;
    const float* xp = p.x_p + static_cast<size_t>(c.n) * p.C * plane;
    // main head: softmax prob of the target, weighted CE, argmax
    float vm[CMAX];
    float mx = -FLT_MAX, vt = 0.f;
    int amax = 0;
#pragma unroll
    for (int k = 0; k < CMAX; ++k) {
      if (k < p.C) {
        vm[k] = interp(xm + k * plane, c, p.w);
        if (vm[k] > mx) { mx = vm[k]; amax = k; }
        if (k == t) vt = vm[k];
      }
    }
    float se = 0.f;
#pragma unroll
    for (int k = 0; k < CMAX; ++k) if (k < p.C) se += expf(vm[k] - mx);
    const float lse = mx + logf(se);
    const float wt = (valid1 && p.class_w) ? __ldg(p.class_w + t) : 1.f;
    const float pm = expf(vt - lse);
    const float cem = wt * (lse - vt);
    acc[A_ACC] = (static_cast<long>(amax) == t64) ? 1.f : 0.f;   // pixel_acc counts every pixel (label >= 0), utils.py:31
    // aux head: plain weighted CE, reduction none (ignored pixels contribute 0)
    if (valid1) {
      float mp = -FLT_MAX, vtp = 0.f, sp = 0.f;
#pragma unroll
      for (int k = 0; k < CMAX; ++k) {
        if (k < p.C) {
          vm[k] = interp(xp + k * plane, c, p.w);
          mp = fmaxf(mp, vm[k]);
          if (k == t) vtp = vm[k];
        }
      }
#pragma unroll
      for (int k = 0; k < CMAX; ++k) if (k < p.C) sp += expf(vm[k] - mp);
      acc[A_CE_P] = wt * (mp + logf(sp) - vtp);
      if (p.aux_ce) p.aux_ce[pix] = acc[A_CE_P];
    }
    // boundary head
    const float xd = interp(p.x_d + static_cast<size_t>(c.n) * plane, c, p.w);
    const float z = p.bd_gt[pix];
    const float bce = fmaxf(xd, 0.f) - xd * z + log1pf(expf(-fabsf(xd)));
    if (z == 1.f) { acc[A_BCE_POS] = bce; acc[A_NPOS] = 1.f; }
    else if (z == 0.f) { acc[A_BCE_NEG] = bce; acc[A_NNEG] = 1.f; }
    const float sg = 1.f / (1.f + expf(-xd));
    const bool valid2 = valid1 && sg > p.bd_threshold;
    if (valid1) { acc[A_NV1] = 1.f; if (pm < p.ohem_thres) acc[A_NLT1] = 1.f; }
    if (valid2) { acc[A_NV2] = 1.f; if (pm < p.ohem_thres) acc[A_NLT2] = 1.f; }
    p.ws_p[pix] = pm;
    p.ws_ce[pix] = cem;
    p.ws_flags[pix] = static_cast<unsigned char>((valid1 ? 1 : 0) | (valid2 ? 2 : 0));
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < A_PIXEND; ++i) {
    const float v = warp_sum_f(acc[i]);
    if (lane == 0) red[i][warp] = static_cast<double>(v);
  }
  __syncthreads();
  if (threadIdx.x < A_PIXEND) {
    double v = 0.0;
    for (int q = 0; q < 8; ++q) v += red[threadIdx.x][q];
    if (v != 0.0) atomicAdd(p.accum + threadIdx.x, v);
  }
}

// --------------------------------------------------------------------------------------- OHEM threshold
__global__ void crit_select_init_kernel(CritParams p, SelState* st) {
  if (threadIdx.x < 2) {
    const int s = threadIdx.x;
    const double n = p.accum[A_NV1 + s], nlt = p.accum[A_NLT1 + s];
    const double k = fmin(static_cast<double>(p.min_kept), n - 1.0);   // index of min_value in the sorted list
    st->thr[s] = p.ohem_thres;
    st->prefix[s] = 0;
    st->rank[s] = k > 0 ? static_cast<unsigned>(k) : 0u;
    // p_sorted[k] < thres  <=>  at least k+1 valid pixels have p < thres  => threshold stays at thres
    st->need[s] = (n >= 1.0 && !(nlt >= k + 1.0)) ? 1u : 0u;
  }
}
// histogram of the next digit of p's bit pattern among candidates that match the prefix so far
__global__ void __launch_bounds__(256) crit_hist_kernel(CritParams p, const SelState* st, unsigned* hist, int shift,
                                                        int bits, int prefix_bits) {
  if (!st->need[0] && !st->need[1]) return;
  __shared__ unsigned sh[2][4096];
  const int nb = 1 << bits;
  for (int i = threadIdx.x; i < 2 * 4096; i += blockDim.x) (&sh[0][0])[i] = 0;
  __syncthreads();
  const long npix = static_cast<long>(p.N) * p.H * p.W;
  const unsigned need0 = st->need[0], need1 = st->need[1], pre0 = st->prefix[0], pre1 = st->prefix[1];
  for (long pix = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; pix < npix;
       pix += static_cast<long>(gridDim.x) * blockDim.x) {
    const unsigned f = p.ws_flags[pix];
    if (!f) continue;
    const unsigned u = __float_as_uint(p.ws_p[pix]);
    const unsigned hi = prefix_bits ? (u >> (32 - prefix_bits)) : 0u;
    const unsigned d = (u >> shift) & (nb - 1);
    if (need0 && (f & 1) && hi == pre0) atomicAdd(&sh[0][d], 1u);
    if (need1 && (f & 2) && hi == pre1) atomicAdd(&sh[1][d], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nb; i += blockDim.x) {
    if (sh[0][i]) atomicAdd(hist + i, sh[0][i]);
    if (sh[1][i]) atomicAdd(hist + 4096 + i, sh[1][i]);
  }
}
__global__ void crit_pick_kernel(SelState* st, unsigned* hist, int bits, int last, float thres) {
  const int s = threadIdx.x;
  if (s < 2 && st->need[s]) {
    const int nb = 1 << bits;
    unsigned r = st->rank[s], cum = 0;
    int d = 0;
    for (; d < nb; ++d) {
      const unsigned hcnt = hist[s * 4096 + d];
      if (cum + hcnt > r) break;
      cum += hcnt;
    }
    if (d >= nb) d = nb - 1;
    st->rank[s] = r - cum;
    st->prefix[s] = (st->prefix[s] << bits) | static_cast<unsigned>(d);
    if (last) st->thr[s] = fmaxf(__uint_as_float(st->prefix[s]), thres);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * 4096; i += blockDim.x) hist[i] = 0;
}

// --------------------------------------------------------------------------------------- pass 2
__global__ void __launch_bounds__(256) crit_sum_kernel(CritParams p, const SelState* st) {
  __shared__ double red[4][8];
  const long npix = static_cast<long>(p.N) * p.H * p.W;
  const float thr1 = st->thr[0], thr2 = st->thr[1];
  double a[4] = {0, 0, 0, 0};
  for (long pix = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; pix < npix;
       pix += static_cast<long>(gridDim.x) * blockDim.x) {
    const unsigned f = p.ws_flags[pix];
    if (!f) continue;
    const float pm = p.ws_p[pix], ce = p.ws_ce[pix];
    if ((f & 1) && pm < thr1) { a[0] += ce; a[2] += 1.0; }
    if ((f & 2) && pm < thr2) { a[1] += ce; a[3] += 1.0; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double v = warp_sum(a[i]);
    if (lane == 0) red[i][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double v = 0.0;
    for (int q = 0; q < 8; ++q) v += red[threadIdx.x][q];
    if (v != 0.0) atomicAdd(p.accum + A_S1 + threadIdx.x, v);
  }
}

__global__ void crit_final_kernel(CritParams p, const SelState* st) {
  if (threadIdx.x != 0) return;
  const double nhw = static_cast<double>(p.N) * p.H * p.W;
  const double* a = p.accum;
  const double aux = p.bw0 * a[A_CE_P] / nhw;
  const double ohem1 = a[A_S1] / a[A_K1];   // 0/0 -> NaN like torch's mean of an empty selection
  const double ohem2 = a[A_S2] / a[A_K2];
  const double nb = a[A_NPOS] + a[A_NNEG];
  const double bce = nb > 0 ? ((a[A_NNEG] / nb) * a[A_BCE_POS] + (a[A_NPOS] / nb) * a[A_BCE_NEG]) / nhw : 0.0;
  const double loss_b = p.coeff_bce * bce;
  const double loss_s = aux + p.bw1 * ohem1;
  float* o = p.out;
  // out-of-range labels: the reference's gather faults (device assert); here the loss is poisoned and o[12] reports the count
  const double poison = a[A_NBAD] > 0 ? __longlong_as_double(0x7ff8000000000000LL) : 0.0;
  o[0] = static_cast<float>(loss_s + loss_b + p.sb * ohem2 + poison);   // == losses.mean() of the reference
  o[1] = static_cast<float>(loss_s + poison);                           // == loss_list[0].mean()
  o[2] = static_cast<float>(loss_b);
  o[3] = static_cast<float>(a[A_ACC] / (a[A_NGE0] + 1e-10));            // pixel_acc: label >= 0 pixels, utils/utils.py:29-35
  o[12] = static_cast<float>(a[A_NBAD]);
  o[13] = static_cast<float>(a[A_NGE0]);
  o[14] = 0.f;
  o[15] = 0.f;
  o[4] = static_cast<float>(ohem1);
  o[5] = static_cast<float>(ohem2);
  o[6] = st->thr[0];
  o[7] = st->thr[1];
  o[8] = static_cast<float>(a[A_NV1]);
  o[9] = static_cast<float>(a[A_NV2]);
  o[10] = static_cast<float>(a[A_K1]);
  o[11] = static_cast<float>(a[A_K2]);
}

// --------------------------------------------------------------------------------------- backward
// Block = 32 x 8 label pixels; gradients are first accumulated into a shared tile of the low-res maps
// (<= 6 x 3 low-res pixels x (2C+1) channels for the x8 case), then flushed with one global atomic each.
constexpr int kTileW = 32, kTileH = 8, kLW = 8, kLH = 4;
__global__ void __launch_bounds__(256) crit_backward_kernel(CritParams p, const SelState* st) {
  extern __shared__ float tile[];   // [(2C+1)][kLH][kLW]
  const int CH = 2 * p.C + 1;
  for (int i = threadIdx.x; i < CH * kLH * kLW; i += blockDim.x) tile[i] = 0.f;
  __syncthreads();
  const int tiles_x = (p.W + kTileW - 1) / kTileW, tiles_y = (p.H + kTileH - 1) / kTileH;
  const int n = blockIdx.x / (tiles_x * tiles_y);
  const int rem = blockIdx.x - n * tiles_x * tiles_y;
  const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
  const int x = tx * kTileW + (threadIdx.x & 31), y = ty * kTileH + (threadIdx.x >> 5);
  // low-res origin of this tile
  const int ly0 = lerp_ac(ty * kTileH, p.h, p.H).i0, lx0 = lerp_ac(tx * kTileW, p.w, p.W).i0;
  const double nhw = static_cast<double>(p.N) * p.H * p.W;
  if (x < p.W && y < p.H) {
    const long pix = (static_cast<long>(n) * p.H + y) * p.W + x;
    const PixelCtx c = make_ctx(pix, p.H, p.W, p.h, p.w);
    const unsigned f = p.ws_flags[pix];
    const long t64 = p.labels[pix];
    const int t = (f & 1) ? static_cast<int>(t64) : 0;
    const size_t plane = static_cast<size_t>(p.h) * p.w;
    const int a00 = (c.ly.i0 - ly0) * kLW + (c.lx.i0 - lx0), a01 = (c.ly.i0 - ly0) * kLW + (c.lx.i1 - lx0);
    const int a10 = (c.ly.i1 - ly0) * kLW + (c.lx.i0 - lx0), a11 = (c.ly.i1 - ly0) * kLW + (c.lx.i1 - lx0);
    const size_t lplane = static_cast<size_t>(p.h) * p.w;
    auto scatter = [&](int ch, float g) {
      if (p.direct_scatter) {
        float* dst = ch < p.C ? p.g_p + (static_cast<size_t>(n) * p.C + ch) * lplane
                              : (ch < 2 * p.C ? p.g_m + (static_cast<size_t>(n) * p.C + (ch - p.C)) * lplane
                                              : p.g_d + static_cast<size_t>(n) * lplane);
        atomicAdd(dst + c.ly.i0 * p.w + c.lx.i0, g * c.w00);
        atomicAdd(dst + c.ly.i0 * p.w + c.lx.i1, g * c.w01);
        atomicAdd(dst + c.ly.i1 * p.w + c.lx.i0, g * c.w10);
        atomicAdd(dst + c.ly.i1 * p.w + c.lx.i1, g * c.w11);
        return;
      }
      float* tb = tile + ch * kLH * kLW;
      atomicAdd(tb + a00, g * c.w00);
      atomicAdd(tb + a01, g * c.w01);
      atomicAdd(tb + a10, g * c.w10);
      atomicAdd(tb + a11, g * c.w11);
    };
    const float wt = ((f & 1) && p.class_w) ? __ldg(p.class_w + t) : 1.f;
    if (f & 1) {
      // main head: coefficient from both OHEM selections
      const float pm = p.ws_p[pix];
      float coef = 0.f;
      if (pm < st->thr[0] && p.accum[A_K1] > 0) coef += static_cast<float>(p.bw1 / p.accum[A_K1]);
      if ((f & 2) && pm < st->thr[1] && p.accum[A_K2] > 0) coef += static_cast<float>(p.sb / p.accum[A_K2]);
      if (coef != 0.f) {
        const float* xm = p.x_m + static_cast<size_t>(n) * p.C * plane;
        float vm[kMaxC], mx = -FLT_MAX, se = 0.f;
        for (int k = 0; k < p.C; ++k) { vm[k] = interp(xm + k * plane, c, p.w); mx = fmaxf(mx, vm[k]); }
        for (int k = 0; k < p.C; ++k) { vm[k] = __expf(vm[k] - mx); se += vm[k]; }
        const float inv = 1.f / se;
        for (int k = 0; k < p.C; ++k) scatter(p.C + k, coef * wt * (vm[k] * inv - (k == t ? 1.f : 0.f)));
      }
      // aux head
      {
        const float* xp = p.x_p + static_cast<size_t>(n) * p.C * plane;
        float vp[kMaxC], mx = -FLT_MAX, se = 0.f;
        for (int k = 0; k < p.C; ++k) { vp[k] = interp(xp + k * plane, c, p.w); mx = fmaxf(mx, vp[k]); }
        for (int k = 0; k < p.C; ++k) { vp[k] = __expf(vp[k] - mx); se += vp[k]; }
        const float inv = 1.f / se, cf = static_cast<float>(p.bw0 / nhw) * wt;
        for (int k = 0; k < p.C; ++k) scatter(k, cf * (vp[k] * inv - (k == t ? 1.f : 0.f)));
      }
    }
    // boundary head
    {
      const float xd = interp(p.x_d + static_cast<size_t>(n) * plane, c, p.w);
      const float z = p.bd_gt[pix];
      const double nb = p.accum[A_NPOS] + p.accum[A_NNEG];
      float om = 0.f;
      if (z == 1.f) om = static_cast<float>(p.accum[A_NNEG] / nb);
      else if (z == 0.f) om = static_cast<float>(p.accum[A_NPOS] / nb);
      const float sg = 1.f / (1.f + __expf(-xd));
      scatter(2 * p.C, static_cast<float>(p.coeff_bce / nhw) * om * (sg - z));
    }
  }
  __syncthreads();
  if (p.direct_scatter) return;
  const size_t plane = static_cast<size_t>(p.h) * p.w;
  for (int i = threadIdx.x; i < CH * kLH * kLW; i += blockDim.x) {
    const float g = tile[i];
    if (g == 0.f) continue;
    const int ch = i / (kLH * kLW), r = i % (kLH * kLW);
    const int ly = ly0 + r / kLW, lx = lx0 + r % kLW;
    if (ly >= p.h || lx >= p.w) continue;
    float* dst = ch < p.C ? p.g_p + (static_cast<size_t>(n) * p.C + ch) * plane
                          : (ch < 2 * p.C ? p.g_m + (static_cast<size_t>(n) * p.C + (ch - p.C)) * plane
                                          : p.g_d + static_cast<size_t>(n) * plane);
    atomicAdd(dst + static_cast<size_t>(ly) * p.w + lx, g);
  }
}

// Tiled backward: same 32 x 8 label tile per block, but (a) the low-res logits of the tile's footprint are staged
// in shared memory once, (b) the class loops are register-resident (CMAX), and (c) the transposed interpolation
// along x is a segmented warp reduction -- lanes that share the left low-res column (contiguous runs of ~8 lanes)
// are summed with shuffles and only the run heads touch the shared gradient tile, which removes the ~8-way
// same-address shared-atomic serialisation of the plain scatter.
template <int CMAX>
__global__ void __launch_bounds__(256) crit_backward_tiled_kernel(CritParams p, const SelState* st) {
  extern __shared__ float sm[];
  const int CH = 2 * p.C + 1;
  constexpr int kLP = kLH * kLW;
  float* lo = sm;                 // [CH][kLH][kLW] staged logits: [0,C) aux head, [C,2C) main head, 2C boundary
  float* tile = sm + CH * kLP;    // [CH][kLH][kLW] gradient accumulators
  const int tiles_x = (p.W + kTileW - 1) / kTileW, tiles_y = (p.H + kTileH - 1) / kTileH;
  const int n = blockIdx.x / (tiles_x * tiles_y);
  const int rem = blockIdx.x - n * tiles_x * tiles_y;
  const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
  const int ly0 = lerp_ac(ty * kTileH, p.h, p.H).i0, lx0 = lerp_ac(tx * kTileW, p.w, p.W).i0;
  const size_t plane = static_cast<size_t>(p.h) * p.w;
  for (int i = threadIdx.x; i < CH * kLP; i += blockDim.x) {
    const int ch = i / kLP, r = i - ch * kLP;
    const int ly = min(ly0 + r / kLW, p.h - 1), lx = min(lx0 + r % kLW, p.w - 1);
    const float* src = ch < p.C ? p.x_p + (static_cast<size_t>(n) * p.C + ch) * plane
                                : (ch < 2 * p.C ? p.x_m + (static_cast<size_t>(n) * p.C + (ch - p.C)) * plane
                                                : p.x_d + static_cast<size_t>(n) * plane);
    lo[i] = __ldg(src + static_cast<size_t>(ly) * p.w + lx);
    tile[i] = 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int xr = tx * kTileW + lane, yr = ty * kTileH + (threadIdx.x >> 5);
  const bool inside = xr < p.W && yr < p.H;
  const int x = min(xr, p.W - 1), y = min(yr, p.H - 1);
  const Lerp ly = lerp_ac(y, p.h, p.H), lx = lerp_ac(x, p.w, p.W);
  const int a00 = (ly.i0 - ly0) * kLW + (lx.i0 - lx0), a01 = (ly.i0 - ly0) * kLW + (lx.i1 - lx0);
  const int a10 = (ly.i1 - ly0) * kLW + (lx.i0 - lx0), a11 = (ly.i1 - ly0) * kLW + (lx.i1 - lx0);
  // runs of lanes with the same left column: predicate bit b = "lane + 2^b belongs to my run"
  unsigned same = 0;
#pragma unroll
  for (int b = 0; b < 5; ++b) {
    const int other = __shfl_down_sync(0xffffffffu, lx.i0, 1 << b);
    if (lane + (1 << b) < 32 && other == lx.i0) same |= 1u << b;
  }
  const int prev = __shfl_up_sync(0xffffffffu, lx.i0, 1);
  const bool head = lane == 0 || prev != lx.i0;
  const float wy0 = 1.f - ly.l, wy1 = ly.l, wx0 = 1.f - lx.l, wx1 = lx.l;
  auto value = [&](int ch) {   // torch's evaluation order: rows first, then columns
    const float* q = lo + ch * kLP;
    return wy0 * (wx0 * q[a00] + wx1 * q[a01]) + wy1 * (wx0 * q[a10] + wx1 * q[a11]);
  };
  auto contribute = [&](int ch, float g) {   // executed by all 32 lanes
    float v0 = g * wx0, v1 = g * wx1;
#pragma unroll
    for (int b = 0; b < 5; ++b) {
      const float t0 = __shfl_down_sync(0xffffffffu, v0, 1 << b), t1 = __shfl_down_sync(0xffffffffu, v1, 1 << b);
      if (same & (1u << b)) { v0 += t0; v1 += t1; }
    }
    if (head) {
      float* tb = tile + ch * kLP;
      if (v0 != 0.f) { atomicAdd(tb + a00, v0 * wy0); atomicAdd(tb + a10, v0 * wy1); }
      if (v1 != 0.f) { atomicAdd(tb + a01, v1 * wy0); atomicAdd(tb + a11, v1 * wy1); }
    }
  };
  const double nhw = static_cast<double>(p.N) * p.H * p.W;
  const long pix = (static_cast<long>(n) * p.H + y) * p.W + x;
  const unsigned f = inside ? p.ws_flags[pix] : 0u;
  const int t = (f & 1) ? static_cast<int>(p.labels[pix]) : -1;
  const float wt = ((f & 1) && p.class_w) ? __ldg(p.class_w + t) : 1.f;
  float v[CMAX];
  // main head: coefficient from both OHEM selections
  float coef = 0.f;
  if (f & 1) {
    const float pm = p.ws_p[pix];
    if (pm < st->thr[0] && p.accum[A_K1] > 0) coef += static_cast<float>(p.bw1 / p.accum[A_K1]);
    if ((f & 2) && pm < st->thr[1] && p.accum[A_K2] > 0) coef += static_cast<float>(p.sb / p.accum[A_K2]);
  }
  if (__any_sync(0xffffffffu, coef != 0.f)) {
    float mx = -FLT_MAX, se = 0.f;
#pragma unroll
    for (int k = 0; k < CMAX; ++k) if (k < p.C) { v[k] = value(p.C + k); mx = fmaxf(mx, v[k]); }
#pragma unroll
    for (int k = 0; k < CMAX; ++k) if (k < p.C) { v[k] = __expf(v[k] - mx); se += v[k]; }
    const float cf = coef * wt, inv = 1.f / se;
#pragma unroll
    for (int k = 0; k < CMAX; ++k) if (k < p.C) contribute(p.C + k, cf * (v[k] * inv - (k == t ? 1.f : 0.f)));
  }
  // aux head
  if (__any_sync(0xffffffffu, (f & 1) != 0)) {
    float mx = -FLT_MAX, se = 0.f;
#pragma unroll
    for (int k = 0; k < CMAX; ++k) if (k < p.C) { v[k] = value(k); mx = fmaxf(mx, v[k]); }
#pragma unroll
    for (int k = 0; k < CMAX; ++k) if (k < p.C) { v[k] = __expf(v[k] - mx); se += v[k]; }
    const float cf = (f & 1) ? static_cast<float>(p.bw0 / nhw) * wt : 0.f, inv = 1.f / se;
#pragma unroll
    for (int k = 0; k < CMAX; ++k) if (k < p.C) contribute(k, cf * (v[k] * inv - (k == t ? 1.f : 0.f)));
  }
  // boundary head
  {
    float g = 0.f;
    if (inside) {
      const float xd = value(2 * p.C);
      const float z = p.bd_gt[pix];
      const double nb = p.accum[A_NPOS] + p.accum[A_NNEG];
      float om = 0.f;
      if (z == 1.f) om = static_cast<float>(p.accum[A_NNEG] / nb);
      else if (z == 0.f) om = static_cast<float>(p.accum[A_NPOS] / nb);
      const float sg = 1.f / (1.f + __expf(-xd));
      g = static_cast<float>(p.coeff_bce / nhw) * om * (sg - z);
    }
    contribute(2 * p.C, g);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < CH * kLP; i += blockDim.x) {
    const float g = tile[i];
    if (g == 0.f) continue;
    const int ch = i / kLP, r = i - ch * kLP;
    const int gy = ly0 + r / kLW, gx = lx0 + r % kLW;
    if (gy >= p.h || gx >= p.w) continue;
    float* dst = ch < p.C ? p.g_p + (static_cast<size_t>(n) * p.C + ch) * plane
                          : (ch < 2 * p.C ? p.g_m + (static_cast<size_t>(n) * p.C + (ch - p.C)) * plane
                                          : p.g_d + static_cast<size_t>(n) * plane);
    atomicAdd(dst + static_cast<size_t>(gy) * p.w + gx, g);
  }
}

// --------------------------------------------------------------------------------------- run kernels
// Fast path of both passes when the low-res footprint of a 256 x 8 label tile fits shared memory (always the case
// for the x8 heads): every thread owns a RUN of 8 consecutive label pixels of one row.
//  * the low-res logits of the tile footprint are staged in smem once (LDS instead of 4 global loads per class);
//  * the interpolation is one 4-term dot product with per-pixel weights (same value as torch's two-stage lerp up
//    to fp32 rounding), the softmax uses ex2 on pre-scaled arguments;
//  * backward: the 8 pixels of a run share their left low-res column (or switch once), so the transposed
//    interpolation is accumulated in REGISTERS and flushed to the smem gradient tile at most twice per run --
//    lanes of a warp are 8 pixels (~1 low-res column) apart and therefore hit distinct addresses.
constexpr int kPX = 8, kRW = 32 * kPX, kRH = 8, kRLW = 36, kRLH = 4, kRLP = kRLW * kRLH;
constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// ---- run kernels, column-owned form.  ncu on the first form (12 x 1024 x 1024, C = 19): 1106 / 2213 warp instructions per
// label pixel (forward / backward), 66 % / 45 % issue-active, i.e. INSTRUCTION bound: four shared loads + four multiply-adds
// per class and pixel for the interpolation, one runtime `k < C` test per class in every loop, and (backward) a gradient
// flush that diverges inside the warp.  Here
//  * the class count is a template constant for the dataset values (C = 19 Cityscapes, 11 CamVid): no per-class predicates;
//  * a thread owns ONE LOW-RES COLUMN INTERVAL of one label row: all label pixels x with floor(x * (w-1)/(W-1)) == c (8 or 9
//    consecutive pixels for the x8 heads).  They share their two low-res columns, so the vertically interpolated logits
//    c0[k] = lerp_y(col c), d[k] = lerp_y(col c+1) - c0[k] are loaded ONCE per thread and a pixel costs one fma per class;
//    in the backward pass the transposed interpolation accumulates in registers for the whole run and is flushed once --
//    no reload / flush inside the pixel loop, hence no divergence (a first run-of-8-pixels form reloaded on a column change,
//    which happens at a different pixel in every lane: the warp executed the reload block at almost every pixel);
//  * the target logit is re-interpolated on its own (7 instructions) instead of a select per class; ex2 / lg2 everywhere;
//  * labels / boundary targets come in and p / CE / flags go out through shared memory, so global IO stays coalesced.
constexpr int kCT = 32;              // low-res columns per tile (one per lane)
constexpr int kMaxRun = 10;          // label pixels per low-res column: <= ceil((W-1)/(w-1)) + 1
constexpr int kCS = kCT * (kMaxRun + 1);   // padded smem row: pixel j of column lane c at c * 11 + j (odd stride: conflict-free)
__device__ __forceinline__ float lg2f(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// per-pixel code byte: bits 0-4 class, 5 valid (label is a class index and not ignore_label), 6 label out of range, 7 label >= 0
__device__ __forceinline__ unsigned label_code(long t64, long ignore, int C) {
  const bool inrange = t64 >= 0 && t64 < C;
  const bool valid = t64 != ignore && inrange;
  return (valid ? (static_cast<unsigned>(t64) | 0x20u) : 0u) | ((t64 != ignore && !inrange) ? 0x40u : 0u) | (t64 >= 0 ? 0x80u : 0u);
}
// the low-res column lerp_ac assigns to label pixel x (its i0), and the first label pixel of column c (W for c >= w) --
// both with lerp_ac's own float arithmetic, so that "the pixels of column c" is exactly {x : lerp_ac(x).i0 == c}
__device__ __forceinline__ int col_of(int x, float scale, int w) { return min(static_cast<int>(scale * static_cast<float>(x)), w - 1); }
__device__ __forceinline__ int first_x_of_col(int c, float scale, int w, int W) {
  if (c <= 0) return 0;
  if (c >= w) return W;
  int x = scale > 0.f ? static_cast<int>(static_cast<float>(c) / scale) : W;
  x = max(0, min(x, W));
  while (x > 0 && col_of(x - 1, scale, w) >= c) --x;
  while (x < W && col_of(x, scale, w) < c) ++x;
  return x;
}
struct ColTile {
  int n, ty, tx, ly0, lx0, xb, xe;
  float scale;
};
__device__ __forceinline__ ColTile col_tile(const CritParams& p) {
  ColTile t;
  const int tiles_x = (p.w + kCT - 1) / kCT, tiles_y = (p.H + kRH - 1) / kRH;
  t.n = blockIdx.x / (tiles_x * tiles_y);
  const int rem = blockIdx.x - t.n * tiles_x * tiles_y;
  t.ty = rem / tiles_x;
  t.tx = rem - t.ty * tiles_x;
  t.scale = p.W > 1 ? static_cast<float>(p.w - 1) / static_cast<float>(p.W - 1) : 0.f;
  t.ly0 = lerp_ac(t.ty * kRH, p.h, p.H).i0;
  t.lx0 = t.tx * kCT;
  t.xb = first_x_of_col(t.lx0, t.scale, p.w, p.W);
  t.xe = first_x_of_col(t.lx0 + kCT, t.scale, p.w, p.W);
  return t;
}
// lo[ch][kRLH][kRLW]: [0,C) aux head, [C,2C) main head, 2C boundary.  The per-element index arithmetic (two divisions, clamps,
// a three-way pointer select) cost 122 instructions per staged value in the first version -- 335 per label pixel, more than
// the softmax: the clamped source offsets are tabulated once per block, the heads are separate loops.
__device__ __forceinline__ void stage_lowres_col(const CritParams& p, const ColTile& t, int C, float* lo, float* zero) {
  __shared__ int s_off[kRLP];
  for (int r = threadIdx.x; r < kRLP; r += blockDim.x) {
    const int ly = min(t.ly0 + r / kRLW, p.h - 1), lx = min(t.lx0 + r % kRLW, p.w - 1);
    s_off[r] = ly * p.w + lx;
  }
  __syncthreads();
  const size_t plane = static_cast<size_t>(p.h) * p.w;
  const float* base[3] = {p.x_p + static_cast<size_t>(t.n) * C * plane, p.x_m + static_cast<size_t>(t.n) * C * plane,
                          p.x_d + static_cast<size_t>(t.n) * plane};
#pragma unroll
  for (int hd = 0; hd < 3; ++hd) {
    const int nval = (hd == 2 ? 1 : C) * kRLP;
    float* dst = lo + hd * C * kRLP;
    float* zdst = zero ? zero + hd * C * kRLP : nullptr;
    const float* src = base[hd];
    for (int i = threadIdx.x; i < nval; i += blockDim.x) {
      const int ch = i / kRLP, r = i - ch * kRLP;
      // asynchronous 4-byte copies: all of a thread's ~22 loads are in flight at once (with plain loads the loop waited for
      // each value before storing it -- ncu: a quarter of the kernel's samples sat on that store)
      const unsigned sdst = static_cast<unsigned>(__cvta_generic_to_shared(dst + i));
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sdst), "l"(src + static_cast<size_t>(ch) * plane + s_off[r]) : "memory");
      if (zdst) zdst[i] = 0.f;
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void stage_lowres_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// smem slot (within a row) of label pixel x of this tile
__device__ __forceinline__ int col_slot(int x, const ColTile& t, int w, int W) {
  const int c = col_of(x, t.scale, w);
  return (c - t.lx0) * (kMaxRun + 1) + (x - first_x_of_col(c, t.scale, w, W));
}

template <int CMAX, bool EXACT>
__global__ void __launch_bounds__(256, 3) crit_pixel_run_kernel(CritParams p) {
  extern __shared__ float lo[];                       // [CH][kRLH][kRLW]
  __shared__ double red[A_PIXEND][8];
  __shared__ unsigned char s_lab[kRH * kCS], s_zc[kRH * kCS], s_fl[kRH * kCS];   // staged label / boundary-target codes, flags
  __shared__ float s_pm[kRH * kCS], s_ce[kRH * kCS];
  const int C = EXACT ? CMAX : p.C;
  const ColTile t = col_tile(p);
  stage_lowres_col(p, t, C, lo, nullptr);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // coalesced staging of the tile's labels and boundary targets: a thread takes the same pixel column(s) of all 8 rows
  for (int x = t.xb + static_cast<int>(threadIdx.x); x < t.xe; x += 256) {
    const int slot = col_slot(x, t, p.w, p.W);
    const int nrow = min(kRH, p.H - t.ty * kRH);
    const long pix0 = (static_cast<long>(t.n) * p.H + t.ty * kRH) * p.W + x;
    long lab[kRH];
    float zz[kRH];
#pragma unroll
    for (int i = 0; i < kRH; ++i) {      // all 16 loads of the column in flight before the first use
      const long pix = pix0 + static_cast<long>(i < nrow ? i : 0) * p.W;
      lab[i] = __ldg(p.labels + pix);
      zz[i] = __ldg(p.bd_gt + pix);
    }
#pragma unroll
    for (int i = 0; i < kRH; ++i) {
      if (i < nrow) {
        s_lab[i * kCS + slot] = static_cast<unsigned char>(label_code(lab[i], p.ignore_label, C));
        s_zc[i * kCS + slot] = static_cast<unsigned char>(zz[i] == 1.f ? 1u : (zz[i] == 0.f ? 0u : 2u));
      }
    }
  }
  stage_lowres_wait();
  __syncthreads();
  const int y = t.ty * kRH + warp;
  const int c = t.lx0 + lane;
  float acc[A_PIXEND];
#pragma unroll
  for (int i = 0; i < A_PIXEND; ++i) acc[i] = 0.f;
  const int my = warp * kCS + lane * (kMaxRun + 1);
  if (y < p.H && c < p.w) {
    const int xs = first_x_of_col(c, t.scale, p.w, p.W);
    const int npx = first_x_of_col(c + 1, t.scale, p.w, p.W) - xs;
    const Lerp ly = lerp_ac(y, p.h, p.H);
    const int i1 = lane + (c < p.w - 1 ? 1 : 0);                 // staged column of lerp_ac(x).i1
    const int r0 = (ly.i0 - t.ly0) * kRLW, r1 = (ly.i1 - t.ly0) * kRLW;
    const float wy0 = 1.f - ly.l, wy1 = ly.l;
    const float logit_thr = __logf(p.bd_threshold / (1.f - p.bd_threshold));   // sigmoid(x) > thr  <=>  x > logit(thr)
    const float fc = static_cast<float>(c);
    float c0[CMAX], d[CMAX];
    // ---------------- main head: p(target), weighted CE, argmax hit; OHEM counters
#pragma unroll
    for (int k = 0; k < CMAX; ++k) {
      if (EXACT || k < C) {
        const float* q = lo + (C + k) * kRLP;
        const float a = fmaf(wy1, q[r1 + lane], wy0 * q[r0 + lane]), b = fmaf(wy1, q[r1 + i1], wy0 * q[r0 + i1]);
        c0[k] = a; d[k] = b - a;
      }
    }
#pragma unroll 1
    for (int j = 0; j < npx; ++j) {
      const float l = t.scale * static_cast<float>(xs + j) - fc;   // == lerp_ac(x).l (same product, i0 == c)
      const unsigned code = s_lab[my + j];
      const bool valid1 = (code & 0x20u) != 0;
      const int tg = code & 0x1Fu;
      float mx = -FLT_MAX;
      int amax = 0;
#pragma unroll
      for (int k = 0; k < CMAX; ++k) {
        if (EXACT || k < C) {
          const float v = fmaf(l, d[k], c0[k]);
          if (v > mx) { mx = v; amax = k; }
        }
      }
      float se = 0.f;
      const float mxs = -mx * kLog2e;
#pragma unroll
      for (int k = 0; k < CMAX; ++k) if (EXACT || k < C) se += ex2f(fmaf(fmaf(l, d[k], c0[k]), kLog2e, mxs));
      // the target's logit, interpolated on its own with the same operations as c0 / d above (bit-identical to v[tg])
      const float* qt = lo + (C + tg) * kRLP;
      const float ta = fmaf(wy1, qt[r1 + lane], wy0 * qt[r0 + lane]), tb = fmaf(wy1, qt[r1 + i1], wy0 * qt[r0 + i1]);
      const float vt = fmaf(l, tb - ta, ta);
      // (full-precision log / exp here: p feeds the OHEM order statistic, where a pixel within rounding of the threshold flips
      //  in or out of the selection)
      const float lse = mx + logf(se);
      const float pm = expf(vt - lse);
      const float wt = (valid1 && p.class_w) ? __ldg(p.class_w + tg) : 1.f;
      if (valid1 && amax == tg) acc[A_ACC] += 1.f;             // pixel_acc numerator: prediction == label (label >= 0), utils.py:31-33
      if (code & 0x80u) acc[A_NGE0] += 1.f;
      if (code & 0x40u) acc[A_NBAD] += 1.f;
      s_pm[my + j] = pm;
      s_ce[my + j] = wt * (lse - vt);
      if (valid1) { acc[A_NV1] += 1.f; if (pm < p.ohem_thres) acc[A_NLT1] += 1.f; }
    }
    // ---------------- boundary head: weighted BCE sums, bd_label validity
    {
      const float* q = lo + 2 * C * kRLP;
      const float b0 = fmaf(wy1, q[r1 + lane], wy0 * q[r0 + lane]);
      const float bd = fmaf(wy1, q[r1 + i1], wy0 * q[r0 + i1]) - b0;
#pragma unroll 1
      for (int j = 0; j < npx; ++j) {
        const float l = t.scale * static_cast<float>(xs + j) - fc;
        const float xd = fmaf(l, bd, b0);
        const unsigned zc = s_zc[my + j];
        // BCE with logits: max(x, 0) - x z + log(1 + exp(-|x|))
        const float sp = __logf(1.f + __expf(-fabsf(xd)));
        if (zc == 1u) { acc[A_BCE_POS] += fmaxf(xd, 0.f) - xd + sp; acc[A_NPOS] += 1.f; }
        else if (zc == 0u) { acc[A_BCE_NEG] += fmaxf(xd, 0.f) + sp; acc[A_NNEG] += 1.f; }
        const bool valid1 = (s_lab[my + j] & 0x20u) != 0;
        const bool valid2 = valid1 && xd > logit_thr;
        if (valid2) { acc[A_NV2] += 1.f; if (s_pm[my + j] < p.ohem_thres) acc[A_NLT2] += 1.f; }
        s_fl[my + j] = static_cast<unsigned char>((valid1 ? 1 : 0) | (valid2 ? 2 : 0));
      }
    }
    // ---------------- aux head: plain weighted CE over valid pixels (reduction none; ignored pixels contribute 0)
#pragma unroll
    for (int k = 0; k < CMAX; ++k) {
      if (EXACT || k < C) {
        const float* q = lo + k * kRLP;
        const float a = fmaf(wy1, q[r1 + lane], wy0 * q[r0 + lane]), b = fmaf(wy1, q[r1 + i1], wy0 * q[r0 + i1]);
        c0[k] = a; d[k] = b - a;
      }
    }
#pragma unroll 1
    for (int j = 0; j < npx; ++j) {
      const unsigned code = s_lab[my + j];
      if (!(code & 0x20u)) continue;
      const int tg = code & 0x1Fu;
      const float l = t.scale * static_cast<float>(xs + j) - fc;
      float mp = -FLT_MAX;
#pragma unroll
      for (int k = 0; k < CMAX; ++k) if (EXACT || k < C) mp = fmaxf(mp, fmaf(l, d[k], c0[k]));
      float sp = 0.f;
      const float mps = -mp * kLog2e;
#pragma unroll
      for (int k = 0; k < CMAX; ++k) if (EXACT || k < C) sp += ex2f(fmaf(fmaf(l, d[k], c0[k]), kLog2e, mps));
      const float* qt = lo + tg * kRLP;
      const float ta = fmaf(wy1, qt[r1 + lane], wy0 * qt[r0 + lane]), tb = fmaf(wy1, qt[r1 + i1], wy0 * qt[r0 + i1]);
      const float vtp = fmaf(l, tb - ta, ta);
      const float wt = p.class_w ? __ldg(p.class_w + tg) : 1.f;
      const float cep = wt * (fmaf(mp, kLog2e, lg2f(sp)) * kLn2 - vtp);
      acc[A_CE_P] += cep;
      if (p.aux_ce) p.aux_ce[(static_cast<long>(t.n) * p.H + y) * p.W + xs + j] = cep;   // (rare path: the map is pre-zeroed)
    }
  }
  __syncthreads();
  // coalesced write-out of p / CE / flags
  for (int x = t.xb + static_cast<int>(threadIdx.x); x < t.xe; x += 256) {
    const int slot = col_slot(x, t, p.w, p.W);
#pragma unroll
    for (int i = 0; i < kRH; ++i) {
      const int yy = t.ty * kRH + i;
      if (yy >= p.H) break;
      const long pix = (static_cast<long>(t.n) * p.H + yy) * p.W + x;
      p.ws_p[pix] = s_pm[i * kCS + slot];
      p.ws_ce[pix] = s_ce[i * kCS + slot];
      p.ws_flags[pix] = s_fl[i * kCS + slot];
    }
  }
#pragma unroll
  for (int i = 0; i < A_PIXEND; ++i) {
    const float v = warp_sum_f(acc[i]);
    if (lane == 0) red[i][warp] = static_cast<double>(v);
  }
  __syncthreads();
  if (threadIdx.x < A_PIXEND) {
    double v = 0.0;
    for (int q = 0; q < 8; ++q) v += red[threadIdx.x][q];
    if (v != 0.0) atomicAdd(p.accum + threadIdx.x, v);
  }
}

// backward.  Phase 1 (coalesced): per pixel a code word {target | kept-by-OHEM-1 | kept-by-OHEM-2 | valid} and the
// boundary target go to smem.  Phase 2: thread = (label row, low-res column); the transposed interpolation of its run
// accumulates in registers (a0: column c, a1: column c + 1), the a1 of the left neighbour lane is folded in with one shuffle
// and every lane adds its column to the two low-res rows of the shared gradient tile; one global atomic per tile element.
template <int CMAX, bool EXACT>
__global__ void __launch_bounds__(256, 2) crit_backward_run_kernel(CritParams p, const SelState* st) {
  extern __shared__ float sm[];
  const int C = EXACT ? CMAX : p.C;
  const int CH = 2 * C + 1;
  float* lo = sm;
  float* tile = sm + CH * kRLP;
  int* s_code = reinterpret_cast<int*>(tile + CH * kRLP);   // [kRH][kCS]
  float* s_z = reinterpret_cast<float*>(s_code + kRH * kCS);
  const ColTile t = col_tile(p);
  stage_lowres_col(p, t, C, lo, tile);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float thr0 = st->thr[0], thr1 = st->thr[1];
  for (int x = t.xb + static_cast<int>(threadIdx.x); x < t.xe; x += 256) {
    const int slot = col_slot(x, t, p.w, p.W);
    const int nrow = min(kRH, p.H - t.ty * kRH);
    const long pix0 = (static_cast<long>(t.n) * p.H + t.ty * kRH) * p.W + x;
    unsigned fl[kRH];
    float pmv[kRH], zz[kRH];
    int lab[kRH];
#pragma unroll
    for (int i = 0; i < kRH; ++i) {      // all loads of the column in flight before the first use
      const long pix = pix0 + static_cast<long>(i < nrow ? i : 0) * p.W;
      fl[i] = p.ws_flags[pix];
      pmv[i] = p.ws_p[pix];
      lab[i] = static_cast<int>(__ldg(p.labels + pix));
      zz[i] = __ldg(p.bd_gt + pix);
    }
#pragma unroll
    for (int i = 0; i < kRH; ++i) {
      if (i < nrow) {
        int code = 0;
        if (fl[i] & 1) {
          code = (lab[i] & 0xFF) | 0x400;
          if (pmv[i] < thr0) code |= 0x100;
          if ((fl[i] & 2) && pmv[i] < thr1) code |= 0x200;
        }
        s_code[i * kCS + slot] = code;
        s_z[i * kCS + slot] = zz[i];
      }
    }
  }
  stage_lowres_wait();
  __syncthreads();
  const int y = t.ty * kRH + warp;
  const int c = t.lx0 + lane;
  if (y < p.H) {                                               // (whole warps: the shuffles below stay convergent)
    const bool live = c < p.w;
    const int xs = live ? first_x_of_col(c, t.scale, p.w, p.W) : 0;
    const int npx = live ? first_x_of_col(c + 1, t.scale, p.w, p.W) - xs : 0;
    const Lerp ly = lerp_ac(y, p.h, p.H);
    const int i1 = lane + (c < p.w - 1 ? 1 : 0);
    const int r0 = (ly.i0 - t.ly0) * kRLW, r1 = (ly.i1 - t.ly0) * kRLW;
    const float wy0 = 1.f - ly.l, wy1 = ly.l;
    const double nhw = static_cast<double>(p.N) * p.H * p.W;
    // an empty OHEM selection (the reference raises IndexError, criterion.py:73) contributes no gradient; the loss is NaN
    const float c_k1 = p.accum[A_K1] > 0 ? static_cast<float>(p.bw1 / p.accum[A_K1]) : 0.f;
    const float c_k2 = p.accum[A_K2] > 0 ? static_cast<float>(p.sb / p.accum[A_K2]) : 0.f;
    const float c_aux = static_cast<float>(p.bw0 / nhw);
    const int* my_code = s_code + warp * kCS + lane * (kMaxRun + 1);
    const float* my_z = s_z + warp * kCS + lane * (kMaxRun + 1);
    const float fc = static_cast<float>(c);
    // column of the tile that receives a1 of lane 31 (or of the image's last column, where i1 == i0)
    float a0[CMAX], a1[CMAX], c0[CMAX], d[CMAX];
#pragma unroll 1
    for (int hd = 0; hd < 2; ++hd) {   // 0: main head (OHEM coefficients), 1: aux head
      const int ch0 = hd == 0 ? C : 0;
#pragma unroll
      for (int k = 0; k < CMAX; ++k) {
        if (EXACT || k < C) {
          const float* q = lo + (ch0 + k) * kRLP;
          const float a = fmaf(wy1, q[r1 + lane], wy0 * q[r0 + lane]), b = fmaf(wy1, q[r1 + i1], wy0 * q[r0 + i1]);
          c0[k] = a; d[k] = b - a;
          a0[k] = 0.f; a1[k] = 0.f;
        }
      }
#pragma unroll 1
      for (int j = 0; j < npx; ++j) {
        const int code = my_code[j];
        if (!(code & 0x400)) continue;
        const int tg = code & 0xFF;
        const float wt = p.class_w ? __ldg(p.class_w + tg) : 1.f;
        const float cf = wt * (hd == 0 ? ((code & 0x100) ? c_k1 : 0.f) + ((code & 0x200) ? c_k2 : 0.f) : c_aux);
        if (cf == 0.f) continue;
        const float l = t.scale * static_cast<float>(xs + j) - fc;
        float v[CMAX];
        float mx = -FLT_MAX, se = 0.f;
#pragma unroll
        for (int k = 0; k < CMAX; ++k) if (EXACT || k < C) { v[k] = fmaf(l, d[k], c0[k]); mx = fmaxf(mx, v[k]); }
        const float mxs = -mx * kLog2e;
#pragma unroll
        for (int k = 0; k < CMAX; ++k) if (EXACT || k < C) { v[k] = ex2f(fmaf(v[k], kLog2e, mxs)); se += v[k]; }
        const float sc = __fdividef(cf, se), g0 = 1.f - l, g1 = l;
#pragma unroll
        for (int k = 0; k < CMAX; ++k) {
          if (EXACT || k < C) {
            const float g = fmaf(v[k], sc, k == tg ? -cf : 0.f);
            a0[k] = fmaf(g, g0, a0[k]);
            a1[k] = fmaf(g, g1, a1[k]);
          }
        }
      }
      // flush: column (lane) gets a0 + the left neighbour's a1; lane 31 / the image's last column keep their a1 for column i1
      const bool own_a1 = lane == 31 || i1 == lane;
#pragma unroll
      for (int k = 0; k < CMAX; ++k) {
        if (EXACT || k < C) {
          float* tb = tile + (ch0 + k) * kRLP;
          const float left = __shfl_up_sync(0xffffffffu, a1[k], 1);
          float mine = a0[k] + (lane > 0 ? left : 0.f);
          if (i1 == lane) { mine += a1[k]; }                   // last image column: both taps are column c
          atomicAdd(tb + r0 + lane, mine * wy0);
          atomicAdd(tb + r1 + lane, mine * wy1);
          if (own_a1 && i1 != lane) { atomicAdd(tb + r0 + i1, a1[k] * wy0); atomicAdd(tb + r1 + i1, a1[k] * wy1); }
        }
      }
    }
    // boundary head (one channel): same run accumulation
    {
      const double nb = p.accum[A_NPOS] + p.accum[A_NNEG];
      const float om_pos = static_cast<float>(p.accum[A_NNEG] / nb), om_neg = static_cast<float>(p.accum[A_NPOS] / nb);
      const float c_bce = static_cast<float>(p.coeff_bce / nhw);
      float b0 = 0.f, b1 = 0.f;
      float* tb = tile + 2 * C * kRLP;
      const float* q = lo + 2 * C * kRLP;
      const float q0 = fmaf(wy1, q[r1 + lane], wy0 * q[r0 + lane]);
      const float qd = fmaf(wy1, q[r1 + i1], wy0 * q[r0 + i1]) - q0;
#pragma unroll 1
      for (int j = 0; j < npx; ++j) {
        const float z = my_z[j];
        const float om = z == 1.f ? om_pos : (z == 0.f ? om_neg : 0.f);
        if (om == 0.f) continue;
        const float l = t.scale * static_cast<float>(xs + j) - fc;
        const float xd = fmaf(l, qd, q0);
        const float sg = 1.f / (1.f + __expf(-xd));
        const float g = c_bce * om * (sg - z);
        b0 = fmaf(g, 1.f - l, b0);
        b1 = fmaf(g, l, b1);
      }
      const float left = __shfl_up_sync(0xffffffffu, b1, 1);
      float mine = b0 + (lane > 0 ? left : 0.f);
      if (i1 == lane) mine += b1;
      atomicAdd(tb + r0 + lane, mine * wy0);
      atomicAdd(tb + r1 + lane, mine * wy1);
      if (lane == 31 && i1 != lane) { atomicAdd(tb + r0 + i1, b1 * wy0); atomicAdd(tb + r1 + i1, b1 * wy1); }
    }
  }
  __syncthreads();
  const size_t plane = static_cast<size_t>(p.h) * p.w;
  for (int i = threadIdx.x; i < CH * kRLP; i += blockDim.x) {
    const float g = tile[i];
    if (g == 0.f) continue;
    const int ch = i / kRLP, r = i - ch * kRLP;
    const int gy = t.ly0 + r / kRLW, gx = t.lx0 + r % kRLW;
    if (gy >= p.h || gx >= p.w) continue;
    float* dst = ch < C ? p.g_p + (static_cast<size_t>(t.n) * C + ch) * plane
                        : (ch < 2 * C ? p.g_m + (static_cast<size_t>(t.n) * C + (ch - C)) * plane
                                      : p.g_d + static_cast<size_t>(t.n) * plane);
    atomicAdd(dst + static_cast<size_t>(gy) * p.w + gx, g);
  }
}

// does the low-res footprint of a (tw x th) label tile fit a (lw x lh) staging tile?
inline bool footprint_fits(const CritParams& p, int tw, int th, int lw, int lh) {
  auto span = [](int t, int in, int out) { return out > 1 ? (static_cast<long>(t - 1) * (in - 1)) / (out - 1) + 3 : 1; };
  const long sw = std::min<long>(span(std::min(tw, p.W), p.w, p.W), p.w), sh = std::min<long>(span(std::min(th, p.H), p.h, p.H), p.h);
  return sw <= lw && sh <= lh;
}
// column-owned run kernels: 8 label rows must span <= kRLH low-res rows and a low-res column <= kMaxRun label pixels
inline bool column_runs_fit(const CritParams& p) {
  auto span = [](int t, int in, int out) { return out > 1 ? (static_cast<long>(t - 1) * (in - 1)) / (out - 1) + 3 : 1; };
  const long sh = std::min<long>(span(std::min(kRH, p.H), p.h, p.H), p.h);
  if (p.w < 2 || p.W < 2) return false;
  const long per_col = (static_cast<long>(p.W) - 1 + (p.w - 2)) / (p.w - 1) + 1;    // ceil((W-1)/(w-1)) + 1
  return sh <= kRLH && per_col <= kMaxRun;
}
template <int CMAX, bool EXACT>
cudaError_t launch_run_kernels(const CritParams& p, const SelState* sel, bool backward, cudaStream_t st) {
  const unsigned blocks = static_cast<unsigned>(p.N) * ((p.w + kCT - 1) / kCT) * ((p.H + kRH - 1) / kRH);
  const size_t lo_bytes = static_cast<size_t>(2 * p.C + 1) * kRLP * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    const int max_bytes = (2 * CMAX + 1) * kRLP * static_cast<int>(sizeof(float));
    cudaError_t e = cudaFuncSetAttribute(crit_pixel_run_kernel<CMAX, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(crit_backward_run_kernel<CMAX, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             2 * max_bytes + 2 * kRH * kCS * 4);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  if (!backward) crit_pixel_run_kernel<CMAX, EXACT><<<blocks, 256, lo_bytes, st>>>(p);
  else crit_backward_run_kernel<CMAX, EXACT><<<blocks, 256, 2 * lo_bytes + 2 * kRH * kCS * 4, st>>>(p, sel);
  return cudaGetLastError();
}
cudaError_t launch_run(const CritParams& p, const SelState* sel, bool backward, cudaStream_t st) {
  if (p.C == 19) return launch_run_kernels<19, true>(p, sel, backward, st);    // Cityscapes
  if (p.C == 11) return launch_run_kernels<11, true>(p, sel, backward, st);    // CamVid
  if (p.C <= 12) return launch_run_kernels<12, false>(p, sel, backward, st);
  if (p.C <= 20) return launch_run_kernels<20, false>(p, sel, backward, st);
  return launch_run_kernels<kMaxC, false>(p, sel, backward, st);
}

// --------------------------------------------------------------------------------------- post-processing
// Fused x8 bilinear upsample (align_corners=True) + argmax (+ confusion-matrix histogram): replaces
// F.interpolate -> [exp] -> argmax -> .cpu() of datasets/base_dataset.py:136-150 / tools/custom.py:90-92 and
// get_confusion_matrix of utils/utils.py:129-152 without materialising the [N,C,H,W] fp32 tensor (159 MB per
// 1024x2048 image).  The interpolation is evaluated in torch's operation order WITHOUT fma contraction so the class
// index is bit-identical to a plain fp32 restatement of that formula (what the parity tests compare against).
struct PostParams {
  const float* x;            // [N,C,h,w] fp32 logits
  unsigned char* pred;       // [N,H,W] class index (optional)
  const int64_t* labels;     // [N,H,W] (optional, with conf)
  unsigned long long* conf;  // [C*C] += (row = label, column = prediction)
  const unsigned* cell_mask; // [N,h,w] candidate classes per low-res cell (argmax_cell_mask_kernel) or nullptr
  int N, C, h, w, H, W, staged;
  long ignore_label;
};
// Candidate classes per low-res cell: inside the cell spanned by the nodes (cy,cx)..(cy+1,cx+1) every upsampled logit
// is a convex combination of the four corner logits, and fp32 rounding of that combination is monotone in each corner
// value.  So class k can never be the FIRST maximum anywhere in the cell if some class j dominates it at all four
// corners with j < k (ties go to the lower index), or strictly with a margin far above the rounding error when j > k.
// Testing only the <= 4 corner leaders keeps it O(C); typical cells end up with one or two candidates, and the
// per-pixel kernel then interpolates just those -- the result is identical to the exhaustive loop.
template <int CMAX>
__global__ void __launch_bounds__(128) argmax_cell_mask_kernel(const float* __restrict__ x, int N, int C, int h, int w,
                                                               unsigned* __restrict__ mask) {
  const long cell = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (cell >= static_cast<long>(N) * h * w) return;
  const int cx = static_cast<int>(cell % w);
  const long t1 = cell / w;
  const int cy = static_cast<int>(t1 % h);
  const int n = static_cast<int>(t1 / h);
  const int x1 = min(cx + 1, w - 1), y1 = min(cy + 1, h - 1);
  const size_t plane = static_cast<size_t>(h) * w;
  const float* xn = x + static_cast<size_t>(n) * C * plane;
  float v[CMAX][4], best[4], lv[4][4];
  int lead[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) { best[c] = -FLT_MAX; lead[c] = 0; }
#pragma unroll
  for (int k = 0; k < CMAX; ++k) {
    if (k < C) {
      const float* q = xn + k * plane;
      v[k][0] = __ldg(q + cy * w + cx); v[k][1] = __ldg(q + cy * w + x1);
      v[k][2] = __ldg(q + y1 * w + cx); v[k][3] = __ldg(q + y1 * w + x1);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (k == 0 || v[k][c] > best[c]) {
          best[c] = v[k][c]; lead[c] = k;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) lv[c][q4] = v[k][q4];
        }
      }
    }
  }
  unsigned m = 0;
#pragma unroll
  for (int k = 0; k < CMAX; ++k) {
    if (k < C) {
      bool dominated = false;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = lead[c];
        if (j == k) continue;
        bool dom = true;
        if (j < k) {
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) dom = dom && lv[c][q4] >= v[k][q4];
        } else {
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const float margin = 1e-4f * fmaxf(1.f, fmaxf(fabsf(lv[c][q4]), fabsf(v[k][q4])));
            dom = dom && (lv[c][q4] - v[k][q4] > margin);
          }
        }
        dominated = dominated || dom;
      }
      if (!dominated) m |= 1u << k;
    }
  }
  mask[cell] = m;
}

template <int CMAX>
__global__ void __launch_bounds__(256) upsample_argmax_kernel(PostParams p) {
  extern __shared__ float lo[];                    // [C][kRLH][kRLW] when staged
  __shared__ unsigned hist[kMaxC * kMaxC];
  const int tiles_x = (p.W + kRW - 1) / kRW, tiles_y = (p.H + kRH - 1) / kRH;
  const int n = blockIdx.x / (tiles_x * tiles_y);
  const int rem = blockIdx.x - n * tiles_x * tiles_y;
  const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
  const int ly0 = lerp_ac(ty * kRH, p.h, p.H).i0, lx0 = lerp_ac(tx * kRW, p.w, p.W).i0;
  const size_t plane = static_cast<size_t>(p.h) * p.w;
  const float* xn = p.x + static_cast<size_t>(n) * p.C * plane;
  if (p.staged && !p.cell_mask) {
    for (int i = threadIdx.x; i < p.C * kRLP; i += blockDim.x) {
      const int ch = i / kRLP, r = i - ch * kRLP;
      const int ly = min(ly0 + r / kRLW, p.h - 1), lx = min(lx0 + r % kRLW, p.w - 1);
      lo[i] = __ldg(xn + ch * plane + static_cast<size_t>(ly) * p.w + lx);
    }
  }
  if (p.conf) for (int i = threadIdx.x; i < p.C * p.C; i += blockDim.x) hist[i] = 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int y = ty * kRH + warp;
  if (y < p.H) {
    const Lerp ly = lerp_ac(y, p.h, p.H);
    const float wy1 = ly.l, wy0 = 1.f - ly.l;
    const long row = (static_cast<long>(n) * p.H + y) * p.W;
#pragma unroll 1
    for (int j = 0; j < kPX; ++j) {
      const int x = tx * kRW + j * 32 + lane;
      if (x >= p.W) break;
      const Lerp lx = lerp_ac(x, p.w, p.W);
      const float wx1 = lx.l, wx0 = 1.f - lx.l;
      float best = -FLT_MAX;
      int arg = 0;
      if (p.cell_mask) {
        unsigned m = __ldg(p.cell_mask + (static_cast<size_t>(n) * p.h + ly.i0) * p.w + lx.i0);
        if ((m & (m - 1)) == 0) {
          arg = __ffs(m) - 1;                      // a single candidate: no interpolation needed
        } else {
          for (; m; m &= m - 1) {                  // ascending class index: the first maximum wins
            const int k = __ffs(m) - 1;
            const float* q = xn + k * plane;
            const float a = __ldg(q + ly.i0 * p.w + lx.i0), b = __ldg(q + ly.i0 * p.w + lx.i1);
            const float d = __ldg(q + ly.i1 * p.w + lx.i0), e = __ldg(q + ly.i1 * p.w + lx.i1);
            const float top = __fadd_rn(__fmul_rn(wx0, a), __fmul_rn(wx1, b));
            const float bot = __fadd_rn(__fmul_rn(wx0, d), __fmul_rn(wx1, e));
            const float v = __fadd_rn(__fmul_rn(wy0, top), __fmul_rn(wy1, bot));
            if (v > best) { best = v; arg = k; }
          }
        }
      }
#pragma unroll
      for (int k = 0; k < CMAX; ++k) {
        if (k < p.C && !p.cell_mask) {
          float a, b, d, e;
          if (p.staged) {
            const float* q = lo + k * kRLP;
            const int r0 = (ly.i0 - ly0) * kRLW - lx0, r1 = (ly.i1 - ly0) * kRLW - lx0;
            a = q[r0 + lx.i0]; b = q[r0 + lx.i1]; d = q[r1 + lx.i0]; e = q[r1 + lx.i1];
          } else {
            const float* q = xn + k * plane;
            a = __ldg(q + ly.i0 * p.w + lx.i0); b = __ldg(q + ly.i0 * p.w + lx.i1);
            d = __ldg(q + ly.i1 * p.w + lx.i0); e = __ldg(q + ly.i1 * p.w + lx.i1);
          }
          // h0 * (w0 * a + w1 * b) + h1 * (w0 * d + w1 * e), every product and sum rounded (no fma)
          const float top = __fadd_rn(__fmul_rn(wx0, a), __fmul_rn(wx1, b));
          const float bot = __fadd_rn(__fmul_rn(wx0, d), __fmul_rn(wx1, e));
          const float v = __fadd_rn(__fmul_rn(wy0, top), __fmul_rn(wy1, bot));
          if (v > best) { best = v; arg = k; }     // first maximum wins, like torch.argmax / np.argmax
        }
      }
      if (p.pred) p.pred[row + x] = static_cast<unsigned char>(arg);
      if (p.conf) {
        const long t = p.labels[row + x];
        if (t != p.ignore_label && t >= 0 && t < p.C) atomicAdd(&hist[static_cast<int>(t) * p.C + arg], 1u);
      }
    }
  }
  if (p.conf) {
    __syncthreads();
    for (int i = threadIdx.x; i < p.C * p.C; i += blockDim.x)
      if (hist[i]) atomicAdd(p.conf + i, static_cast<unsigned long long>(hist[i]));
  }
}

template <int CMAX>
cudaError_t launch_post(const PostParams& p, cudaStream_t st) {
  const unsigned blocks = static_cast<unsigned>(p.N) * ((p.W + kRW - 1) / kRW) * ((p.H + kRH - 1) / kRH);
  const size_t smem = (p.staged && !p.cell_mask) ? static_cast<size_t>(p.C) * kRLP * sizeof(float) : 0;
  if (p.cell_mask) {
    const long cells = static_cast<long>(p.N) * p.h * p.w;
    argmax_cell_mask_kernel<CMAX><<<static_cast<unsigned>((cells + 127) / 128), 128, 0, st>>>(p.x, p.N, p.C, p.h, p.w,
                                                                                             const_cast<unsigned*>(p.cell_mask));
  }
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(upsample_argmax_kernel<CMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
  }
  upsample_argmax_kernel<CMAX><<<blocks, 256, smem, st>>>(p);
  return cudaGetLastError();
}

// --------------------------------------------------------------------------------------- x8 upsample (returned outputs)
__global__ void __launch_bounds__(256) upsample_ac_kernel(const float* __restrict__ x, int NC, int h, int w,
                                                          float* __restrict__ out, int H, int W) {
  const long total = static_cast<long>(NC) * H * W;
  const long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int X = static_cast<int>(idx % W);
  const long t = idx / W;
  const int Y = static_cast<int>(t % H);
  const long nc = t / H;
  const Lerp ly = lerp_ac(Y, h, H), lx = lerp_ac(X, w, W);
  const float* p = x + nc * h * w;
  const float a = __ldg(p + ly.i0 * w + lx.i0), b = __ldg(p + ly.i0 * w + lx.i1);
  const float d = __ldg(p + ly.i1 * w + lx.i0), e = __ldg(p + ly.i1 * w + lx.i1);
  out[idx] = (1.f - ly.l) * ((1.f - lx.l) * a + lx.l * b) + ly.l * ((1.f - lx.l) * d + lx.l * e);
}

}  // namespace

size_t criterion_workspace_bytes(int N, int H, int W) {
  const size_t npix = static_cast<size_t>(N) * H * W;
  return npix * 9 + 4096 /*alignment*/ + A_COUNT * sizeof(double) + sizeof(SelState) + 2 * 4096 * sizeof(unsigned) + 1024;
}

cudaError_t criterion_launch(CritParams p, void* workspace, bool backward, cudaStream_t st) {
  if (p.C > kMaxC) return cudaErrorInvalidValue;
  const size_t npix = static_cast<size_t>(p.N) * p.H * p.W;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  size_t off = 0;
  p.ws_p = reinterpret_cast<float*>(ws + off); off = al(off + npix * 4);
  p.ws_ce = reinterpret_cast<float*>(ws + off); off = al(off + npix * 4);
  p.ws_flags = ws + off; off = al(off + npix);
  p.accum = reinterpret_cast<double*>(ws + off); off = al(off + A_COUNT * sizeof(double));
  SelState* sel = reinterpret_cast<SelState*>(ws + off); off = al(off + sizeof(SelState));
  unsigned* hist = reinterpret_cast<unsigned*>(ws + off);
  cudaError_t e;
  if ((e = cudaMemsetAsync(p.accum, 0, A_COUNT * sizeof(double) + 512 + 2 * 4096 * sizeof(unsigned) + 256, st)) != cudaSuccess) return e;
  if (p.aux_ce && (e = cudaMemsetAsync(p.aux_ce, 0, npix * sizeof(float), st)) != cudaSuccess) return e;
  const unsigned blocks = static_cast<unsigned>((npix + 255) / 256);
  const bool run_path = column_runs_fit(p);
  if (run_path) { if ((e = launch_run(p, nullptr, false, st)) != cudaSuccess) return e; }
  else if (p.C <= 12) crit_pixel_kernel<12><<<blocks, 256, 0, st>>>(p);
  else if (p.C <= 20) crit_pixel_kernel<20><<<blocks, 256, 0, st>>>(p);
  else crit_pixel_kernel<kMaxC><<<blocks, 256, 0, st>>>(p);
  crit_select_init_kernel<<<1, 32, 0, st>>>(p, sel);
  const unsigned hb = blocks < 1184 ? blocks : 1184;
  // exact k-th order statistic of p: radix select over the 32-bit pattern (12 + 12 + 8 bits)
  crit_hist_kernel<<<hb, 256, 0, st>>>(p, sel, hist, 20, 12, 0);
  crit_pick_kernel<<<1, 256, 0, st>>>(sel, hist, 12, 0, p.ohem_thres);
  crit_hist_kernel<<<hb, 256, 0, st>>>(p, sel, hist, 8, 12, 12);
  crit_pick_kernel<<<1, 256, 0, st>>>(sel, hist, 12, 0, p.ohem_thres);
  crit_hist_kernel<<<hb, 256, 0, st>>>(p, sel, hist, 0, 8, 24);
  crit_pick_kernel<<<1, 256, 0, st>>>(sel, hist, 8, 1, p.ohem_thres);
  crit_sum_kernel<<<hb, 256, 0, st>>>(p, sel);
  crit_final_kernel<<<1, 32, 0, st>>>(p, sel);
  if (backward) {
    const size_t lowres = static_cast<size_t>(p.N) * p.h * p.w;
    if ((e = cudaMemsetAsync(p.g_p, 0, lowres * p.C * 4, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(p.g_m, 0, lowres * p.C * 4, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(p.g_d, 0, lowres * 4, st)) != cudaSuccess) return e;
    if (run_path) return launch_run(p, sel, true, st);
    const int tiles = ((p.W + kTileW - 1) / kTileW) * ((p.H + kTileH - 1) / kTileH);
    // low-res footprint of a 32 x 8 label tile (+1 for the i1 neighbour, +1 for the fractional start)
    auto span = [](int t, int in, int out) { return out > 1 ? (static_cast<long>(t - 1) * (in - 1)) / (out - 1) + 3 : 1; };
    p.direct_scatter = (span(kTileW, p.w, p.W) > kLW || span(kTileH, p.h, p.H) > kLH) ? 1 : 0;
    const size_t smem = static_cast<size_t>(2 * p.C + 1) * kLH * kLW * sizeof(float);
    if (p.direct_scatter) {
      crit_backward_kernel<<<p.N * tiles, 256, smem, st>>>(p, sel);
    } else if (p.C <= 12) {
      crit_backward_tiled_kernel<12><<<p.N * tiles, 256, 2 * smem, st>>>(p, sel);
    } else if (p.C <= 20) {
      crit_backward_tiled_kernel<20><<<p.N * tiles, 256, 2 * smem, st>>>(p, sel);
    } else {
      crit_backward_tiled_kernel<kMaxC><<<p.N * tiles, 256, 2 * smem, st>>>(p, sel);
    }
  }
  return cudaGetLastError();
}

cudaError_t postprocess_launch(const float* x, int N, int C, int h, int w, int H, int W, unsigned char* pred,
                               const int64_t* labels, long ignore_label, unsigned long long* conf, unsigned* cell_mask_ws,
                               cudaStream_t st) {
  if (C < 1 || C > kMaxC) return cudaErrorInvalidValue;
  PostParams p;
  p.x = x; p.pred = pred; p.labels = labels; p.conf = conf; p.cell_mask = cell_mask_ws; p.N = N; p.C = C; p.h = h; p.w = w; p.H = H; p.W = W;
  p.ignore_label = ignore_label;
  CritParams fp;   // footprint test only
  fp.h = h; fp.w = w; fp.H = H; fp.W = W;
  p.staged = footprint_fits(fp, kRW, kRH, kRLW, kRLH) ? 1 : 0;
  if (C <= 12) return launch_post<12>(p, st);
  if (C <= 20) return launch_post<20>(p, st);
  return launch_post<kMaxC>(p, st);
}

cudaError_t upsample_ac_launch(const float* x, int NC, int h, int w, float* out, int H, int W, cudaStream_t st) {
  const long total = static_cast<long>(NC) * H * W;
  upsample_ac_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(x, NC, h, w, out, H, W);
  return cudaGetLastError();
}

}  // namespace pidnet
