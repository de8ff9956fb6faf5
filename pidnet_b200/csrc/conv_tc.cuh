// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a), NHWC bf16 -> fp32 TMEM accumulators.
//
//   D[pixel, cout] = act( sum_{src, tap, cin} A_src[pixel + tap, cin] * W[cout, (src,tap,cin)] + bias[cout] (+ R[pixel, cout]) )
//
// * M tile = 128 output pixels = a TN x TH x TW patch of the NHWC output; one TMA 4-D box load per
//   (tap, 64/32-channel chunk) brings the shifted input patch (zero-filled outside the image ==
//   the conv padding) into a 128-row K-major SWIZZLE_128B/64B tile that tcgen05.mma consumes directly.
// * stride-2 convs read one of four "parity" tensor maps (the even/odd sub-lattices of the input),
//   so every load is a plain dense box; 1x1 convs are the 1-tap case; a second source appends
//   K-slices (the 1x1 downsample of a residual block, or a channel concat) to the same accumulator.
// * BatchNorm is folded into W/bias on the host; bias, residual add and ReLU run on the TMEM->register
//   epilogue; the bf16 tile is staged in swizzled smem and written with TMA (clipped at ragged edges),
//   or written as fp32 NCHW planes for the logits.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace pidnet {

constexpr int kConvMaxMaps = 6;
constexpr int kConvMaxTaps = 9;

struct ConvSrc {
  int ntaps;                    // 1 (1x1) or up to 9 (3x3)
  int chunks;                   // ceil(Cin / BK)
  uint32_t taps[kConvMaxTaps];  // map index | (dh+8) << 8 | (dw+8) << 16
};

// kOutNHWCsplit: the fp32 result y is stored as TWO bf16 values, hi = bf16(y) in channels [0, Cout) and lo = bf16(y - hi) in
// channels [Cout, 2 Cout) of a 2*Cout-channel NHWC tensor (split-bf16 storage, ~16 mantissa bits; the "fp32_head" mode)
enum ConvOutMode { kOutNHWCbf16 = 0, kOutNCHWf32 = 1, kOutNHWCsplit = 2 };

struct ConvParams {
  CUtensorMap tmA[kConvMaxMaps];  // activation maps
  CUtensorMap tmB;                // packed weights [Cout_pad][Ktot], K-major
  CUtensorMap tmR;                // residual (same geometry as tmD)
  CUtensorMap tmD;                // output
  ConvSrc src[2];
  int nsrc;
  int tiles_w, tiles_h;  // tiles per (TN-image group) along W and H
  int TW, TH, TN;
  int N, Ho, Wo, Cout;
  int relu, has_res, out_mode;
  const float* bias;  // [Cout_pad] fp32
  float* out_f32;     // kOutNCHWf32 destination [N][Cout][Ho][Wo]
};

struct ConvLaunch {
  ConvParams p;
  int BN, BK;
  dim3 grid;
};

// ---- conv3x3 stride-1 weight-stationary / halo-patch / persistent kernel (conv3_ws.cu)
constexpr int kConv3MaxSmem = 232448;  // 227 KB opt-in limit per CTA on sm_100

struct Conv3Params {
  CUtensorMap tmA;   // input  [N,H,W,C]  box {CK, 10, 18, 1}   (mode 1: flat [rows, C], box {CK, 128, 1, 1})
  CUtensorMap tmA2;  // mode 1: second K-concatenated source
  CUtensorMap tmW;   // packed weights [Cout_pad][9*chunks*CK], box {CK, BN} (CTA pairs: box {CK, BN/2})
  CUtensorMap tmR;   // residual, box {min(BN,64), 8, 16, 1}
  CUtensorMap tmD;   // output,   box {min(BN,64), 8, 16, 1}
  int chunks;        // ceil(Cin / CK), summed over the sources
  int chunks0;       // mode 1: chunks of the first source
  int npatch;        // halo patch ring depth
  int nstage;        // output / residual staging buffers (2 or 3)
  int tiles_w, tiles_h, N;
  int Cout, Ho, Wo;
  int relu, has_res, out_mode;
  const float* bias;
  float* out_f32;
  // training: per-channel batch statistics of the stored (bf16-rounded) output, accumulated by the store warp from the staged
  // tiles -- sum x into stats[r][c], sum x^2 into stats[r][stats_C + c] for replica r = blockIdx.x % kBnStatReplicas of a
  // [kBnStatReplicas][2 * stats_C] fp64 table (the accumulator block of the BatchNorm that follows, train_kernels.cu); or null
  double* stats;
  int stats_C;
};
constexpr int kBnStatReplicas = 8;

struct Conv3Launch {
  Conv3Params p;
  int BN, CK;
  int mode;           // 0: 3x3 stride-1 halo-patch   1: 1x1 stride-1 over flattened pixels
  dim3 grid;          // (persistent CTAs per Cout tile, Cout tiles)
  size_t smem_bytes;
  int pair;           // 1: CTA pairs (2-CTA clusters along x, tcgen05 cta_group::2); grid.x is even
};

cudaError_t conv3_ws_launch(const Conv3Launch& L, cudaStream_t stream);
cudaError_t conv3_ws_init();
int conv3_ws_plan(int mode, int BN, int CK, int chunks, int pair, int want_stages, int* nstage, size_t* smem_bytes);
bool conv3_ws_pair_available(int mode, int BN, int CK);

// Launches the kernel instance for (BN, BK); returns cudaError_t.
cudaError_t conv_tc_launch(const ConvLaunch& L, cudaStream_t stream);
// One-time: opt in to large dynamic shared memory for all instances.
cudaError_t conv_tc_init();
size_t conv_tc_smem_bytes(int BN, int BK);

}  // namespace pidnet
