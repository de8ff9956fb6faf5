// HBM-bound fused kernels of the PIDNet path + the SIMT reference conv (debug / cross-check path).
// All activations are NHWC bf16 with an explicit pixel stride so that channel slices of a wider
// (concat) buffer can be read/written in place; math is fp32, one rounding per stored value.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cuda.h>
#include "conv_tc.cuh"

namespace pidnet {

typedef __nv_bfloat16 bf16;

// A strided NHWC view (ptr already includes the channel offset).
struct View {
  bf16* ptr;
  int N, H, W, C;
  long ps;  // pixel stride in elements (>= C)
};

// ---- stem: fp32 NCHW image -> 3x3 s2 p1 conv (+bias, BN folded) + ReLU -> bf16 NHWC   (pidnet.py:25-27)
// w: [27][Cout] fp32 with k = (ci*3 + r)*3 + s ; bias: [Cout]
cudaError_t stem_conv_launch(const float* x, int N, int H, int W, View out, const float* w, const float* bias,
                             cudaStream_t st);

// ---- the same stem on tcgen05 (software im2col A-producer; stem_tc.cu). Cout in {32, 64}.
struct StemParams {
  CUtensorMap tmD;      // output as flat [rows = N*Ho*Wo][Cout] bf16, box {Cout, 128, 1, 1}, swizzle = Cout*2 bytes
  const float* x;       // fp32 NCHW image
  // alternative input: uint8 HWC BGR frames as cv2.imread delivers them; the reference's
  // input_transform (datasets/base_dataset.py:36-44: [..., ::-1] / 255 - mean, / std) is applied on load
  const uint8_t* x_u8;  // [N,H,W,3] or nullptr
  const float* lut;     // [3][256] device table: normalised value of byte b for model (RGB) channel c
  const uint8_t* w_swz; // [Cout][32] bf16 K-major (k = (ci*3+r)*3+s, zero padded), pre-swizzled SWIZZLE_64B
  const float* bias;    // [Cout]
  int H, W, Ho, Wo;
  long rows, tiles;     // N*Ho*Wo, ceil(rows / 128)
  int relu;             // 1: inference (bias + ReLU); 0: training (raw conv output, BatchNorm follows)
  double* stats;        // training: [kBnStatReplicas][2 * Cout] fp64 table receiving sum x / sum x^2 of the stored output (or null)
};
cudaError_t stem_tc_launch(const StemParams& p, int Cout, int num_sms, cudaStream_t st);
// ---- fused stem (stem2_tc.cu): conv1.0+BN+ReLU -> conv1.3+BN+ReLU, the C-channel half-resolution intermediate stays in smem
struct Stem2Params {
  CUtensorMap tmD;        // conv1.3 output [N,H2,W2,C] bf16 NHWC, box {C, 8, 16, 1}, swizzle = C*2 bytes
  CUtensorMap tmX;        // pipelined kernel: the fp32 NCHW image as [N,3,H,W], box {36, 67, 3, 1}, no swizzle
  const float* x;         // fp32 NCHW image (or null)
  const uint8_t* x_u8;    // uint8 HWC BGR frames (or null), normalised through `lut` as in StemParams
  const float* lut;
  const uint8_t* w1_swz;  // conv1.0: [C][32] bf16 K-major, pre-swizzled SWIZZLE_64B (== StemParams::w_swz)
  const float* bias1;     // [C]
  const uint8_t* w2_swz;  // conv1.3: [9 taps][C out][C in] bf16 K-major tiles, pre-swizzled (C*2-byte rows)
  const float* bias2;     // [C]
  int H, W;               // image
  int H1, W1;             // conv1.0 output (ceil(H/2), ceil(W/2))
  int H2, W2;             // conv1.3 output
  int tiles_w, tiles_h, N;  // 16 x 8 output tiles
};
cudaError_t stem2_tc_launch(const Stem2Params& p, int C, int num_sms, int pipelined, cudaStream_t st);
// training: fp32 [Cout][3][3][3] master weights -> the pre-swizzled bf16 [Cout][32] tile of the kernel above
cudaError_t stem_pack_launch(const float* w, uint8_t* w_swz, int Cout, cudaStream_t st);

// ---- PagFM fuse (model_utils.py:292-312 after the low-res algebra of DESIGN.md):
//   low = [y | z | t | pad] at (h,w);  s = <x, U(z)> + U(t);  g = sigmoid(s);  out = relu((1-g) x + g U(y))
cudaError_t pag_fuse_launch(View x, View low, View out, int relu, cudaStream_t st);

// ---- generic elementwise: out = act(s * (a + U(b)) + t)   (a, b, s/t optional; U = bilinear, align_corners=False)
cudaError_t upadd_launch(View a, View b, View out, const float* s, const float* t, int relu, cudaStream_t st);
// same with a residual added AFTER the affine map: out = act(s * (a + U(b)) + t + r)
cudaError_t upadd_res_launch(View a, View b, View r, View out, const float* s, const float* t, int relu, cudaStream_t st);

// ---- average pool (count_include_pad) + affine + ReLU; k == 0 means global average pool
// n / d through the launch-time magic-number division the elementwise kernels decode their thread index with (host copy)
uint32_t fastdiv_debug(uint32_t n, uint32_t d);

// up to five upsample-add jobs of identical output geometry in ONE launch (PAPPM: relu(bn_k(scale0 + U(scale_k))), k = 1..4,
// and relu(bn_0(scale0)) as the job "null + U(scale0)" whose interpolation is the identity)
cudaError_t upadd_batch_launch(int njobs, const View* a, const View* b, const View* out, const float* const* s,
                               const float* const* t, int relu, cudaStream_t st);

// ---- the four pooled branches of PAPPM / DAPPM (AvgPool 5/2/2, 9/4/4, 17/8/8 and the global pool, each + BN + ReLU) from ONE
// pass over x: a block builds the summed-area table of a 32-channel slice of one image in shared memory (fp32) and every
// pooled output is four table look-ups.  out[0..2]: k = 5, 9, 17; out[3]: 1x1.  s / t: the four BN affines, [4][C].
// Returns cudaErrorNotSupported when the table does not fit in shared memory (caller falls back to pool_affine_launch).
// ew / se / te (optional): two full-resolution by-products relu(se[j] * x + te[j]) written in the same pass (PAPPM's
// scale0 and shortcut operands), so x is read once for all six consumers.
cudaError_t pool_pyramid_launch(View x, const View out[4], const float* s, const float* t, const View ew[2], const float* se,
                                const float* te, cudaStream_t st);
cudaError_t pool_affine_launch(View x, View out, int k, int stride, int pad, const float* s, const float* t, int relu,
                               cudaStream_t st);

// ---- Light_Bag operand producer (model_utils.py:328-334): e = sigmoid(d);
//   out[..., 0:C] = (1-e) U(i) + p ;  out[..., C:2C] = U(i) + e p
cudaError_t lightbag_uv_launch(View p, View i_low, View d, View out, cudaStream_t st);

// ---- Bag blend (model_utils.py:375-377) + the pre-activation BN/ReLU of dfm.conv:
//   out = relu(s * (e p + (1-e) U(i)) + t)
cudaError_t bag_blend_launch(View p, View i_low, View d, View out, const float* s, const float* t, cudaStream_t st);

// ---- SIMT restatement of conv_tc's contract on raw pointers (same packed weights / tap tables).
struct RefMap {
  const bf16* ptr;
  int C, W, H, N;
  long sW, sH, sN;  // strides in elements
};
struct ConvRefParams {
  RefMap maps[kConvMaxMaps];
  ConvSrc src[2];
  int nsrc, BK;
  const bf16* wpk;  // [Cout_pad][Ktot]
  long Ktot;
  const float* bias;
  RefMap res;       // residual view (ptr == nullptr -> none)
  bf16* out;        // NHWC view
  long o_sW, o_sH, o_sN;
  float* out_f32;   // NCHW
  int N, Ho, Wo, Cout, relu, out_mode;
};
cudaError_t conv_ref_launch(const ConvRefParams& p, cudaStream_t st);

}  // namespace pidnet
