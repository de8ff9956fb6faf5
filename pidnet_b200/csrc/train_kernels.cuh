// Training-mode kernels (see train_kernels.cu, wgrad_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "kernels.cuh"

namespace pidnet {

// ---- BatchNorm with batch statistics
// sums: device double[2*C] scratch (sum, sum of squares | sum dz', sum dz'*xhat)
// `sums` must be zero on entry (the arena is cleared at plan time; finalize / backward / add_sums clear it after use)
cudaError_t bn_stats_launch(View x, double* sums, int num_sms, cudaStream_t st);
// mean / invstd / folded (scale, shift) for the apply kernel; running stats updated in place when non-null
// (momentum 0.1, unbiased variance; `conv_bias` is added to the tracked mean for convs whose bias we drop)
cudaError_t bn_finalize_launch(double* sums, int C, double count, const float* gamma, const float* beta,
                               const float* conv_bias, float* mean, float* invstd, float* scale, float* shift,
                               float* run_mean, float* run_var, cudaStream_t st);
// z = act(scale*x + shift (+res)) is the existing upadd kernel (kernels.cuh).  Backward:
//   dz' = dz*[z>0];  dgamma += sum dz'*xhat;  dbeta += sum dz';  dx (+)= gamma*invstd*(dz' - mean(dz') - xhat*mean(dz'*xhat));
//   dres (+)= dz'
cudaError_t bn_backward_launch(View x, View dz, View z, View dx, View dres, const float* mean, const float* invstd,
                               const float* gamma, double* sums, float* coef /*[3*C] scratch*/, int relu, int acc_dx,
                               int acc_dres, float* dgamma, float* dbeta, int num_sms, cudaStream_t st);

// Fused single-launch forms (persistent grid + grid barrier, loaded vectors parked in shared memory between the
// statistics pass and the apply pass).  Each launch gets its own accumulator block of bn_fused_acc_bytes(C) bytes (`sums` at its
// start, `sync` at bn_fused_sync_offset(C)); the kernels never clear it -- the caller zeroes all blocks of a pass with one memset.
bool bn_fused_supported(int C);
size_t bn_fused_acc_bytes(int C);
size_t bn_fused_sync_offset(int C);
// `scale` / `shift` ([C] each): the folded affine the forward applied, kept for the backward; mask_x = 1 (ReLU without a residual):
// the backward recomputes the ReLU mask from x with that affine instead of reading the stored output z.
cudaError_t bn_forward_fused_launch(View x, View res, View z, const float* gamma, const float* beta, const float* conv_bias,
                                    float* mean, float* invstd, float* scale, float* shift, float* run_mean, float* run_var,
                                    double* sums, unsigned* sync, int relu, int num_sms, cudaStream_t st, int pre = 0);
// (pre = 1: `sums` was filled by the store warp of the conv that produced x -- Conv3Params::stats -- so the launch only derives
// the coefficients and applies them: one read of x, no grid barrier, no cooperative launch)
cudaError_t bn_backward_fused_launch(View x, View dz, View z, View dx, View dres, const float* mean, const float* invstd,
                                     const float* gamma, const float* scale, const float* shift, int mask_x, double* sums,
                                     unsigned* sync, int relu, int acc_dx, int acc_dres, float* dgamma, float* dbeta, int num_sms,
                                     cudaStream_t st);

// ---- fused SGD step over flat fp32 buffers (n a multiple of 4, pointers 16-byte aligned); `first` = no momentum history yet
cudaError_t sgd_step_launch(float* p, const float* g, float* buf, long n, float lr, float momentum, float dampening, float wd,
                            int nesterov, int first, float grad_scale, cudaStream_t st);

// ---- device-side weight packing: fp32 [Cout][Cin_total][k][k] -> bf16 packed rows (forward: row = co; dgrad: row = ci)
struct PackJob {
  const float* src;
  bf16* dst;
  long Ktot, kofs;        // destination row length / offset of this source's K range
  int rows_pad;           // destination rows (zero-filled beyond the valid ones)
  int Cout, Cin, Cin_total, ci_off, k;
  int BK, chunks, ntaps;
  int dgrad;
  unsigned char taps[9];  // (r << 4) | s of the SOURCE weight tap feeding destination tap i
};
cudaError_t pack_weights_launch(const PackJob& j, cudaStream_t st);
// batched form: job table + first-block prefix in device memory, one launch for every conv of the step
unsigned pack_job_blocks(const PackJob& j);
cudaError_t pack_all_launch(const PackJob* dev_jobs, const unsigned* dev_block_start, int njobs, unsigned total_blocks,
                            cudaStream_t st);

cudaError_t nchw_to_nhwc_launch(const float* x, int N, int C, int H, int W, View out, cudaStream_t st);
// dlow (+)= U^T dhi  (bilinear, align_corners=False);  dx (+)= P^T dy (AvgPool count_include_pad; k==0 global)
cudaError_t upsample_transpose_launch(View dhi, View dlow, int accumulate, cudaStream_t st);
cudaError_t pool_transpose_launch(View dy, View dx, int k, int stride, int pad, int accumulate, cudaStream_t st);
// out (+)= a * [mask > 0]  (mask optional)
cudaError_t masked_add_launch(View a, View mask, View out, int accumulate, cudaStream_t st);

// ---- PagFM / Light_Bag in training form (gate saved for backward)
cudaError_t pag_train_fwd_launch(View x, View xk, View yq, View y, View out, float* gate, cudaStream_t st);
cudaError_t pag_train_bwd_launch(View x, View xk, View yq, View y, View out, View dout, const float* gate, View dx,
                                 int acc_dx, View dxk, View t1, View t2, cudaStream_t st);
cudaError_t lightbag_bwd_launch(View p, View il, View d, View duv, View dp, int acc_dp, View dd, int acc_dd, View ti,
                                cudaStream_t st);
// Bag (L topology), train mode: out = e p + (1-e) U(i) with e = sigmoid(d); backward: dp (+)= e g, dd (+)= g (p - U(i)) e (1-e),
// ti = (1-e) g (hi-res gradient of U(i), reduced to the low-res map by upsample_transpose)
cudaError_t bag_train_fwd_launch(View p, View i_low, View d, View out, cudaStream_t st);
cudaError_t bag_train_bwd_launch(View p, View i_low, View d, View dout, View dp, int acc_dp, View dd, int acc_dd, View ti,
                                 cudaStream_t st);
// dst[c] += sums[c]  (bias gradients from a per-channel reduction)
cudaError_t add_sums_launch(double* sums, float* dst, int C, int Cacc, cudaStream_t st);

// ---- weight gradient of the 3-channel stem conv (conv1.0: 3x3, stride 2, pad 1) straight from the caller's fp32 NCHW image:
//   dW[co][ci][r][s] = sum_{n,oh,ow} dY[n,oh,ow,co] * x[n,ci,2oh-1+r,2ow-1+s]      (overwrites dW: [Cout][3][3][3] fp32)
// K = pixels is the whole problem (5.4 GFLOP at 12 x 1024^2): a register-tiled SIMT kernel (one warp owns the full 32 x 27
// gradient of a 32-channel slice for its pixel chunks; 4 co x 8 k accumulators per thread) instead of the tcgen05 wgrad on
// an 8-channel padded NHWC copy of the image (0.52 ms + 0.11 ms for the copy).
cudaError_t stem_wgrad_launch(const float* x, int N, int H, int W, View dy, float* dW, int num_sms, cudaStream_t st);

// ---- wgrad on tcgen05 with MN-major operands (wgrad_tc.cu):
//   dW[co][ci_off+ci][r][s] += sum_pixels dY[p][co] * X[tap(p)][ci]       (fp32 atomics into the torch-layout gradient)
struct WgradParams {
  CUtensorMap tmY;                 // dY  [N,Ho,Wo,Cout]  box {64, 8, 8, 1}
  CUtensorMap tmX[kConvMaxMaps];   // X (parity) maps      box {64, 8, 8, 1}
  uint32_t taps[kConvMaxTaps];     // map | (dh+8) << 8 | (dw+8) << 16 | (r << 24) | (s << 28)
  int ntaps;
  int tiles_w, tiles_h, N;         // 8x8-pixel tiles of the OUTPUT grid
  int Cout, Cin, Cin_total, ci_off, k;
  float* dW;                       // [Cout][Cin_total][k][k] fp32
  float* ws;                       // split-K partials: [split][co tile][z][128 rows][T taps][64 ci] fp32
  int halo;                        // 1: 3x3 stride-1 halo-patch kernel (16x8-pixel tiles: tmY box {64,16,8,1}, tmX[0] box {64,18,10,1},
                                   //    5 taps per CTA, grid.z = ci tiles * 2)
  int direct;                      // 1x1 convs: every CTA adds its partial tile straight into dW with 16-byte vector reductions
                                   //    (red.global.add.v4.f32: a thread's 32 consecutive ci of one co row are contiguous in dW);
                                   //    no workspace, no reduction launch
};
constexpr int kWgradTileFloats = 128 * 64;   // per tap
struct WgradLaunch {
  WgradParams p;
  dim3 grid;  // (pixel splits, co tiles, ci tiles * tap groups)
  int taps_per_group;
};
// defer_reduce: leave the split-K partial tiles in L.p.ws (the caller sums many layers with ONE wgrad_reduce_all_launch)
cudaError_t wgrad_tc_launch(const WgradLaunch& L, cudaStream_t st, bool defer_reduce = false);
cudaError_t wgrad_tc_init();

// One deferred split-K reduction: everything the reduction of a layer needs, without the tensor maps of WgradParams
struct WgReduceJob {
  const float* ws;
  float* dW;
  int Cout, Cin, Cin_total, ci_off, k, ntaps;
  unsigned char tap_rs[kConvMaxTaps];   // (r << 4) | s of tap i
  int stack;                            // 1: partial tiles of wgrad_stack_kernel, 0: of wgrad_tc_kernel<T> / wgrad_halo_kernel
  int T, splits, co_tiles, nz;          // taps per CTA, grid.x, grid.y, grid.z
};
bool wgrad_reduce_job(const WgradLaunch& L, WgReduceJob* job);   // false: this launch reduces itself (direct mode)
unsigned wgrad_reduce_job_blocks(const WgReduceJob& j);
// jobs / block_start: device arrays; block_start[i] = first block of job i in a numbering where job 0 starts at `base`
cudaError_t wgrad_reduce_all_launch(const WgReduceJob* dev_jobs, const unsigned* dev_block_start, int njobs, unsigned base,
                                    unsigned total_blocks, cudaStream_t st);

}  // namespace pidnet
