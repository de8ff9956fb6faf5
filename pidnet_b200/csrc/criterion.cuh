// Fused criterion (OhemCrossEntropy + BondaryLoss + FullModel composition) -- see criterion.cu.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace pidnet {

struct CritParams {
  const float* x_p;   // x_extra_p logits, fp32 NCHW [N,C,h,w]
  const float* x_m;   // x_ logits
  const float* x_d;   // x_extra_d logits [N,1,h,w]
  const int64_t* labels;  // [N,H,W], ignore_label marks ignored pixels
  const float* bd_gt;     // [N,H,W] 0/1
  const float* class_w;   // [C] or nullptr
  int N, C, h, w, H, W;
  long ignore_label;
  float ohem_thres, bd_threshold;
  long min_kept;
  double bw0, bw1, sb, coeff_bce;
  float* out;             // device float[12]
  float *g_p, *g_m, *g_d; // low-res gradients of loss.mean() (backward only)
  float* aux_ce;          // optional [N,H,W]: per-pixel weighted CE of the aux head (0 at ignored pixels) -- the reference's loss MAP
  int direct_scatter;     // backward: 1 = global atomics (footprint does not fit the smem tile)
  // workspace (filled by criterion_launch)
  float* ws_p;
  float* ws_ce;
  unsigned char* ws_flags;
  double* accum;
};

size_t criterion_workspace_bytes(int N, int H, int W);
cudaError_t criterion_launch(CritParams p, void* workspace, bool backward, cudaStream_t st);
// fused x8 upsample (align_corners=True) + argmax -> uint8 labels (optional) and/or confusion-matrix accumulation (optional);
// cell_mask_ws: optional device scratch of N*h*w uint32 enabling the candidate-pruned path (same result, fewer interpolations)
cudaError_t postprocess_launch(const float* x, int N, int C, int h, int w, int H, int W, unsigned char* pred,
                               const int64_t* labels, long ignore_label, unsigned long long* conf, unsigned* cell_mask_ws,
                               cudaStream_t st);
cudaError_t upsample_ac_launch(const float* x, int NC, int h, int w, float* out, int H, int W, cudaStream_t st);

}  // namespace pidnet
