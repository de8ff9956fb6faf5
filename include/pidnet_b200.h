/* pidnet_b200 -- C ABI of the B200-native PIDNet forward engine (libpidnet_b200.so).
 *
 * The reference (Bzdeco/pidnet) has no FFI layer: its operator boundary for this path is the
 * nn.Module call protocol of models/pidnet.py.  Each entry point below names the reference
 * interface it stands in for; the Python shim `pidnet_b200/pidnet.py` keeps the nn.Module surface
 * (same class / constructors / state_dict keys / output contract) and forwards to these symbols
 * through ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions: every function returns 0 on success or a negative error code; the message is
 * available from pidnet_last_error() (thread-local).  Nothing throws across the boundary.  Device
 * pointers are plain `void*` / `float*` in the caller's CUDA context (primary context of the current
 * device); `stream` is a cudaStream_t passed as void*.  A handle is bound to one device and is not
 * re-entrant (one process / one handle per GPU -- the reference's DataParallel thread-per-replica
 * model, tools/train.py:136, is replaced by one process per GPU).
 */
#ifndef PIDNET_B200_H_
#define PIDNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pidnet_engine pidnet_engine;

/* Constructor arguments of the reference model: PIDNet.__init__(m, n, num_classes, planes, ppm_planes,
 * head_planes, augment), models/pidnet.py:19.  get_pred_model / get_seg_model (pidnet.py:184-227) map
 * model names onto these. */
typedef struct pidnet_cfg {
  int m;           /* 2 (S/M) or 3 (L) */
  int n;           /* 3 (S/M) or 4 (L) */
  int num_classes; /* 19 Cityscapes, 11 CamVid */
  int planes;      /* 32 S, 64 M/L */
  int ppm_planes;  /* 96 S/M, 112 L */
  int head_planes; /* 128 S/M, 256 L */
  int augment;     /* 1: forward returns [x_extra_p, x_, x_extra_d] (pidnet.py:177-180) */
} pidnet_cfg;

const char* pidnet_last_error(void);
/* ABI version of this header (bumped on any signature change). */
int pidnet_abi_version(void);
/* Test hook (host only): n / d computed with the launch-time magic-number division the elementwise kernels use to decode
 * their thread index; must equal n / d for every 32-bit n and every d >= 1 (tests/test_host.py). */
unsigned pidnet_debug_fastdiv(unsigned n, unsigned d);

/* replaces PIDNet.__init__ (models/pidnet.py:19-100): creates an engine for one model topology. */
int pidnet_create(const pidnet_cfg* cfg, pidnet_engine** out);
int pidnet_destroy(pidnet_engine* h);

/* replaces nn.Module.load_state_dict for one tensor (models/pidnet.py:193-214, tools/eval.py:68-78):
 * `key` is the reference state_dict key (e.g. "layer3.0.conv1.weight"),
 * `host_data` a contiguous fp32 HOST array of `shape[0..ndim)`.  Integer buffers
 * (num_batches_tracked) are not needed and are ignored.  Marks the plan dirty. */
int pidnet_set_param(pidnet_engine* h, const char* key, const float* host_data, const int64_t* shape, int ndim);

/* Eval-mode plan for input [N,3,H,W] (H, W multiples of 8; pidnet.py:138-139): folds every BatchNorm
 * (eps 1e-5, model_utils.py:8) into its conv in fp32/fp64 on the host, packs weights K-major bf16,
 * lays activations out in one HBM arena (NHWC bf16), encodes the TMA tensor maps.  `arena_bytes`
 * (optional) receives the device bytes the engine allocated. */
int pidnet_plan(pidnet_engine* h, int N, int H, int W, size_t* arena_bytes);

/* replaces PIDNet.forward (models/pidnet.py:136-182) in eval mode.
 *   x_nchw   : device fp32 [N,3,H,W]
 *   out_main : device fp32 [N,num_classes,H/8,W/8]   (x_)
 *   out_p    : device fp32 [N,num_classes,H/8,W/8]   (x_extra_p; augment only, else NULL)
 *   out_d    : device fp32 [N,1,H/8,W/8]             (x_extra_d; augment only, else NULL)
 * All work is enqueued on `stream` (side streams fork/join on it with events); no host sync and no
 * allocation inside.  use_graph != 0 replays a CUDA graph captured for these pointers. */
int pidnet_forward(pidnet_engine* h, void* stream, const float* x_nchw, float* out_main, float* out_p, float* out_d,
                   int use_graph);

/* the same forward fed with camera frames as cv2.imread delivers them.
 *   bgr_hwc  : device uint8 [N,H,W,3], channel order B,G,R
 *   mean_rgb, std_rgb : host double[3] indexed R,G,B (datasets/base_dataset.py:27-28 defaults 0.485/0.456/0.406, 0.229/0.224/0.225)
 * The stem kernel applies the reference's input_transform (datasets/base_dataset.py:36-44, tools/custom.py:52-57:
 * image[..., ::-1] / 255.0 - mean, / std, HWC -> CHW) on load through a 3x256 table evaluated with numpy's exact
 * arithmetic (float32 division, then float64 subtract / divide each rounded back to float32), so the result equals
 * pidnet_forward on the host-normalised fp32 NCHW tensor while the H2D copy shrinks 4x (6.3 MB vs 25.2 MB per 1024x2048 frame). */
int pidnet_forward_u8(pidnet_engine* h, void* stream, const unsigned char* bgr_hwc, const double* mean_rgb, const double* std_rgb,
                      float* out_main, float* out_p, float* out_d, int use_graph);

/* Kernel launches issued by one forward (for bench.py's `gpu_launches`). */
int pidnet_num_launches(pidnet_engine* h);
/* Algorithmic conv FLOPs (2*MACs, no padding waste) of one forward at the planned shape. */
double pidnet_conv_flops(pidnet_engine* h);

/* Measurement support (bench.py roofline): runs the planned ops ONE AT A TIME on `stream` with a CUDA
 * event pair around each launch and returns the device time of every launch in ms (cap >= number of
 * launches).  pidnet_op_info describes launch i: its layer name, kernel family label, algorithmic
 * FLOPs (2*MACs, no padding) and algorithmic bytes (each distinct input and the output once). */
int pidnet_profile(pidnet_engine* h, void* stream, const float* x_nchw, float* out_main, float* out_p, float* out_d,
                   float* ms_per_op, int cap);
int pidnet_op_info(pidnet_engine* h, int i, char* name, int name_cap, char* kernel, int kernel_cap, double* flops,
                   double* bytes, int* lane);

/* Options (set before pidnet_plan): "conv_impl" = 0 tcgen05 (default) | 1 SIMT restatement (debug
 * cross-check); "lanes" = 1 | 3 concurrent branch streams (default 3); "use_ws" = 1 (default) | 0: use
 * the weight-stationary halo-patch kernel for 3x3 stride-1 convs; "use_pair" = 1 (default: CTA pairs /
 * tcgen05 cta_group::2 for 3x3 layers with Cin >= 128) | 0 | 2 (wherever the pair instance exists);
 * "ws_stages" = 3 (default) | 2 staging buffers of the weight-stationary kernels; "use_stem2" = 2
 * (default: conv1.0 -> conv1.3 fused in one kernel, warp-specialised pipeline for 32-channel stems) | 1 (lock-step
 * fused kernel) | 0 (two kernels); "use_pyramid" = 1 (default: the pooled PAPPM / DAPPM branches from one summed-area-table
 * kernel) | 0 (one pooling kernel per branch); "fp32_head" = 0 (default) | 1: the final segmentation head (final_layer: the
 * logits path, model_utils.py:100-112) in split-bf16 arithmetic -- fp32 weights as hi + lo bf16 pairs, the hidden tensor stored
 * as hi + lo, three K-concatenated tcgen05 GEMM terms with fp32 accumulation: given its input the head then agrees with fp32
 * arithmetic to ~1e-5 (north star: "fp32 logits within 1e-3"; the rest of the network stays bf16).  The same switches can be set for a whole
 * process with PIDNET_WS_PAIR / PIDNET_WS_STAGES / PIDNET_STEM2 (A/B measurements). */
int pidnet_set_option(pidnet_engine* h, const char* name, int value);

/* Debug: copy a named intermediate (e.g. "layer3", "pag3", "spp"; see engine.cu) to the host as fp32
 * NCHW.  `shape4` receives N,C,H,W; `host_out` may be NULL to query the shape only. */
int pidnet_debug_tensor(pidnet_engine* h, const char* name, float* host_out, int64_t* shape4);

/* ---- single-op entry points (kernel-level parity tests; device pointers, NHWC bf16 unless noted) ---- */

/* nn.Conv2d (+folded affine, +residual, +ReLU) on one NHWC bf16 tensor.  w: host fp32 [Cout][Cin/groups][k][k]
 * (already BN-folded), bias: host fp32 [Cout] or NULL, res: device NHWC bf16 [N,Ho,Wo,Cout] or NULL.
 * out_nhwc (bf16) or out_nchw_f32 (exactly one non-NULL).  k in {1,3}, pad = k/2, stride in {1,2}.
 * impl: 0 tcgen05 (weight-stationary halo kernel -- CTA pairs for Cin >= 128 -- where it applies, else the generic one),
 * 1 SIMT restatement, 2 generic tcgen05 kernel only, 3 weight-stationary kernels without CTA pairs and with two staging
 * buffers, 4 without CTA pairs. */
int pidnet_op_conv2d(void* stream, const void* x_nhwc, int N, int H, int W, int Cin, const float* w, const float* bias,
                     int Cout, int k, int stride, int groups, const void* res, int relu, void* out_nhwc,
                     float* out_nchw_f32, int impl);
/* conv1.0 + BN + ReLU of the stem: fp32 NCHW image -> bf16 NHWC (pidnet.py:25-27). w: host [Cout][3][3][3]. */
int pidnet_op_stem(void* stream, const float* x_nchw, int N, int H, int W, const float* w, const float* bias, int Cout,
                   void* out_nhwc);
/* PagFM fuse: x [N,H,W,C], low [N,h,w,2C+8] = [y|z|t|pad] -> out [N,H,W,C]. */
int pidnet_op_pag(void* stream, const void* x, const void* low, void* out, int N, int H, int W, int C, int h, int w,
                  int relu);
/* out = act(s*(a + bilinear(b)) + t); a/b/s/t optional (NULL). b is [N,h,w,C]. s,t device fp32 [C]. */
int pidnet_op_upadd(void* stream, const void* a, const void* b, void* out, int N, int H, int W, int C, int h, int w,
                    const float* s, const float* t, int relu);
/* AvgPool2d(k,stride,pad,count_include_pad) (k==0: global) + affine + ReLU. */
int pidnet_op_pool(void* stream, const void* x, void* out, int N, int H, int W, int C, int k, int stride, int pad,
                   const float* s, const float* t, int relu);
/* Light_Bag operand producer: out [N,H,W,2C]. */
int pidnet_op_lightbag(void* stream, const void* p, const void* i_low, const void* d, void* out, int N, int H, int W,
                       int C, int h, int w);
/* Bag blend + BN + ReLU: out [N,H,W,C]. */
int pidnet_op_bag(void* stream, const void* p, const void* i_low, const void* d, void* out, int N, int H, int W, int C,
                  int h, int w, const float* s, const float* t);

/* ---- criterion: replaces FullModel.forward after `outputs = self.model(inputs)` (utils/utils.py:41-57) with
 * OhemCrossEntropy (utils/criterion.py:43-99; thres / min_kept / class weights / ignore label) and BondaryLoss
 * (utils/criterion.py:102-132).  The global yacs values the reference reads (LOSS.BALANCE_WEIGHTS, LOSS.SB_WEIGHTS,
 * MODEL.ALIGN_CORNERS=True, TRAIN.IGNORE_LABEL) are passed explicitly.
 *   x_p, x_m, x_d : device fp32 NCHW low-res logits [N,C,h,w], [N,C,h,w], [N,1,h,w] (the three PIDNet outputs)
 *   labels        : device int64 [N,H,W];  bd_gt: device fp32 [N,H,W] (0/1)
 *   out16 (device fp32[16]): loss (== losses.mean()), loss_s (== loss_list[0].mean()), loss_b, pixel acc (over label >= 0
 *                   pixels, utils/utils.py:29-35), ohem(x_,labels), ohem(x_,bd_label), the two OHEM thresholds, the two valid
 *                   counts, |K1|, |K2|, [12] = number of labels that are neither ignore_label nor in [0, C) (the reference's
 *                   gather faults on them: such pixels are dropped, never used as an index, and the loss is NaN),
 *                   [13] = number of labels >= 0, [14..15] reserved.  An empty OHEM set (the reference raises IndexError,
 *                   utils/criterion.py:73) gives a NaN loss and contributes no gradient.
 *   grad_*        : optional device fp32 buffers shaped like the logits; receive d(loss.mean())/d(logits)
 *   aux_ce_map    : optional device fp32 [N,H,W]: the per-pixel term of the reference's loss -- nn.CrossEntropyLoss(reduction=
 *                   'none') of x_extra_p (weighted, 0 at ignored pixels; utils/criterion.py:50-60,94).  The reference's
 *                   FullModel returns loss as a [1,N,H,W] MAP (that term times BALANCE_WEIGHTS[0] plus the scalar terms);
 *                   callers that need the map shape rebuild it from this (pidnet_b200.FullModel(loss_map=True))
 * No host synchronisation; `workspace` needs pidnet_criterion_workspace_bytes(N,H,W) device bytes. */
typedef struct pidnet_criterion_cfg {
  int64_t ignore_label;        /* TRAIN.IGNORE_LABEL (255) */
  int64_t ohem_keep;           /* LOSS.OHEMKEEP (131072) */
  float ohem_thres;            /* LOSS.OHEMTHRES (0.9) */
  float bd_threshold;          /* 0.8, utils/utils.py:52 */
  float balance_weight_aux;    /* LOSS.BALANCE_WEIGHTS[0] (0.4) */
  float balance_weight_main;   /* LOSS.BALANCE_WEIGHTS[1] (1.0) */
  float sb_weight;             /* LOSS.SB_WEIGHTS (1.0) */
  float coeff_bce;             /* BondaryLoss(coeff_bce=20.0) */
} pidnet_criterion_cfg;
size_t pidnet_criterion_workspace_bytes(int N, int H, int W);
int pidnet_criterion(void* stream, const float* x_p, const float* x_m, const float* x_d, int N, int C, int h, int w,
                     const int64_t* labels, const float* bd_gt, int H, int W, const float* class_weights,
                     const pidnet_criterion_cfg* cfg, void* workspace, size_t workspace_bytes, float* out16,
                     float* grad_p, float* grad_m, float* grad_d, float* aux_ce_map);
/* F.interpolate(x, size=(H,W), mode='bilinear', align_corners=True) on fp32 NCHW (utils/utils.py:44-46) */
int pidnet_upsample_align_corners(void* stream, const float* x, int NC, int h, int w, float* out, int H, int W);

/* ---- training step: replaces, for one per-GPU batch shard, the reference's
 *   losses, _, acc, loss_list = model(images, labels, bd_gts); loss = losses.mean(); loss.backward()
 * (utils/function.py:43-48 with FullModel utils/utils.py:37-57 in train mode): train-mode forward (BatchNorm batch
 * statistics + running-stat update, momentum 0.1), fused criterion, and the backward pass of every op into the
 * caller's fp32 gradient buffers.  Parameters are BOUND, not copied: `dev_param` / `dev_grad` point at the storage of
 * the torch parameters (and of the running_mean / running_var buffers, dev_grad = NULL), keys are the reference
 * state_dict keys.  Gradients of different ranks are summed by the caller (NCCL all-reduce of the flat buffer).
 * S, M and L topologies, augment = 1. */
typedef struct pidnet_trainer pidnet_trainer;
int pidnet_train_create(const pidnet_cfg* cfg, pidnet_trainer** out);
int pidnet_train_destroy(pidnet_trainer* h);
int pidnet_train_bind(pidnet_trainer* h, const char* key, float* dev_param, float* dev_grad, const int64_t* shape, int ndim);
int pidnet_train_plan(pidnet_trainer* h, int N, int H, int W, size_t* arena_bytes);
/* out16: device fp32[16] as in pidnet_criterion; out_main/out_p/out_d: optional device copies of the low-res logits.
 * backward: 1 = forward + criterion + backward of everything in this call; 0 = forward + criterion values only;
 *           2 = forward + criterion values AND logit gradients, the network backward is left to pidnet_train_backward
 *               (this is how `loss.backward()` of the reference loop, utils/function.py:47, triggers it). */
int pidnet_train_step(pidnet_trainer* h, void* stream, const float* x_nchw, const int64_t* labels, const float* bd_gt,
                      const float* class_weights, const pidnet_criterion_cfg* cfg, int backward, float* out16,
                      float* out_main, float* out_p, float* out_d, float* aux_ce_map /* optional, as in pidnet_criterion */);
/* Backward of the last train-mode forward of this handle (pidnet_train_forward, or pidnet_train_step with backward 0 / 2):
 * what autograd runs for `outputs = self.model(inputs)` (utils/utils.py:39) when a loss built on the three outputs calls
 * .backward().  g_main / g_p / g_d: device fp32 gradients w.r.t. x_ / x_extra_p / x_extra_d shaped like the logits, or all
 * NULL to use the gradients the engine's own criterion left behind (backward = 2).  x_nchw: the image of that forward (the
 * stem's weight gradient reads it).  segment: -1 = the whole backward; k in [0, pidnet_train_num_segments) = the k-th of
 * the consecutive op ranges the backward is split into -- call them in order; after range k the flat-gradient ranges
 * reported by pidnet_train_segment_ranges(k) are final, so the caller can overlap their all-reduce with range k+1
 * (SURVEY 8e "bucketed in reverse layer order"; replaces DataParallel's reduce, tools/train.py:136). */
int pidnet_train_backward(pidnet_trainer* h, void* stream, const float* x_nchw, const float* g_main, const float* g_p,
                          const float* g_d, int segment);
int pidnet_train_num_segments(pidnet_trainer* h);
/* begin_end: int64 pairs [begin, end) in floats relative to grad_base (the start of the caller's flat gradient buffer);
 * begin_end may be NULL to query n_pairs only. */
int pidnet_train_segment_ranges(pidnet_trainer* h, int segment, const float* grad_base, int64_t* begin_end, int cap_pairs,
                                int* n_pairs);
/* PIDNet.forward in train mode (models/pidnet.py:136-182 with nn.BatchNorm2d in training mode): batch statistics, running
 * statistics updated, the three low-res outputs copied to the optional device buffers; no criterion, no backward */
int pidnet_train_forward(pidnet_trainer* h, void* stream, const float* x_nchw, float* out_main, float* out_p, float* out_d);
int pidnet_train_num_launches(pidnet_trainer* h, int* fwd, int* bwd);
/* options: "use_graph" 1 (default: the step replays two CUDA graphs after one eager step) | 0 (eager launches);
 *          "overlap_wgrad" 1 (default: weight-gradient GEMMs run on a side stream next to the dgrad chain) | 0;
 *          "fused_bn" 1 (default: single-launch BatchNorm kernels with a grid barrier) | 0 (3-kernel form; re-plan);
 *          "conv_stats" bit mask (default 7): the forward batch statistics of a conv -> BatchNorm pair are accumulated by the conv
 *                       kernel itself (bit 0: 3x3 weight-stationary kernel, bit 1: 1x1 weight-stationary kernel, bit 2: stem kernel) and
 *                       the BatchNorm launch only applies them | 0 (every BatchNorm makes its own statistics pass); re-plan;
 *          "wgrad_halo" 1 (default: halo-patch weight-gradient kernel for 3x3 stride-1 convs) | 0 (tap-by-tap; re-plan);
 *          "wgrad_stack" 1 (default: for Cout <= 64 the halo kernel's stacked-tap form, two filter taps per MMA) | 0 (re-plan);
 *          "bwd_segments" 4 (default) .. 16: number of ranges pidnet_train_backward splits the backward into */
int pidnet_train_set_option(pidnet_trainer* h, const char* name, int value);
int pidnet_train_profile(pidnet_trainer* h, void* stream, const float* x_nchw, const int64_t* labels, const float* bd_gt,
                         const float* class_weights, const pidnet_criterion_cfg* cfg, char* buf, size_t cap, float* crit_ms);
int pidnet_train_debug_tensor(pidnet_trainer* h, const char* name, int grad, float* host_out, int64_t* shape4);

/* ---- post-processing.
 * Fused bilinear upsample (align_corners=True) + argmax of the [N,C,h,w] fp32 logits to label size:
 *   pred      (optional) uint8 [N,H,W]: torch.argmax(F.interpolate(pred, size, 'bilinear', align_corners=True), 1)
 *             -- datasets/base_dataset.py:136-150 (`.exp()` is monotonic), tools/custom.py:90-92;
 *   confusion (optional) uint64 [C*C] += histogram of (label, prediction) over labels != ignore_label
 *             -- get_confusion_matrix, utils/utils.py:129-152 (row = ground truth, column = prediction).
 * The [N,C,H,W] tensor is never materialised.  All pointers are device pointers.
 * cell_mask_ws (optional): scratch of N*h*w uint32; when given, a first kernel marks per low-res cell the classes that can
 * win anywhere inside it (all others are dominated at the four corners) and the per-pixel kernel interpolates only those --
 * identical output, several times fewer interpolations on real logits. */
int pidnet_postprocess(void* stream, const float* logits, int N, int C, int h, int w, int H, int W, unsigned char* pred,
                       const int64_t* labels, int64_t ignore_label, unsigned long long* confusion, unsigned* cell_mask_ws);

/* ---- optimizer step on flat buffers.
 * Replaces torch.optim.SGD.step() as configured in tools/train.py:139-148 (momentum, weight decay, optional Nesterov) for ALL
 * parameters in one launch:  d = grad_scale * g + wd * p;  buf = first_step ? d : momentum * buf + (1 - dampening) * d;
 * d = nesterov ? d + momentum * buf : buf;  p -= lr * d.   n: floats (multiple of 4), pointers 16-byte aligned device fp32.
 * The poly learning-rate schedule of utils/utils.py:154-160 is host arithmetic (pidnet_b200/optim.py:adjust_learning_rate). */
int pidnet_sgd_step(void* stream, float* param, const float* grad, float* momentum_buf, int64_t n, float lr, float momentum,
                    float dampening, float weight_decay, int nesterov, int first_step, float grad_scale);

#ifdef __cplusplus
}
#endif
#endif /* PIDNET_B200_H_ */
