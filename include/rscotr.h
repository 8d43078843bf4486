/*
 * rscotr.h -- C ABI of librscotr_b200.so: the sm_100a kernels behind the RSCoTr
 * co-training hot path.
 *
 * Conventions (all entry points):
 *   - plain C: raw DEVICE pointers, explicit int sizes, a dtype enum, a CUDA
 *     stream passed as void* (cudaStream_t); no torch / C++ types.
 *   - the caller owns every buffer (inputs, outputs, workspaces); the library
 *     never allocates and keeps no global state besides a thread-local error
 *     string.  Calls are asynchronous on `stream`.
 *   - return 0 on success; RSC_ERR_* otherwise, message via rsc_last_error().
 *   - there is NO CPU path: a call on a machine without a CUDA device fails
 *     with RSC_ERR_CUDA.
 *
 * Each function cites the reference interface it stands in for.  The reference
 * (Li-Qingyun/RSCoTr) is pure Python over mmcv-full 1.6.1 / mmdet 2.25.1; the
 * only FFI on its path is mmcv's ext_module (ms_deform_attn_*,
 * sigmoid_focal_loss_*); the Swin ops are eager ATen chains inside
 * mmdet/models/backbones/swin.py which the fused kernels below replace.
 */
#ifndef RSCOTR_H_
#define RSCOTR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSC_OK 0
#define RSC_ERR_INVALID 1 /* bad argument (shape, dtype, alignment, null pointer) */
#define RSC_ERR_CUDA 2    /* launch / runtime error reported by CUDA */

/* element type of activation tensors ("x" pointers).  Parameters that the
 * reference keeps in fp32 (sampling locations, attention weights, LayerNorm
 * statistics, gradient accumulators) are always float. */
#define RSC_F32 0
#define RSC_BF16 1

const char *rsc_last_error(void);
int rsc_version(void);
/* number of kernels launched by this library in the calling process since the
 * last rsc_reset_launch_count() (bench.py reports it as gpu_launches). */
int64_t rsc_launch_count(void);
void rsc_reset_launch_count(void);

/* ------------------------------------------------------------------------
 * Window index maps.  Replaces: mmdet ShiftWindowMSA.forward's
 *   F.pad -> torch.roll(-shift) -> window_partition   (and the inverse
 *   window_reverse -> roll(+shift) -> crop), mmdet/models/backbones/swin.py,
 * reached from models/multi/multitask_learner.py:83 (self.backbone(img)).
 * SURVEY 8a rows a3/a5.  Integer index arithmetic: bit exact.
 * ---------------------------------------------------------------------- */
/* idx_out[B*nW*ws*ws] (int64): flat source token index into (B,H,W) for every
 * slot of the padded+shifted+partitioned tensor, -1 for a padded slot. */
int rsc_window_index_partition(int64_t *idx_out, int B, int H, int W, int ws, int shift, void *stream);
/* idx_out[B*H*W] (int64): for every token of (B,H,W) the flat slot in the
 * (B*nW, ws, ws) window tensor it is read back from by
 * window_reverse -> roll(+shift) -> crop. */
int rsc_window_index_reverse(int64_t *idx_out, int B, int H, int W, int ws, int shift, void *stream);
/* x (B,H,W,C) -> windows (B*nW, ws*ws, C), zero fill for padded slots. */
int rsc_window_partition(const void *x, void *windows, int B, int H, int W, int C, int ws, int shift, int dtype,
                         void *stream);
/* windows (B*nW, ws*ws, C) -> x (B,H,W,C) (crop fused). */
int rsc_window_reverse(const void *windows, void *x, int B, int H, int W, int C, int ws, int shift, int dtype,
                       void *stream);

/* ------------------------------------------------------------------------
 * Fused (shifted-)window attention core.  Replaces, in one kernel, the chain
 *   pad, roll, window_partition, reshape/permute qkv, q*scale, q@k^T,
 *   + relative_position_bias_table[index], + shift mask (0/-100), softmax,
 *   attn@v, transpose/reshape, window_reverse, roll back, crop
 * of mmdet ShiftWindowMSA.forward / WindowMSA.forward (SURVEY 8a rows a3-a5;
 * restated in oracle/swin.py::shift_window_msa).  The qkv and proj Linear
 * layers stay outside (GEMMs).
 *
 *   qkv       (B,H,W,3*C)  un-padded tokens; layout per token [q|k|v] x [head][32]
 *   qkv_bias  (3*C) float or NULL: value of a zero-padded token's q|k|v row
 *             (the reference pads AFTER norm1, so a padded row is Linear(0) = bias)
 *   bias_table((2*ws-1)^2, heads) float   -- relative_position_bias_table
 *   out       (B,H,W,C)
 * head_dim = C/heads must be 32, ws must be 7, shift in {0, 3}.
 * ---------------------------------------------------------------------- */
int rsc_wmsa_fwd(const void *qkv, const float *qkv_bias, const float *bias_table, void *out, int B, int H, int W,
                 int C, int heads, int ws, int shift, float scale, int dtype, void *stream);
/* Backward.  dqkv (B,H,W,3*C) is fully written.  dbias_table ((2*ws-1)^2,heads)
 * and dqkv_bias (3*C) (gradient reaching the qkv bias through padded rows;
 * may be NULL iff qkv_bias is NULL) are float and ACCUMULATED into (caller
 * zero-fills or carries a running gradient). */
int rsc_wmsa_bwd(const void *qkv, const float *qkv_bias, const float *bias_table, const void *dout, void *dqkv,
                 float *dbias_table, float *dqkv_bias, int B, int H, int W, int C, int heads, int ws, int shift,
                 float scale, int dtype, void *stream);

/* ------------------------------------------------------------------------
 * PatchMerging gather + LayerNorm.  Replaces mmdet PatchMerging.forward up to
 * (not including) the reduction Linear: NLC->NCHW, corner pad to even,
 * nn.Unfold(2, stride 2) (merged channel = c*4 + kh*2 + kw), LayerNorm(4C).
 * SURVEY 8a row a6; oracle/swin.py::patch_merging.
 *   x (B,H,W,C) -> y (B, ceil(H/2)*ceil(W/2), 4C); mean/rstd (B*L_out) float
 * ---------------------------------------------------------------------- */
int rsc_patch_merge_ln_fwd(const void *x, const float *gamma, const float *beta, void *y, float *mean, float *rstd,
                           int B, int H, int W, int C, float eps, int dtype, void *stream);
/* dx (B,H,W,C) fully written; dgamma/dbeta (4C) float ACCUMULATED. */
int rsc_patch_merge_ln_bwd(const void *x, const float *gamma, const float *mean, const float *rstd, const void *dy,
                           void *dx, float *dgamma, float *dbeta, int B, int H, int W, int C, int dtype,
                           void *stream);
/* Which kernels serve rsc_patch_merge_ln_*: 1 = one token per warp step (default), 2 = several tokens in flight per warp,
 * operands kept in registers (faster, opt-in: see DESIGN.md section 9), 0 = as the environment says
 * (RSC_PATCH_MERGE_V2=1 selects 2).  Results agree to fp32 rounding. */
int rsc_set_patch_merge_variant(int variant);

/* ------------------------------------------------------------------------
 * LayerNorm over the last dimension.  Replaces ATen native_layer_norm(+backward)
 * behind nn.LayerNorm in mmdet SwinBlock (norm1/norm2), SwinTransformer.norm{i},
 * PatchEmbed.norm and mmcv BaseTransformerLayer.norms.
 *   x (rows,C) in_dtype -> y (rows,C) out_dtype; mean/rstd (rows) float.
 * bwd: dy is out_dtype, dx is in_dtype (fully written); dgamma/dbeta (C) float ACCUMULATED.
 * C % 4 == 0, C <= 1024.
 * ---------------------------------------------------------------------- */
int rsc_layernorm_fwd(const void *x, const float *gamma, const float *beta, void *y, float *mean, float *rstd,
                      int64_t rows, int C, float eps, int in_dtype, int out_dtype, void *stream);
int rsc_layernorm_bwd(const void *x, const float *gamma, const float *mean, const float *rstd, const void *dy,
                      void *dx, float *dgamma, float *dbeta, int64_t rows, int C, int in_dtype, int out_dtype,
                      void *stream);

/* ------------------------------------------------------------------------
 * Multi-scale deformable attention.  Drop-in for mmcv-full 1.6.1
 *   ext_module.ms_deform_attn_forward / ms_deform_attn_backward
 * (mmcv/ops/multi_scale_deform_attn.py::MultiScaleDeformableAttnFunction),
 * reached from models/multi/bbox_head/transformer.py:211,258,
 * models/multi/seg_head/pixel_decoder.py:134, cls_head/pixel_decoder.py:95.
 * SURVEY 8a row a11.
 *   value   (B,Nv,heads,32)          dtype
 *   spatial_shapes (L,2) int64 (h,w); level_start_index (L) int64   [device]
 *   sampling_loc (B,Nq,heads,L,P,2) float in [0,1] (x,y)
 *   attn_weight  (B,Nq,heads,L,P)   float
 *   out     (B,Nq,heads*32)          dtype
 * heads*L*P... constraints: head_dim == 32, heads % 4 == 0, L*P == 16 is the
 * tuned case; any L*P that is a multiple of 2 and <= 64 is accepted.
 * im2col_step is accepted and ignored (the whole batch is one launch).
 * ---------------------------------------------------------------------- */
int rsc_msda_fwd(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                 const float *sampling_loc, const float *attn_weight, void *out, int B, int Nv, int Nq, int heads,
                 int L, int P, int im2col_step, int dtype, void *stream);
/* grad_value (B,Nv,heads,32) is FLOAT regardless of dtype and ACCUMULATED with
 * atomics (caller zero-fills, as mmcv does); grad_loc / grad_weight are float
 * and fully written. */
int rsc_msda_bwd(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                 const float *sampling_loc, const float *attn_weight, const void *grad_out, float *grad_value,
                 float *grad_loc, float *grad_weight, int B, int Nv, int Nq, int heads, int L, int P,
                 int im2col_step, int dtype, void *stream);

/* Fused tail of mmcv MultiScaleDeformableAttention.forward (ops/multi_scale_deform_attn.py): softmax over the
 * L*P attention logits + sampling_locations = reference_points + offsets / (W_l, H_l) (2-d references) or
 * + offsets / P * ref_wh * 0.5 (4-d references) + the sampling op above, in one kernel each way.
 * offsets (B,Nq,heads,L,P,2) and logits (B,Nq,heads,L*P) are the raw outputs of the two Linears (`off_dtype`),
 * ref (B,Nq,L,ref_dim) fp32 with ref_dim 2 or 4 (no gradient is produced for it: the reference feeds detached /
 * constant reference points), L*P must be 16.  grad_value is fp32 and ACCUMULATED (zero it first).
 * off_row_stride / logit_row_stride: elements between consecutive queries' rows of offsets / logits (and of their
 * gradients); 0 = dense (heads*L*P*2 and heads*L*P).  With strides the two matrices can be column ranges of ONE
 * (B*Nq, heads*L*P*3) matrix -- the output of a single GEMM over the stacked sampling_offsets / attention_weights
 * weights -- and the two gradients column ranges of the matrix that GEMM's backward reads. */
int rsc_msda_fused_fwd(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                       const void *offsets, const void *logits, const float *ref, void *out, int B, int Nv, int Nq,
                       int heads, int L, int P, int ref_dim, int dtype, int off_dtype, int off_row_stride,
                       int logit_row_stride, void *stream);
int rsc_msda_fused_bwd(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                       const void *offsets, const void *logits, const float *ref, const void *grad_out,
                       float *grad_value, void *grad_offsets, void *grad_logits, int B, int Nv, int Nq, int heads,
                       int L, int P, int ref_dim, int dtype, int off_dtype, int off_row_stride, int logit_row_stride,
                       void *stream);

/* ------------------------------------------------------------------------
 * Global average pool.  Replaces mmcls GlobalAveragePooling
 * (AdaptiveAvgPool2d((1,1))) called from
 * models/multi/cls_head/slvl_cls_head.py:14-18.  SURVEY 8a row a12.
 * channels_last = 0: x (B,C,HW) -> y (B,C);  1: x (B,HW,C) -> y (B,C).
 * ---------------------------------------------------------------------- */
int rsc_gap_fwd(const void *x, void *y, int B, int C, int HW, int channels_last, int dtype, void *stream);
int rsc_gap_bwd(const void *dy, void *dx, int B, int C, int HW, int channels_last, int dtype, void *stream);

/* ------------------------------------------------------------------------
 * Bilinear resize, align_corners=False, NCHW.  Replaces F.interpolate /
 * mmseg.ops.resize at models/multi/seg_head/mask2former_head.py:126-130 and
 * mmseg BaseDecodeHead.losses (mask2former_head.py:204).  SURVEY rows a18/a19.
 *   x (N,Hi,Wi) planes -> y (N,Ho,Wo) planes, N = B*C.
 * bwd: dx fully written (gather formulation, no atomics).
 * ---------------------------------------------------------------------- */
int rsc_bilinear_fwd(const void *x, void *y, int N, int Hi, int Wi, int Ho, int Wo, int dtype, void *stream);
int rsc_bilinear_bwd(const void *dy, void *dx, int N, int Hi, int Wi, int Ho, int Wo, int dtype, void *stream);
/* channels-last variants: x (B,Hi,Wi,C) -> y (B,Ho,Wo,C), C a multiple of the 16-byte channel vector (8 bf16 / 4 float);
 * same arithmetic; the maps that the convolution / norm kernels produce need no NCHW transpose around the resize
 * (mmseg `resize` in seg_head/pixel_decoder.py:55-64 and UPerHead's top-down path / fpn_outs resizes). */
int rsc_bilinear_cl_fwd(const void *x, void *y, int B, int C, int Hi, int Wi, int Ho, int Wo, int dtype, void *stream);
int rsc_bilinear_cl_bwd(const void *dy, void *dx, int B, int C, int Hi, int Wi, int Ho, int Wo, int dtype, void *stream);

/* ------------------------------------------------------------------------
 * Sigmoid focal loss.  Drop-in for mmcv ext_module.sigmoid_focal_loss_forward
 * / _backward (mmdet FocalLoss -> mmcv.ops.sigmoid_focal_loss), reached from
 * models/multi/bbox_head/mmdet_detr_head/detr_head.py:384 and dino_head.py:272.
 *   input (N,C) dtype logits; target (N) int64 in [0,C] (C = background);
 *   output / grad_input (N,C) float, element-wise (reduction by the caller).
 * ---------------------------------------------------------------------- */
int rsc_sigmoid_focal_loss_fwd(const void *input, const int64_t *target, float *output, int N, int C, float gamma,
                               float alpha, int dtype, void *stream);
int rsc_sigmoid_focal_loss_bwd(const void *input, const int64_t *target, float *grad_input, int N, int C,
                               float gamma, float alpha, int dtype, void *stream);

/* ------------------------------------------------------------------------
 * DINO / Deformable-DETR loss hot path.
 * Replaces, per training step, the 7 x B HungarianAssigner.assign calls of
 * models/multi/bbox_head/mmdet_detr_head/detr_head.py:475-543 (cost matrix with
 * mmdet FocalLossCost / BBoxL1Cost(xywh) / IoUCost(giou) + scipy
 * linear_sum_assignment behind a device->host sync each) and the 13 loss_single /
 * loss_dn_single calls of detr_head.py:333-416 and dino_head.py:236-365.
 *
 * Query element (l, b, q) of a segment lives at row (l*B + b)*NqTot + q0 + q of
 * cls (rows x C logits, `dtype`) and box (rows x 4 fp32, cxcywh in [0,1]).
 *
 * rsc_det_match: P = L*B problems.  gt_labels[G] / gt_boxes[G,4] (xyxy pixels) are
 *   concatenated over the images, gt_start[B+1] their offsets, img_wh[B,2] = (w,h).
 *   cost: caller workspace of P*max_gt*Nq floats, receives cost[p][i][q];
 *   assign[P,Nq]: global gt index matched to query q or -1; gt_norm (optional, [G,4]):
 *   normalised cxcywh of the gt boxes.  Assignment = exact rectangular linear sum
 *   assignment (fp64 duals), i.e. what scipy returns whenever the optimum is unique.
 * rsc_det_loss_fwd: out[row,3] += (loss_cls, loss_bbox, loss_iou) of layer l, already
 *   multiplied by the loss weights and divided by cls_factor[l] / pos_factor[l] (device
 *   floats, >= 1); row = out_row0 + (last_first ? (l == L-1 ? 0 : l+1) : l).
 *   assign is [L,B,Nq] (assign_layer_stride = B*Nq) or [B,Nq] shared by all layers (0).
 * rsc_det_loss_bwd: dout[rows,3] -> dcls (same layout/dtype as cls), dbox (fp32); writes
 *   every element of the segment.
 * ---------------------------------------------------------------------- */
int rsc_det_match(const void *cls, const float *box, const int64_t *gt_labels, const float *gt_boxes,
                  const int *gt_start, const float *img_wh, int P, int B, int NqTot, int q0, int Nq, int C, int max_gt,
                  float w_cls, float w_reg, float w_iou, float alpha, float gamma, float eps, float *cost, int *assign,
                  float *gt_norm, int dtype, void *stream);
int rsc_det_loss_fwd(const void *cls, const float *box, const int *assign, const int64_t *gt_labels,
                     const float *gt_norm, const float *img_wh, const float *cls_factor, const float *pos_factor,
                     float *out, int L, int B, int NqTot, int q0, int Nq, int C, int assign_layer_stride, int out_row0,
                     int last_first, float gamma, float alpha, float w_cls, float w_l1, float w_iou, float eps, int dtype,
                     void *stream);
int rsc_det_loss_bwd(const void *cls, const float *box, const int *assign, const int64_t *gt_labels,
                     const float *gt_norm, const float *img_wh, const float *cls_factor, const float *pos_factor,
                     const float *dout, void *dcls, float *dbox, int L, int B, int NqTot, int q0, int Nq, int C,
                     int assign_layer_stride, int out_row0, int last_first, float gamma, float alpha, float w_cls,
                     float w_l1, float w_iou, float eps, int dtype, void *stream);

/* ------------------------------------------------------------------------
 * Fused bilinear up-sampling (align_corners=False) + softmax cross-entropy
 * (SURVEY 8a row a19, 8f rank 2).  Replaces mmseg 0.28 BaseDecodeHead.losses
 * (resize(seg_logit, label size) + CrossEntropyLoss(ignore_index) + accuracy) called at
 * models/multi/seg_head/mask2former_head.py:204; the up-sampled (B,C,H,W) logits are
 * never materialised.  logits (B,C,h,w) `dtype`, label (B,H,W) int64, H >= h, W >= w,
 * scale <= 9.5.  fwd: stats[3] += {sum over non-ignored pixels of CE, #pixels whose
 * argmax == label, #non-ignored pixels}; lse (B,H,W) fp32 receives logsumexp per pixel
 * (+inf on ignored pixels) for the backward.  bwd: dlogits = gscale[0] *
 * d(sum CE)/d(logits), gscale a DEVICE float (upstream gradient x loss weight / #pixels);
 * ws: float scratch of B*h*w*4*C elements (per-cell tap sums, combined without atomics).
 * ---------------------------------------------------------------------- */
int rsc_upsample_ce_fwd(const void *logits, const int64_t *label, float *lse, float *stats, int B, int C, int h, int w,
                        int H, int W, int ignore_index, int dtype, void *stream);
int rsc_upsample_ce_bwd(const void *logits, const int64_t *label, const float *lse, const float *gscale, void *dlogits,
                        float *ws, int B, int C, int h, int w, int H, int W, int dtype, void *stream);

/* ------------------------------------------------------------------------
 * Fused residual-stream passes of the Swin block (mmdet 2.25.1
 * SwinBlock.forward: x = x + DropPath(attn(LN(x))); x = x + DropPath(FFN(LN(x)))).
 * rsc_add_ln_fwd:  r = identity + (x + bias) * scale[row / rows_per_sample];
 *                  n = LayerNorm(r; gamma, beta, eps); mean / rstd saved per row.
 *   identity, x, r_out, n_out: (rows, C) `dtype`; bias (C) and scale (samples) are fp32
 *   and may be NULL; C in {96,192,384,768,128,256,512,1024} (rsc_add_ln_supported).
 * rsc_add_ln_bwd:  dr = dr_ext + LayerNormBackward(dn)  -> d_identity;
 *                  dx = dr * scale (written only when dx != NULL; dx == NULL means dx = dr);
 *                  dbias += column sums of dx; dgamma / dbeta accumulated (all fp32).
 * rsc_bias_act_fwd: y = act(h + bias); act 0 = GELU, exact erf form (torch.nn.GELU default; the
 *                   bf16 path evaluates erf to 1.5e-7), act 1 = ReLU (the mmcv FFN of the encoder /
 *                   decoder layers, cnn/bricks/transformer.py FFN), act 2 (bf16 only, opt-in) = GELU with the
 *                   normal CDF evaluated as a fitted logistic, |error| < 2.6e-5.
 * rsc_bias_act_bwd: dh = dy * act'(h + bias); dbias += column sums of dh.
 * ---------------------------------------------------------------------- */
int rsc_add_ln_supported(int C);
int rsc_add_ln_fwd(const void *identity, const void *x, const float *bias, const float *scale, const float *gamma,
                   const float *beta, void *r_out, void *n_out, float *mean, float *rstd, int64_t rows,
                   int64_t rows_per_sample, int C, float eps, int dtype, void *stream);
int rsc_add_ln_bwd(const void *r, const float *gamma, const float *mean, const float *rstd, const void *dn,
                   const void *dr_ext, const float *scale, void *d_identity, void *dx, float *dgamma, float *dbeta,
                   float *dbias, int64_t rows, int64_t rows_per_sample, int C, int dtype, void *stream);
int rsc_bias_act_fwd(const void *h, const float *bias, void *y, int64_t rows, int C, int act, int dtype, void *stream);
int rsc_bias_act_bwd(const void *h, const float *bias, const void *dy, void *dh, float *dbias, int64_t rows, int C,
                     int act, int dtype, void *stream);

/* ------------------------------------------------------------------------
 * Patch-embedding gather (mmdet PatchEmbed = Conv2d(3, 96, k4, s4) -> LayerNorm).
 * x (B,Cin,H,W) NCHW `in_dtype`, H % 4 == W % 4 == 0 -> y (B*(H/4)*(W/4), Cin*16) `out_dtype`,
 * column order (c, kh, kw) = the Conv2d weight layout, so the projection is y @ weight.view(out, Cin*16)^T.
 * ---------------------------------------------------------------------- */
int rsc_patchify4(const void *x, void *y, int B, int Cin, int H, int W, int in_dtype, int out_dtype, void *stream);

/* ------------------------------------------------------------------------
 * Flat fused AdamW (+ gradient-clip scale).  Replaces mmcv OptimizerHook's
 * clip_grad_norm_ scaling + torch.optim.AdamW.step over one contiguous fp32 range
 * (optimizer built by mtl/utils/optimizer.py:25-55).
 * lr / step / clip_coef are DEVICE scalars (float); step holds the 1-based step
 * count t; clip_coef may be NULL (= 1).  torch.optim.AdamW arithmetic.
 * param_bf16 (may be NULL): bf16 shadow of `param`, rewritten in the same pass.
 * ---------------------------------------------------------------------- */
int rsc_adamw_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, const float *lr,
                   float lr_mult, float beta1, float beta2, float eps, float weight_decay, const float *step,
                   const float *clip_coef, void *param_bf16, void *stream);

/* y[c] += sum_r x[r][c]  (rows,C) -> (C) float, ACCUMULATED: the bias gradient of a Linear
 * layer (replaces ATen's sum(0) reduce in AddmmBackward / nn.Linear backward). */
int rsc_colsum(const void *x, float *y, int64_t rows, int C, int dtype, void *stream);

/* ------------------------------------------------------------------------
 * Linear layers as tcgen05 GEMMs with the surrounding element-wise passes in the
 * epilogue.  Replaces, for the bf16 compute path, nn.Linear (ATen addmm -> cuBLAS)
 * + the activation kernel behind mmcv FFN / the Swin MLP / qkv / proj
 * (cfg MTL_slvlcls_...py:9-25, :34-50) and their
 * autograd backward (mm, mm, sum(0), GeluBackward / threshold_backward).
 * All matrices are row-major bf16 with leading dimensions in ELEMENTS (multiples
 * of 8), 16-byte aligned; accumulation is fp32; bias / db are float.
 *
 * rsc_linear_fwd:  Y (M,N) = act(X (M,K) W (N,K)^T + bias)
 *     act 0: none.  act 1: GELU (torch's erf form); the pre-activation
 *     H = X W^T + bias is ALSO written (bf16) to `h` for the backward.  act 2: ReLU.
 * rsc_linear_dx:   dX (M,K) = (dY (M,N) W (N,K)) * act'(aux)
 *     act 1: aux = the saved H; act 2: aux = the saved Y (sign only); aux (M,K)
 *     shares dX's leading dimension.  act 0: aux ignored.
 * rsc_linear_dw:   dW (N,K) += dY (M,N)^T X (M,K)   (float, ACCUMULATED),
 *                  db (N)   += column sums of dY     (float, ACCUMULATED; may be NULL)
 * ---------------------------------------------------------------------- */
int rsc_linear_fwd(const void *x, const void *w, const float *bias, void *y, void *h, int64_t M, int N, int K, int64_t ldx,
                   int64_t ldw, int64_t ldy, int act, void *stream);
/* (r, n) = (identity + (X W^T + bias) * scale[row / rows_per_sample], LayerNorm(r) * gamma + beta): the Linear, the
 * residual add with the DropPath scale of the row's sample, and the NEXT LayerNorm in one kernel (the pair
 * nn.Linear -> rsc_add_ln_fwd of a Swin block's attention / MLP tail and of a post-norm transformer layer, SURVEY 8a rows
 * a2 / a9).  N (the normalised width) must fit one tile: N % 32 == 0, N <= 256.  identity / r_out / n_out are (M,N)
 * contiguous bf16; scale (M / rows_per_sample) float or NULL; mean / rstd (M) float are written for rsc_add_ln_bwd. */
int rsc_linear_add_ln_fwd(const void *x, const void *w, const float *bias, const void *identity, const float *scale,
                          const float *gamma, const float *beta, void *r_out, void *n_out, float *mean, float *rstd,
                          int64_t M, int N, int K, int64_t ldx, int64_t ldw, int64_t rows_per_sample, float eps,
                          void *stream);
int rsc_linear_dx(const void *dy, const void *w, const void *aux, void *dx, int64_t M, int N, int K, int64_t lddy,
                  int64_t ldw, int64_t lddx, int act, void *stream);
int rsc_linear_dw(const void *dy, const void *x, float *dw, float *db, int64_t M, int N, int K, int64_t lddy, int64_t ldx,
                  int64_t lddw, void *stream);
/* SMs the persistent rsc_linear_* kernels size their single wave for when the GEMM has fewer than `below_rows` token rows
 * (0 = every GEMM); sms = 0 restores all 148.  Used by the data-parallel step engine to leave room for the gradient
 * exchange that runs next to the small det / seg GEMMs. */
int rsc_set_gemm_sms(int sms, int64_t below_rows);

/* ------------------------------------------------------------------------
 * Convolutions as im2col GEMMs and the PPM pooling, channels-last maps (B,H,W,C).
 * Replaces nn.Conv2d (cuDNN) behind mmdet ChannelMapper (cfg MTL_slvlcls_...py:26-33),
 * the lateral / output / mask-feature convs of seg_head/pixel_decoder.py:39-64,158-170 and
 * mmseg UPerHead / PPM, and nn.AdaptiveAvgPool2d of PPM.
 * A k x k convolution is rsc_im2col_fwd + rsc_linear_fwd on the gathered matrix with the
 * weight laid out (Cout, kh, kw, Cin); a 1x1 convolution is rsc_linear_fwd alone on the
 * (B*H*W, C) token matrix.  Backward: rsc_linear_dx / rsc_linear_dw, then rsc_im2col_bwd.
 *   col  (B*Ho*Wo, kh*kw*C), column = (tap = i*kw + j, c); Ho = (H + 2 pad - kh)/stride + 1
 *   rsc_im2col_bwd: dx (B,H,W,C) = adjoint gather, fully written (no atomics)
 *   rsc_adaptive_avgpool_*: bins [floor(i H / S), ceil((i+1) H / S)) as torch
 * C % 8 == 0 (bf16) / C % 4 == 0 (float); 16-byte aligned pointers.
 * ---------------------------------------------------------------------- */
int rsc_im2col_fwd(const void *x, void *col, int B, int H, int W, int C, int kh, int kw, int stride, int pad, int dtype,
                   void *stream);
int rsc_im2col_bwd(const void *dcol, void *dx, int B, int H, int W, int C, int kh, int kw, int stride, int pad, int dtype,
                   void *stream);
int rsc_adaptive_avgpool_fwd(const void *x, void *y, int B, int H, int W, int C, int S, int dtype, void *stream);
int rsc_adaptive_avgpool_bwd(const void *dy, void *dx, int B, int H, int W, int C, int S, int dtype, void *stream);

/* ------------------------------------------------------------------------
 * Decoder attention core: O = softmax(Q K^T * scale + mask) V for nn.MultiheadAttention
 * as mmcv's MultiheadAttention wrapper calls it from the DINO decoder's self-attention
 * (models/multi/bbox_head/dino_head.py -> DinoTransformerDecoder; constant denoising mask
 * of query_denoising.py:167-190) and from the Mask2Former-style seg decoder
 * (models/multi/seg_head/mask2former_head.py:174-197; cross-attention masked by the previous
 * layer's mask prediction, :111-139).  Replaces F.multi_head_attention_forward's bmm / softmax /
 * bmm (8f rank 2).  bf16, head_dim 32, fp32 accumulation.
 *   q (Lq,B,H,32) / k, v (Lk,B,H,32) / out (Lq,B,H,32): element strides *_sl (sequence) and
 *   *_sb (batch), heads contiguous (32 apart); all strides multiples of 8.
 *   mask_bits: NULL, or uint32 words [image (stride mask_sb words, 0 = shared)][Lq][ceil(Lk/32)];
 *   bit (k & 31) of word k >> 5 set = key k masked for that query.  A fully masked row yields 0.
 *   lse (B*H, Lq) float: log2-domain log-sum-exp (saved for the backward).
 *   nsplit > 1 splits the keys over CTAs (use rsc_attn_nsplit); ws then holds
 *   nsplit*B*H*Lq*34 floats.
 * rsc_attn_bwd: dq32 (Lq,B,H,32) float, fully written; dk / dv bf16 with their own strides;
 *   delta_ws (B*H*Lq) float scratch.  `out` / `dout` share the strides o_sl / o_sb.
 * rsc_m2f_mask_bits: mask_pred (rows = B*Q, Hi, Wi) logits -> bits (rows, ceil(Ho*Wo/32)):
 *   bilinear resize (align_corners=False) to the key grid (Ho,Wo), masked = sigmoid < 0.5,
 *   rows with every key masked are cleared (mask2former_head.py:134-139, :177-178).
 * rsc_pack_mask_bits: boolean mask (rows, Lk) bytes, non-zero = masked -> bits.
 * ---------------------------------------------------------------------- */
int rsc_attn_nsplit(int B, int H, int Lq, int Lk);
int rsc_attn_fwd(const void *q, const void *k, const void *v, const void *mask_bits, void *out, float *lse, float *ws, int B,
                 int H, int Lq, int Lk, int head_dim, int64_t q_sl, int64_t q_sb, int64_t k_sl, int64_t k_sb, int64_t v_sl,
                 int64_t v_sb, int64_t o_sl, int64_t o_sb, int64_t mask_sb, int nsplit, float scale, void *stream);
int rsc_attn_bwd(const void *q, const void *k, const void *v, const void *mask_bits, const void *out, const void *dout,
                 const float *lse, float *delta_ws, float *dq32, void *dk, void *dv, int B, int H, int Lq, int Lk, int head_dim,
                 int64_t q_sl, int64_t q_sb, int64_t k_sl, int64_t k_sb, int64_t v_sl, int64_t v_sb, int64_t o_sl, int64_t o_sb,
                 int64_t dk_sl, int64_t dk_sb, int64_t dv_sl, int64_t dv_sb, int64_t mask_sb, float scale, void *stream);
int rsc_m2f_mask_bits(const void *mask_pred, void *bits, int rows, int Hi, int Wi, int Ho, int Wo, int dtype, void *stream);
int rsc_pack_mask_bits(const void *mask_u8, void *bits, int64_t rows, int Lk, void *stream);

/* ------------------------------------------------------------------------
 * GroupNorm / BatchNorm (training statistics) on channels-last maps, optional fused ReLU.
 * Replaces nn.GroupNorm / nn.BatchNorm2d (+ nn.ReLU) of mmcv ConvModule behind mmdet
 * ChannelMapper (cfg MTL_slvlcls_...py:26-33, GN-32), seg_head/pixel_decoder.py:39-64
 * (GN-32, ReLU on the output convs) and mmseg UPerHead / FCNHead (BN + ReLU)
 * (SURVEY 8a rows a8, a17, a20).
 *   x, y, dy, dx: (R, P, C), C innermost.  A statistic group = (row r, Cg consecutive channels)
 *   over the P pixels.  GroupNorm(G) on (B,H,W,C): R = B, P = H*W, Cg = C/G.
 *   BatchNorm2d (training): R = 1, P = B*H*W, Cg = 1; run_mean / run_var (C) are then updated
 *   with `momentum` (unbiased variance), NULL otherwise.
 *   stats (R, C/Cg, 2) float: mean, rstd (written by fwd, read by bwd).
 *   ws: float scratch, R*C*2 (fwd) / R*C*2 + R*(C/Cg)*2 (bwd).  dgamma / dbeta (C) float, ACCUMULATED.
 *   relu: y = max(norm(x), 0); the backward masks dy with the recomputed sign.
 * rsc_norm_supported: C % (16 / element size) == 0, C <= 2048, 256 % (C / vector) == 0.
 * ---------------------------------------------------------------------- */
int rsc_norm_supported(int C, int dtype);
int rsc_groupnorm_fwd(const void *x, const float *gamma, const float *beta, void *y, float *stats, float *ws, int R, int P, int C,
                      int Cg, float eps, int relu, float *run_mean, float *run_var, float momentum, int dtype, void *stream);
int rsc_groupnorm_bwd(const void *x, const void *dy, const float *gamma, const float *beta, const float *stats, void *dx,
                      float *dgamma, float *dbeta, float *ws, int R, int P, int C, int Cg, int relu, int dtype, void *stream);

/* ------------------------------------------------------------------------
 * Device half of a deferred Normalize (mmcv / mmcls / mmdet / mmseg `Normalize` on the loader side,
 * configs/_base_/{cls,det,seg}/*.py img_norm_cfg): the loader ships uint8 batches (4x fewer H2D bytes),
 * this kernel produces out[b][c] = (img[b][flip ? C-1-c : c] - mean[c]) * inv_std[c] for pixels inside the
 * image's valid (h, w) = valid_hw[b] and 0 in the right / bottom padding (the reference pads AFTER normalising).
 * img (B,C,H,W) uint8, out (B,C,H,W) float / bf16, W % 4 == 0; valid_hw (B,2) int32 or NULL.
 * ---------------------------------------------------------------------- */
int rsc_normalize_u8(const void *img, void *out, const float *mean, const float *inv_std, const int *valid_hw, int B, int C, int H,
                     int W, int flip, int out_dtype, void *stream);

/* ------------------------------------------------------------------------
 * Gradient all-reduce (mean) inside the NVSwitch: replaces the ncclAllReduce behind the reference's DDP wrapper
 * (mtl/apis/train.py:37-46) when the flat gradient buffer lives in symmetric memory with an NVLS multicast
 * mapping.  mc = multicast address of element 0; [lo, hi) float elements, multiples of 4.  Rank `rank` of `world`
 * reduces its 1/world share: multimem.ld_reduce.add (sum over all GPUs, computed by the switch) -> x scale ->
 * multimem.st (to every GPU).  The caller brackets the call with cross-GPU barriers (all ranks have written the range /
 * all shares are stored).  ctas: CTAs of 512 threads (0 = 64).
 * ---------------------------------------------------------------------- */
int rsc_nvls_allreduce_mean(void *mc, int64_t lo, int64_t hi, int rank, int world, float scale, int ctas, void *stream);

/* ------------------------------------------------------------------------
 * Iterative box refinement: out = sigmoid(tmp + inverse_sigmoid(ref, eps)), the chain
 * `tmp + inverse_sigmoid(reference)` -> `.sigmoid()` of DinoTransformerDecoder.forward and DINOHead.forward
 * (models/multi/bbox_head/dino_head.py; mmdet inverse_sigmoid: clamp(ref,0,1), log(max(x,eps) / max(1-x,eps))).
 * tmp: `dtype` (the reg branch output), ref / out / dout / dref: float, n elements.  bwd: dtmp (`dtype`) =
 * dout * out * (1 - out); dref (may be NULL) = dtmp * d inverse_sigmoid / d ref with torch's clamp conventions.
 * ---------------------------------------------------------------------- */
int rsc_box_refine_fwd(const void *tmp, const float *ref, float *out, int64_t n, float eps, int dtype, void *stream);
int rsc_box_refine_bwd(const float *out, const float *ref, const float *dout, void *dtmp, float *dref, int64_t n, float eps,
                       int dtype, void *stream);

/* ------------------------------------------------------------------------
 * Backward of a SMALL Linear layer (y = x W^T + b with fewer than ~4096 token rows: decoder projections, reg / cls
 * branches, mask-embedding MLPs of dino_head.py / mask2former_head.py) in one launch instead of the library's three
 * (mm, addmm, column sum):  dx (M,K) = dy (M,N) w (N,K) [bf16];  dw (N,K) += dy^T x [float, ACCUMULATED];
 * db (N) += column sums of dy [float, ACCUMULATED].  All inputs bf16 row-major, leading dimensions in elements
 * (multiples of 8), 16-byte aligned; dx / dw / db may be NULL (db needs dw).
 * ---------------------------------------------------------------------- */
int rsc_small_linear_bwd(const void *dy, const void *x, const void *w, void *dx, float *dw, float *db, int M, int N, int K,
                         int64_t lddy, int64_t ldx, int64_t ldw, int64_t lddx, int64_t lddw, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RSCOTR_H_ */
