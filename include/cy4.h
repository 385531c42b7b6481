/*
 * cy4.h -- C-ABI of libcy4.so, the B200 (sm_100a) implementation of the Complex-YOLOv4 training
 * hot path.  Plain pointers and sizes only; no torch types.
 *
 * The reference (maudzung/Complex-YOLOv4-Pytorch) is pure Python and has no FFI of its own
 * (SURVEY.md F1); the boundary a maintainer binds is therefore the set of reference *functions*
 * each entry point replaces, cited per declaration as file:line under /root/reference/src.
 * INTEGRATION.md shows the ctypes stubs.
 *
 * Conventions
 *   - every pointer is a BORROWED device pointer (cudaMalloc'd / torch-owned) unless marked host;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - functions return 0 on success, <0 on error; cy4_last_error() returns a thread-local,
 *     NUL-terminated description of the most recent failure on the calling thread;
 *   - no function synchronises the device, allocates caller-visible memory or throws;
 *   - all entry points are re-entrant across threads.
 */
#ifndef CY4_H
#define CY4_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CY4_VERSION 100

#if defined(CY4_BUILD) && defined(__GNUC__)
#define CY4_API __attribute__((visibility("default")))
#else
#define CY4_API
#endif

/* flags for the rotated-IoU entry points */
#define CY4_F_GIOU 1u /* reference clipper + convex-hull GIoU term (GIoU=True); else exact-intersection IoU */

CY4_API int cy4_version(void);
CY4_API const char *cy4_last_error(void);
/* 0 if a CUDA device of compute capability 10.x is usable by this library, <0 otherwise (host call). */
CY4_API int cy4_device_ok(void);
/* number of CUDA kernels this library has launched in this process (reset != 0: return and clear) */
CY4_API long long cy4_kernel_launches(int reset);
/* Kernels launched by replaying a CUDA graph do not pass through the counting entry points: after each replay the caller adds
 * the number of this library's kernel launches that were captured into that graph (cy4_kernel_launches delta over the capture). */
CY4_API int cy4_note_graph_replay(int n_kernels);

/* ---- rotated-box geometry -------------------------------------------------------------------
 * utils/iou_rotated_boxes_utils.py:98-142  iou_pred_vs_target_boxes (element-wise pairs), with
 * utils/cal_intersection_rotated_boxes.py:42-96 (intersection_area / PolyArea2D) inside.
 * pred6/tgt6: [n,6] fp32 rows (x, y, w, l, im, re).  iou, term: [n] fp32; term is the summand of
 * giou_loss (1 - iou + (C-U)/C with GIoU, 1 - iou without).  gpred6 (optional, may be NULL):
 * [n,6] d term / d pred under the reference's autograd semantics, multiplied by
 * gterm[k] when gterm != NULL (else by 1). */
CY4_API int cy4_rgiou_pairs(const float *pred6, const float *tgt6, int64_t n, uint32_t flags,
                    float *iou, float *term, const float *gterm, float *gpred6, void *stream);

/* Sequential-order (reference `giou_loss += term`, :133) fp32 sum of term[0..n) into out[0]. */
CY4_API int cy4_sum_f32_seq(const float *term, int64_t n, float *out, void *stream);

/* utils/iou_rotated_boxes_utils.py:34-61 get_corners_vectorize: boxes5 [n,5] (x,y,w,l,yaw) -> [n,4,2]. */
CY4_API int cy4_corners(const float *x, const float *y, const float *w, const float *l, const float *yaw,
                int64_t n, float *corners /* [n,4,2] */, void *stream);

/* utils/cal_intersection_rotated_boxes.py:42-96: intersection_area of n quad pairs ([n,4,2] each). */
CY4_API int cy4_quad_intersection_area(const float *rect1, const float *rect2, int64_t n, float *area, void *stream);
/* utils/cal_intersection_rotated_boxes.py:93-96: PolyArea2D of one polygon pts [k,2], k <= 16. */
CY4_API int cy4_poly_area(const float *pts, int k, float *area, void *stream);

/* utils/iou_rotated_boxes_utils.py:64-95: get_polygons_areas_fix_xy + iou_rotated_boxes_targets_vs_anchors.
 * anchors4 [nA,4], tgt4 [nT,4] rows (w, l, im, re), both placed at (100,100); ious [nA,nT]. */
CY4_API int cy4_anchor_iou(const float *anchors4, int nA, const float *tgt4, int64_t nT, float *ious, void *stream);

/* ---- YOLO head --------------------------------------------------------------------------------
 * models/yolo_layer.py:144-253 YoloLayer.forward and :69-142 build_targets.
 * The raw head tensor is addressed through element strides so both the reference's NCHW layout and
 * the engine's NHWC layout work:  value(b, c, y, x) = pred[b*sB + c*sC + y*sH + x*sW].          */
typedef struct cy4_yolo_desc {
    int32_t B, G, nA, nC;         /* batch, grid size, anchors per cell, classes (nC <= 32) */
    int64_t sB, sC, sH, sW;       /* element strides of the raw head tensor */
    float img_size;               /* 608: stride = img_size / G (yolo_layer.py:56) */
    float ignore_thresh;          /* 0.7 (yolo_layer.py:119, strict >) */
    uint32_t use_giou;            /* 1: total = 3.54 giou + 3.54 eular + 64.3 obj + 37.4 cls (:213-215) */
    uint32_t reserved;
} cy4_yolo_desc;

#define CY4_YOLO_NMETRICS 18      /* order of yolo_layer.py:232-251 */

/* bytes of scratch cy4_yolo_loss_* need for (desc, nT); contents opaque, must persist fwd -> bwd */
CY4_API size_t cy4_yolo_workspace_bytes(const cy4_yolo_desc *d, int64_t nT);

/* Decode only (targets=None): output [B, nA*G*G, 7+nC] rows (x,y,w,l in pixels, im, re, conf, cls..),
 * row order anchor-major then gj then gi (yolo_layer.py:184-189). anchors4 [nA,4] = (w/stride, h/stride, im, re). */
CY4_API int cy4_yolo_decode(const cy4_yolo_desc *d, const float *pred, const float *anchors4, float *output, void *stream);

/* Decode + target assignment + losses + metrics in one pass over the head tensor.
 * targets8 [nT,8] rows (image, class, x, y, w, l, im, re) normalised to [0,1) (may be NULL iff nT==0).
 * loss[1]; metrics[18]; status[1] int32: bit0 = a target indexed outside the grid/batch
 * (the reference raises IndexError there; such targets are skipped). */
CY4_API int cy4_yolo_loss_fwd(const cy4_yolo_desc *d, const float *pred, const float *anchors4,
                      const float *targets8, int64_t nT, float *output /* may be NULL */,
                      float *loss, float *metrics, int32_t *status, void *workspace, void *stream);

/* d loss / d pred, scaled by gloss[0]; dpred uses dense strides (dsB, dsC, dsH, dsW) and is fully
 * overwritten for the nA*(7+nC) head channels. */
CY4_API int cy4_yolo_loss_bwd(const cy4_yolo_desc *d, const float *pred, const float *anchors4,
                      const float *targets8, int64_t nT, const float *gloss, const void *workspace,
                      float *dpred, int64_t dsB, int64_t dsC, int64_t dsH, int64_t dsW, void *stream);

/* models/yolo_layer.py:69-142 build_targets with its 13 dense outputs (API parity; the fused
 * loss above does not materialise them).  pred_boxes [B,nA,G,G,6], pred_cls [B,nA,G,G,nC] contiguous.
 * Dense outputs are [B,nA,G,G] (tcls [B,nA,G,G,nC]); masks are uint8 0/1; giou_loss[1] is already
 * divided by nT.  idx (optional) [5,nT] int64: b, best_n, gj, gi, label. */
CY4_API int cy4_build_targets(const cy4_yolo_desc *d, const float *pred_boxes, const float *pred_cls,
                      const float *targets8, int64_t nT, const float *anchors4,
                      float *iou_scores, float *giou_loss, float *class_mask, uint8_t *obj_mask,
                      uint8_t *noobj_mask, float *tx, float *ty, float *tw, float *th, float *tim,
                      float *tre, float *tcls, float *tconf, int64_t *idx, int32_t *status,
                      void *workspace, void *stream);


/* ---- convolution stack on the tensor cores (tcgen05) -----------------------------------------
 * Replaces nn.Conv2d inside models/darknet2pytorch.py:247-278 (create_network's conv blocks) and
 * its autograd backward.  Activations are NHWC fp16 (channel stride ld, so producers can write
 * into channel slices of a route/concat buffer); weights are K-major packed fp16:
 *   fprop  [Cout_pad][kh][kw][Cin]   (cy4_pack_weight_fprop, from the module's OIHW fp32 parameter)
 *   dgrad  [Cin_pad][kh][kw][Cout]   (cy4_pack_weight_dgrad)
 * with *_pad = channels rounded up to a multiple of 32 (zero rows).  Cin must be a multiple of 32
 * (the 3-channel stem goes through cy4_stem_im2col first, see below). */
typedef struct cy4_conv_desc {
    int32_t B, Hi, Wi, Cin;   /* input  [B,Hi,Wi,Cin], channel stride ldx (elements) */
    int32_t Ho, Wo, Cout;     /* output [B,Ho,Wo,Cout], channel stride ldy (elements) */
    int32_t ksize, stride, pad;
    int64_t ldx, ldy;
    uint32_t flags;           /* CY4_CONV_* */
    uint32_t reserved;
} cy4_conv_desc;

#define CY4_CONV_OUT_F32 1u   /* y is fp32 (+ bias when bias != NULL); default fp16 */
#define CY4_CONV_STATS   2u   /* also accumulate per-channel sum / sum of squares of the fp32 results
                                 into ch_sum / ch_sqsum (atomicAdd; caller zeroes them) -- the batch
                                 statistics of the BatchNorm2d that follows (darknet2pytorch.py:260) */
#define CY4_CONV_A_MATRIX 8u  /* x is a plain [B*Ho*Wo, Cin] matrix (1x1/s1/p0 only): tiled TMA instead of
                                 im2col TMA -- used for the stem's explicit im2col matrix */
#define CY4_CONV_ZERO_ACC 16u /* cy4_conv_wgrad: memset dw_acc before accumulating into it */
#define CY4_CONV_ACCUM   4u   /* y += result instead of y = result (fp16), for gradients of tensors
                                 with several consumers (route / shortcut, darknet2pytorch.py:180-219) */

/* Tunables: "conv_cluster" (1|2|4) and "wgrad_cluster" (1|2): CTAs per cluster sharing weight / activation
 * slabs through TMA multicast; "tma_store" (0|1): swizzled-smem + TMA-store epilogue of the fp16 outputs;
 * "kblocks_per_slot" (1..8, default 4): upper bound on the k-blocks of a narrow layer packed into one pipeline slot;
 * "conv1x1_matrix" (0|1): 1 = 1x1 / stride-1 convs read their activation through a plain 2-D tiled TMA instead of im2col mode;
 * "wgrad_wide32" (0|1, default 1): weight gradient of 32-channel inputs issues one N = 32*taps MMA per K step instead of one per tap;
 * "accum_tma" (0|1, default 1): CY4_CONV_ACCUM outputs through TMA reduce-add stores instead of per-thread read-modify-write;
 * "wgrad_pair" (0|1, default 1): CTA-pair weight-gradient kernel for layers with Cout % 256 == 0 and a 128/256-channel X tile;
 * "conv_pair" (0|1, default 1): CTA-pair (tcgen05 cta_group::2, 256-row tiles) kernel for the eligible conv launches;
 * "pdl" (0|1; environment CY4_PDL sets the initial value): programmatic dependent launch of the tensor-core and BN / activation
 *   kernels -- each calls griddepcontrol.launch_dependents first thing and griddepcontrol.wait before its first global access,
 *   so the next kernel's blocks are placed and set up (barriers, TMEM, descriptors) while the previous grid drains;
 * "ew_fwd_blocks_per_sm" (1..32, default 3) / "ew_bwd_blocks_per_sm" (default 2): grid caps of the BN / activation passes;
 * "ew_carveout" (0|1, default 0; read at a kernel's first launch): 1 = the BN / activation passes prefer the maximum shared-memory
 *   carve-out (experiment: measured slower, kept for the record);
 * "dgrad_interleave" (0|1, default 1): merged stride-2 input gradients whose dY exceeds 64 MB walk the four output-parity classes
 *   of a tile in neighbouring work units (three of the four dY reads hit L2) instead of class after class (four DRAM passes);
 * "slab_stats" (0|1, default 1): CY4_CONV_STATS sums are read off the staged fp16 output slab (the statistics of the STORED tensor,
 *   ~100 instructions per 32x32 slab) instead of a shuffle reduce-scatter over the fp32 accumulators (~400, a dependent chain);
 * "debug": bottleneck experiments, honoured only by -DCY4_PROBE side builds (tools/probe_pipeline.py). */
CY4_API int cy4_set_option(const char *name, int value);
/* y = conv(x, w) */
CY4_API int cy4_conv_fwd(const cy4_conv_desc *d, const void *x, const void *w_fprop, void *y, const float *bias,
                         float *ch_sum, float *ch_sqsum, void *stream);
/* cy4_conv_fwd with CY4_CONV_STATS, the statistics taken about a per-channel shift c[Cout] (NULL = 0): ch_sum += sum (y - c),
 * ch_sqsum += sum (y - c)^2.  With c close to the channel mean (the engine passes last step's batch mean) the batch variance
 * ch_sqsum/n - (ch_sum/n)^2 is free of the cancellation that E[y^2] - E[y]^2 suffers once |mean| >> sigma; the mean is
 * c + ch_sum/n (cy4_bn_train_act_fwd takes the same c). */
CY4_API int cy4_conv_fwd_stats(const cy4_conv_desc *d, const void *x, const void *w_fprop, void *y, float *ch_sum, float *ch_sqsum,
                               const float *stat_shift, void *stream);
/* Eval-mode inference: y = act(conv(x, w_folded) + shift[c]) (+ residual), fp16 NHWC.  BatchNorm
 * (running statistics) is folded into the packed weights (cy4_pack_item.fold_scale) and `shift` = beta - mean*scale;
 * Mish / LeakyReLU run in the conv epilogue on the fp32 accumulators; `residual` (NULL or fp16 [B*Ho*Wo, ldr]) is the
 * shortcut input of a fused [shortcut] block.  Replaces nn.Conv2d + nn.BatchNorm2d(eval) + Mish/LeakyReLU (+ shortcut)
 * of models/darknet2pytorch.py:247-278,208-219 as ONE kernel.  d->flags may only carry CY4_CONV_A_MATRIX. */
CY4_API int cy4_conv_fwd_fused(const cy4_conv_desc *d, const void *x, const void *w_fprop, void *y, const float *shift, int act,
                               const void *residual, int64_t ldr, void *stream);
/* dx (+)= conv_transpose(dy, w); d describes the FORWARD conv; dy [B,Ho,Wo,Cout] (ld = ldy),
 * dx [B,Hi,Wi,Cin] (ld = ldx).  stride 1 (any odd k, pad = k/2) and stride 2 (k=3, pad=1, even Hi/Wi). */
CY4_API int cy4_conv_dgrad(const cy4_conv_desc *d, const void *dy, const void *w_dgrad, void *dx, void *stream);
/* Input gradient fused with the first pass of the PRODUCER's BatchNorm/activation backward.  dx is the gradient of the
 * producer's activated output A = act(scale*Y + shift); instead of dA this call stores dz = dA_total * act'(scale*Y + shift)
 * (dA_total includes the previous contents of dx when CY4_CONV_ACCUM is set -- use it on the LAST writer of that
 * gradient) and accumulates sum_dz[c] += sum_m dz, sum_dzy[c] += sum_m dz*Y (fp32 atomics; caller zeroes them).
 * y_producer: the producer's raw conv output, fp16 [B*Hi*Wi, ldyp]; scale/shift: its BatchNorm scale/shift [Cin].
 * Equivalent to cy4_conv_dgrad followed by the reduce half of cy4_bn_act_bwd_reduce (sum_dzx = rstd*(sum_dzy - mean*sum_dz)). */
CY4_API int cy4_conv_dgrad_fused(const cy4_conv_desc *d, const void *dy, const void *w_dgrad, void *dx, const void *y_producer,
                                 int64_t ldyp, const float *scale, const float *shift, int act, float *sum_dz, float *sum_dzy,
                                 void *stream);
/* dw_acc[Cout_pad][kh*kw][Cin] (fp32, caller-zeroed) += dy^T * im2col(x): split-K partial sums are
 * added with atomics.  cy4_unpack_wgrad then writes the OIHW fp32 gradient of the parameter. */
CY4_API int cy4_conv_wgrad(const cy4_conv_desc *d, const void *x, const void *dy, float *dw_acc, void *stream);
/* Host-side tiling of that launch, without launching (tests / tooling): out8 = m_tiles, n_tiles, block_n, taps per CTA,
 * tap groups, 128-pixel k-blocks, split-K factor, grid size.  One 193 KB CTA fits per SM, so the grid runs in waves of
 * <SM count> CTAs; the split minimises waves x (k-blocks per CTA + fixed cost). */
CY4_API int cy4_conv_wgrad_plan(const cy4_conv_desc *d, int32_t *out8);

CY4_API int cy4_pack_weight_fprop(const float *w_oihw, int Cout, int Cin, int ksize, int cin_pad, void *w_packed, void *stream);
CY4_API int cy4_pack_weight_dgrad(const float *w_oihw, int Cout, int Cin, int ksize, void *w_packed, void *stream);
/* gw_oihw (+)= scale * (dscale ? *dscale : 1) * dw_acc, re-laid out to OIHW */
CY4_API int cy4_unpack_wgrad(const float *dw_acc, int Cout, int Cin, int ksize, int cin_pad, float scale, const float *dscale,
                             int accumulate, float *gw_oihw, void *stream);
/* One launch for all conv layers: device arrays of items (pointers are borrowed device pointers). */
typedef struct cy4_pack_item {
    const float *w_oihw;     /* [Cout][Cin][k][k] fp32 */
    void *w_fprop;           /* [cout_pad][k*k][Cin] fp16, or NULL */
    void *w_dgrad;           /* [cin_pad][k*k][cout_pad] fp16, or NULL */
    int32_t Cout, Cin, ksize, cout_pad, cin_pad;
    int32_t tile_begin;      /* index of this item's first 32 x 32 (co x ci) tile in the launch-wide tile list: the prefix sum of
                                ceil(cout_pad/32) * ceil(cin_pad/32) over the preceding items (0 for the first) */
    const float *fold_scale; /* NULL, or [Cout] fp32: w_fprop rows are multiplied by it (eval-mode BatchNorm folding,
                                gamma * rsqrt(running_var + eps); the matching shift goes to cy4_conv_fwd_fused) */
} cy4_pack_item;
typedef struct cy4_unpack_item {
    const float *dw_acc;     /* [Cout_pad][k*k][Cin] fp32 */
    float *gw_oihw;          /* [Cout][Cin][k][k] fp32 */
    int32_t Cout, Cin, ksize;
    int32_t tile_begin;      /* prefix sum of ceil(Cout/32) * ceil(Cin/32) over the preceding items */
} cy4_unpack_item;
CY4_API int cy4_pack_weights_batched(const cy4_pack_item *items_dev, int n, void *stream);
CY4_API int cy4_unpack_wgrad_batched(const cy4_unpack_item *items_dev, int n, const float *dscale, void *stream);
/* Stem: x NCHW fp32 [B,3,H,W] -> im2col matrix [B*Ho*Wo, 32] fp16 (27 taps*channels (r,s,c order) + 5 zeros) */
CY4_API int cy4_stem_im2col(const float *x_nchw, int B, int C, int H, int W, int ksize, int stride, int pad,
                            void *cols /* [B*Ho*Wo, 32] fp16 */, void *stream);

/* ---- bandwidth-bound layers on NHWC fp16 (channel stride ld, channels % 8 == 0) -------------------
 * Replace nn.BatchNorm2d / Mish / nn.LeakyReLU / torch.cat / shortcut add / nn.MaxPool2d /
 * Upsample_expand of models/darknet2pytorch.py:22-28,64-79,180-219,256-285 and their backward. */
#define CY4_ACT_LINEAR 0
#define CY4_ACT_LEAKY  1   /* LeakyReLU(0.1), darknet2pytorch.py:265-266 */
#define CY4_ACT_MISH   2   /* x * tanh(softplus(x)), darknet2pytorch.py:22-28 */

/* Batch statistics (sum / sum of squares from cy4_conv_fwd, count = B*H*W) -> per-channel
 * scale = gamma*rstd, shift = beta - mean*scale; training also updates running_mean/var (momentum,
 * unbiased variance) and num_batches_tracked exactly like nn.BatchNorm2d; training=0 uses them. */
CY4_API int cy4_bn_finalize(const float *ch_sum, const float *ch_sqsum, float count, const float *gamma, const float *beta,
                            float *running_mean, float *running_var, int64_t *num_batches_tracked, float momentum, float eps,
                            int training, int C, float *scale, float *shift, float *mean, float *rstd, void *stream);
/* cy4_bn_finalize(training=1) + cy4_bn_act_fwd in ONE launch: every thread derives scale / shift of its channels from the
 * batch sums; scale / shift / mean / rstd (for the backward pass), running_mean / running_var / num_batches_tracked are
 * written once, with nn.BatchNorm2d's update rule (darknet2pytorch.py:260, momentum, unbiased variance). */
CY4_API int cy4_bn_train_act_fwd(const void *y, int64_t ldy, const float *ch_sum, const float *ch_sqsum, float count, const float *gamma,
                                 const float *beta, float *running_mean, float *running_var, int64_t *num_batches_tracked,
                                 float momentum, float eps, float *scale, float *shift, float *mean, float *rstd, int act,
                                 const void *residual, int64_t ldr, void *out, int64_t ldo, int64_t M, int C,
                                 const float *stat_shift /* the c of cy4_conv_fwd_stats, or NULL */,
                                 float *shift_next /* NULL, or [C]: receives the batch mean (next step's c); must not alias stat_shift */,
                                 void *stream);
/* out = act(y*scale + shift) (+ residual) */
CY4_API int cy4_bn_act_fwd(const void *y, int64_t ldy, const float *scale, const float *shift, int act, const void *residual,
                           int64_t ldr, void *out, int64_t ldo, int64_t M, int C, void *stream);
/* sum_dz[c] += sum_m dz, sum_dzx[c] += sum_m dz*xhat (= d beta, d gamma), dz = dA*act'(z).  For a
 * non-linear activation dA is OVERWRITTEN with dz (fp16) so that the apply pass need not recompute act'. */
CY4_API int cy4_bn_act_bwd_reduce(const void *y, int64_t ldy, void *dA, int64_t ldg, const float *scale, const float *shift,
                                  const float *mean, const float *rstd, int act, int64_t M, int C, float *sum_dz, float *sum_dzx,
                                  void *stream);
/* After cy4_conv_dgrad_fused: converts its raw per-channel sums in place, sum_dzy <- rstd * (sum_dzy - mean * sum_dz)
 * (= sum dz*xhat = d gamma), which is what cy4_bn_act_bwd_reduce would have produced. */
CY4_API int cy4_bn_bwd_fixup(const float *sum_dz, float *sum_dzy_inout, const float *mean, const float *rstd, int C, void *stream);
/* dY = scale*(dz - sum_dz/M - xhat*sum_dzx/M) (training) or scale*dz (eval).  dz_ready != 0: dA already
 * holds dz (cy4_bn_act_bwd_reduce ran on it); otherwise dz = dA*act'(z) is recomputed here. */
CY4_API int cy4_bn_act_bwd_apply(const void *y, int64_t ldy, const void *dA, int64_t ldg, const float *scale, const float *shift,
                                 const float *mean, const float *rstd, const float *sum_dz, const float *sum_dzx, float inv_count,
                                 int training, int act, int dz_ready, void *dY, int64_t ldd, int64_t M, int C, void *stream);
/* out = a (+ b): route copies into / out of concat buffers, shortcut add, gradient accumulation */
CY4_API int cy4_add_copy(const void *a, int64_t lda, const void *b, int64_t ldb, void *out, int64_t ldo, int64_t M, int C, void *stream);
CY4_API int cy4_upsample2x_fwd(const void *in, int64_t ldi, void *out, int64_t ldo, int B, int H, int W, int C, void *stream);
CY4_API int cy4_upsample2x_bwd(const void *gout, int64_t ldo, void *gin, int64_t ldi, int B, int H, int W, int C, int accumulate, void *stream);
CY4_API int cy4_maxpool_fwd(const void *in, int64_t ldi, void *out, int64_t ldo, int B, int H, int W, int C, int k, int stride, int pad, void *stream);
/* gscratch [B,H,W,C] fp32 (caller-zeroed) += routed gradients; follow with cy4_f32_to_f16 */
CY4_API int cy4_maxpool_bwd(const void *in, int64_t ldi, const void *gout, int64_t ldo, float *gscratch, int B, int H, int W, int C,
                            int k, int stride, int pad, void *stream);
/* The same pair with the argmax kept: the forward also writes, per output element, the window offset dy*k+dx of its first maximum
 * (uint8 [B,Ho,Wo,C], k*k <= 255); the backward routes the gradients from those indices instead of re-scanning the windows.
 * workspace: NULL, or 3*B*H*W*C bytes (16-byte aligned): stride-1 "same" pools then run as two separable passes (2k loads per output). */
CY4_API int cy4_maxpool_fwd_idx(const void *in, int64_t ldi, void *out, int64_t ldo, void *argmax, void *workspace, int B, int H, int W,
                                int C, int k, int stride, int pad, void *stream);
CY4_API int cy4_maxpool_bwd_idx(const void *argmax, const void *gout, int64_t ldo, float *gscratch, int B, int H, int W, int C, int k,
                                int stride, int pad, void *stream);
/* dst (+)= fp16(scale * (dscale ? *dscale : 1) * src); dscale is a device scalar */
CY4_API int cy4_f32_to_f16(const float *src, int64_t lds, float scale, const float *dscale, void *dst, int64_t ldd, int64_t M, int C,
                           int accumulate, void *stream);
/* Dynamic loss scale of the fp16 gradient tensors: amax[0] = max(amax[0], max|src|) (caller zeroes it);
 * scale2[0] = 2^k such that amax * 2^k ~ target, scale2[1] = 1 / scale2[0]. */
CY4_API int cy4_absmax_f32(const float *src, int64_t n, float *amax, void *stream);
CY4_API int cy4_make_scale(const float *amax, float target, float *scale2, void *stream);
/* out[c] (+)= scale * sum_m src[m, c]  (bias gradient of the head convs) */
CY4_API int cy4_colsum_f32(const float *src, int64_t lds, int64_t M, int C, float scale, float *out, int accumulate, void *stream);

/* ---- evaluation ---------------------------------------------------------
 * utils/evaluation_utils.py:186-210 iou_rotated_single_vs_multi_boxes_cpu, generalised to all pairs:
 * ious[i*m + j] = IoU(a6[i], b6[j]); rows (x, y, w, l, im, re).  fp32 corners, fp64 polygon intersection
 * (shapely in the reference), iou = reciprocal(area_a + area_b - inter + 1e-16) * inter in fp32. */
CY4_API int cy4_rbox_iou_matrix(const float *a6, int64_t n, const float *b6, int64_t m, float *ious, void *stream);
/* Anchor k-means distance matrix (utils/find_anchors.py:53-59 compute_iou for all boxes at once):
 * ious[i*k + j] = IoU of box i and cluster j, both (w, l, yaw) float64 rows centred at the origin.  Corners as
 * kitti_bev_utils.get_corners (float64 arithmetic stored as float32), polygon areas / intersection / ratio in float64
 * (shapely), + 1e-12 in the denominator, result rounded to float32 like the reference's np.float32 array. */
CY4_API int cy4_kmeans_iou(const double *boxes3, int64_t n, const double *clusters3, int k, float *ious, void *stream);

/* utils/evaluation_utils.py:322-357 post_processing_v2 on the device, one CTA per image.
 * pred [B, N, 7+nC] (x,y,w,l,im,re,conf,cls...); rows with conf >= conf_thresh, sorted by conf*max(cls)
 * (ties: lower row first), greedy rotated NMS over same-class boxes with IoU > nms_thresh, kept box =
 * confidence-weighted mean of the boxes it suppresses.  out9 [B, cy4_nms_max_candidates(), 9] rows
 * (merged box 6, conf, cls_conf, cls_pred), counts[B] kept rows, found[B] rows that passed the filter
 * (found > cy4_nms_max_candidates(): the list was truncated -- treat as an error).
 * workspace: cy4_nms_workspace_bytes(B) bytes. */
CY4_API int cy4_nms_max_candidates(void);
CY4_API size_t cy4_nms_workspace_bytes(int B);
CY4_API int cy4_nms_rotated_v2(const float *pred, int B, int N, int nC, float conf_thresh, float nms_thresh, float *out9,
                               int32_t *counts, int32_t *found, void *workspace, void *stream);

/* utils/evaluation_utils.py:152-183 get_batch_statistics_rotated_bbox: true-positive flags of the
 * detections of every image against targets8 [nT,8] rows (img, cls, x, y, w, l, im, re; x..l in pixels).
 * dets9 [B, max_det, 9] / counts[B] as written by cy4_nms_rotated_v2; tp [B, max_det] (0/1);
 * n_ann[B] annotations found per image (> cy4_eval_max_annotations(): truncated -- treat as an error). */
CY4_API int cy4_eval_max_annotations(void);
CY4_API int cy4_eval_match(const float *dets9, const int32_t *counts, int B, int max_det, const float *targets8, int64_t nT,
                           float iou_thresh, uint8_t *tp, int32_t *n_ann, void *stream);

/* ---- LiDAR -> BEV rasteriser ---------------------------------------------
 * data_process/kitti_bev_utils.py:18-36 removePoints + :39-76 makeBVFeature for a batch of frames.
 * points4: all frames' (x, y, z, intensity) fp32 rows back to back (16-byte aligned), offsets[B+1] row
 * offsets of the frames.  out [B, 3, H, W] fp32: channel 0 intensity of the highest point of the cell
 * (first in file order among equal heights), 1 its height / max_height, 2 min(1, log(count+1)/log(64)).
 * apply_filter = 1: removePoints first (inclusive bounds, z -= minZ); 0: points are already filtered / shifted
 * (the reference's makeBVFeature contract).  dropped[B]: points that fall outside even the reference's
 * (H+1) x (W+1) scratch map (numpy would wrap or raise there).  workspace: cy4_bev_workspace_bytes(). */
typedef struct cy4_bev_desc {
    float minX, maxX, minY, maxY, minZ, maxZ;
    float discretization;   /* metres per cell: (maxX - minX) / H, config/kitti_config.py:36 */
    float max_height;       /* float(abs(maxZ - minZ)), kitti_bev_utils.py:58 */
    int32_t H, W;           /* 608 x 608 */
    int32_t apply_filter;
    int32_t reserved;
} cy4_bev_desc;
CY4_API size_t cy4_bev_workspace_bytes(int B, int H, int W);
CY4_API int cy4_bev_rasterize(const float *points4, const int64_t *offsets, int B, const cy4_bev_desc *desc, float *out,
                              int32_t *dropped, void *workspace, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CY4_H */
