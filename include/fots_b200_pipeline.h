/*
 * fots_b200_pipeline.h -- C ABI of the two small kernels either side of RoIRotate in the end-to-end path
 * (SURVEY.md section 8f-1 and 8f-4).  Same library (librroi_b200.so), same conventions as rroi_align_b200.h:
 * device pointers, enqueue-only on `stream`, 0 = success, negative RROI_B200_ERR_* otherwise.
 *
 * Reference code replaced (chenjun2hao/FOTS.pytorch):
 *   fots_b200_boxes_to_rois   tools/ocr_utils.py:133-145  (per-box Python: quad -> [b, cx, cy, h, w, -angle_deg])
 *   fots_b200_ctc_greedy      tools/ocr_utils.py:183-186 + src/utils.py:93-97 (argmax over classes, collapse
 *                             repeats, drop blank 0 -- done per box on the host in the reference)
 */
#ifndef FOTS_B200_PIPELINE_H_
#define FOTS_B200_PIPELINE_H_

#include "rroi_align_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/*
 * quads      [num_boxes, quad_stride] fp32, the first 8 values of a row are x0,y0,...,x3,y3 in image pixels
 *            (the row format of nms/adaptor.cpp:13-29 has 9: + score)
 * batch_idx  [num_boxes] int32 image index of each box, or NULL for all-zero
 * rois       [num_boxes, 6] fp32 out: [batch_idx, trunc(cx), trunc(cy), h, w, -angle_deg]; arithmetic in fp64 like
 *            the reference's Python floats, angle = -atan2(y2-y1, x2-x1) / 3.1415926535 * 180
 */
int fots_b200_boxes_to_rois(const float* quads, int quad_stride, const int* batch_idx, int num_boxes,
                            float* rois, cudaStream_t stream);

/*
 * logp    [num_seq, num_classes, T] fp32 (the layout forward_ocr returns, tools/models.py:370-379)
 * ids     [num_seq, T] int32 out: decoded class ids, left-aligned, zero-padded
 * lengths [num_seq] int32 out
 * Greedy CTC: per time step the arg-max class (lowest index on ties, like torch.max), then drop repeats and
 * blanks (class 0).  T <= 1024.
 */
int fots_b200_ctc_greedy(const float* logp, int num_seq, int num_classes, int T, int* ids, int* lengths,
                         cudaStream_t stream);

/*
 * Detection decode: the per-pixel loop of nms/adaptor.cpp:76-117 -- threshold the score
 * map, build one quadrangle per positive pixel from the four distances and (sin, cos), in x10000 fixed point --
 * on the GPU, compacted in RASTER ORDER (the reference's row-wise merge nms/nms.h:149-213 depends on that order),
 * so that the CPU Clipper merge receives a compact candidate list instead of three full maps (test.py:86-96 copies
 * all three to the host with three synchronisations).
 *   segm   fp32 [B, h, w]      score map (seg_pred[0] squeezed)
 *   rbox   fp32 [B, 4, h, w]   distances top, bottom, left, right (NCHW as the network emits them)
 *   angle  fp32 [B, 2, h, w]   (sin, cos)
 *   counts int32 [B] out       number of positive pixels per image (may exceed max_per_image; rows beyond it are dropped)
 *   cand   int32 [B, max_per_image, 16] out, one row per positive pixel in raster order:
 *            [0..7] x0,y0,..,x3,y3 as cInt(roundf(10000 * coord))   [8] score (fp32 bits)
 *            [9..12] p_left*p_bt, p_left*p_top, p_right*p_top, p_right*p_bt (fp32 bits)   [13] x  [14] y  [15] 0
 *   scratch int32 [B * ceil(h*w/256)] per-block counts
 * fp32 arithmetic in the reference's operation order without fused multiply-adds (its Makefile builds with -O3 on
 * baseline x86-64), so the fixed-point coordinates are bit-identical; expf() is CUDA's (<= 2 ulp from libm).
 */
int fots_b200_decode_candidates(const float* segm, const float* rbox, const float* angle, int B, int h, int w,
                                float segm_threshold, int max_per_image, int* counts, int* cand, int* scratch,
                                cudaStream_t stream);

/*
 * HOST function (CPU, no stream): the locality-aware merge + standard NMS of the reference's detector
 * post-processing (nms/nms.h:149-213 merge_iou, :112-146 standard_nms, :49-109 PolyMerger; nms/adaptor.cpp:118),
 * over the compact raster-ordered candidate rows fots_b200_decode_candidates produced, after ONE device-to-host copy
 * of `cand[0 .. num_cand)`.  Sequential by construction, hence CPU like the reference's.
 *   cand        host int32 [num_cand, 16], rows as documented above (one image)
 *   w, h        size of the score map the candidates came from
 *   iou_threshold1 / 2   merge thresholds of the two stages (nms/__init__.py:29 passes 0.4 and 0.2)
 *   boxes       host fp32 [max_boxes, 9] out: x0,y0..x3,y3 in the x10000 fixed-point units (divide by 10000 for pixels,
 *               nms/__init__.py:14) + accumulated score, in the order the reference returns them
 *   num_boxes   out: number of polygons kept (may exceed max_boxes; only the first max_boxes are written)
 * Polygon IoU is an exact-input convex clip in double precision (the reference calls Clipper; same value for the convex
 * quadrangles the decode emits, up to Clipper's integer rounding of intersection vertices).
 */
int fots_b200_merge_candidates_host(const int* cand, int num_cand, int w, int h, float iou_threshold1,
                                    float iou_threshold2, float* boxes, int max_boxes, int* num_boxes);
/*
 * The same merge for a micro-batch of B images in one call, dealt to `threads` host threads (images are independent;
 * the per-rank worker pool of the end-to-end step).  cand [B, cap, 16] / counts [B] as fots_b200_decode_candidates
 * writes them (host copies); boxes [B, max_boxes, 9], num_boxes [B] (totals: may exceed max_boxes, rows beyond it are
 * dropped and the caller must check).
 */
int fots_b200_merge_candidates_host_batch(const int* cand, const int* counts, int B, int cap, int w, int h,
                                          float iou_threshold1, float iou_threshold2, float* boxes, int max_boxes,
                                          int* num_boxes, int threads);

/*
 * Fused channels-last InstanceNorm (+ affine) (+ residual add) + leaky-ReLU for the feeder/consumer networks
 * (tools/models.py:41-48 CReLU_IN, :142-166 BasicBlockIn, :87-103 conv_dw_*_in, :336-364 forward_ocr).  torch's
 * instance_norm converts a channels-last tensor to NCHW and back around a batch-norm kernel; this is one
 * statistics pass and one normalise pass over the NHWC tensor, HBM-bound.
 *   x         bf16 [B, HW, C]  (channels-last activations, C % 8 == 0, C <= 1024)
 *   y         bf16 [B, HW, C]  or, with crelu != 0, [B, HW, 2C] = act(IN(concat(x, -x)))  (gamma/beta then [2C])
 *   gamma/beta fp32 [C] ([2C] with crelu) or both NULL (non-affine)
 *   residual  optional bf16 [B, HW, C] added after the affine, before the activation (crelu == 0 only)
 *   workspace fp64 [B, C, 2]; cleared and filled by the call (sum, sum of squares)
 *   slope     leaky-ReLU negative slope: 0 = ReLU, 1 = no activation
 * Statistics are accumulated in fp32 per thread and fp64 across the plane; variance is the biased one.
 */
int fots_b200_instnorm_nhwc_bf16(const void* x, void* y, const float* gamma, const float* beta,
                                 const void* residual, double* workspace, int B, int HW, int C,
                                 float eps, float slope, int crelu, cudaStream_t stream);

/*
 * MaxPool2d((2,1), stride (2,1)) of a channels-last bf16 tensor (`max2`, tools/models.py:344, :360):
 *   x bf16 [N, H, W, C] -> y bf16 [N, H/2, W, C];  C % 8 == 0, 16-byte aligned.  HBM-bound, 16-byte accesses.
 */
int fots_b200_maxpool_h2_nhwc_bf16(const void* x, void* y, int N, int H, int W, int C, cudaStream_t stream);

/*
 * One step of the top-down feature merge of tools/models.py:411-438, fused (channels-last bf16, fp32 arithmetic):
 *     y = (a_lo ? upsample(a_lo) : c_hi)  +  (b_hi ? b_hi * (g_lo ? upsample(sigmoid(g_lo)) : 1) : 0)
 * upsample = bilinear, align_corners=True (torch's source-index arithmetic), from [h, w] to [H, W].
 *   a_lo  bf16 [B, h, w, C] or NULL      c_hi  bf16 [B, H, W, C] or NULL   (exactly one of the two)
 *   b_hi  bf16 [B, H, W, C] or NULL      g_lo  bf16 [B, h, w, 1] attention LOGITS or NULL (needs b_hi)
 *   y     bf16 [B, H, W, C];  C % 8 == 0.
 * torch runs this as upsample + upsample + mul + add (its NHWC bf16 upsample kernel reaches ~0.55 TB/s).
 */
int fots_b200_fpn_merge_nhwc_bf16(const void* a_lo, const void* c_hi, const void* b_hi, const void* g_lo, void* y,
                                  int B, int h, int w, int H, int W, int C, cudaStream_t stream);
/* The same with the gate map already holding sigmoid(logit) as bf16 (what torch's bf16 autocast path interpolates; emitted by
 * fots_b200_conv1x1_to1_nhwc_bf16 with sigmoid = 1): saves four expf + divisions per 16-byte output vector. */
int fots_b200_fpn_merge_prob_nhwc_bf16(const void* a_lo, const void* c_hi, const void* b_hi, const void* g_prob_lo, void* y,
                                       int B, int h, int w, int H, int W, int C, cudaStream_t stream);

/*
 * Stride-1 convolution of channels-last bf16 activations on the sm_100a tensor cores (tcgen05.mma with TMEM
 * accumulators, TMA-staged operands; csrc/conv_tc.cu), bias and leaky-ReLU fused in the epilogue.  Replaces the
 * cuDNN calls behind nn.Conv2d for the dense convolutions of forward_ocr (tools/models.py:336-366: conv5..conv10_s)
 * and of the feeder's BasicBlockIn stages (tools/models.py:142-166).
 *   x     bf16 [N, H, W, Cin]            (NHWC storage; Cin % 64 == 0)
 *   w     bf16 [Cout, R, S, Cin]         (the storage of a channels-last [Cout, Cin, R, S] weight; Cout % 64 == 0)
 *   bias  fp32 [Cout] or NULL
 *   y     bf16 [N, Ho, Wo, Cout],  Ho = H + 2*pad_h - R + 1,  Wo = W + 2*pad_w - S + 1
 *   slope leaky-ReLU negative slope applied to (conv + bias): 1 = none, 0 = ReLU
 * fp32 accumulation; the result is rounded to bf16 once.  All pointers 16-byte aligned.
 */
int fots_b200_conv2d_nhwc_bf16(const void* x, const void* w, const float* bias, void* y, int N, int H, int W,
                               int Cin, int Cout, int R, int S, int pad_h, int pad_w, float slope,
                               cudaStream_t stream);
/*
 * The same convolution (no activation) that also accumulates, in its epilogue, the per-image per-channel sum and sum
 * of squares of the bf16 values it stores -- the statistics pass of the InstanceNorm that follows every such
 * convolution in tools/models.py (:336-338 conv5+batch5, :346-347, :362-364, :148-160 BasicBlockIn) -- so that
 * fots_b200_instnorm_apply_nhwc_bf16 can normalise without reading y a first time.
 *   stats  fp64 [N, Cout, 2], cleared and filled by the call (same format as the instnorm workspace)
 */
int fots_b200_conv2d_stats_nhwc_bf16(const void* x, const void* w, const float* bias, void* y, double* stats,
                                     int N, int H, int W, int Cin, int Cout, int R, int S, int pad_h, int pad_w,
                                     cudaStream_t stream);
/* The same convolution with spatial stride 1 or 2: y [N, (H + 2 pad_h - R) / stride + 1, (W + 2 pad_w - S) / stride + 1, Cout]. */
int fots_b200_conv2d_strided_nhwc_bf16(const void* x, const void* w, const float* bias, void* y, int N, int H, int W,
                                       int Cin, int Cout, int R, int S, int pad_h, int pad_w, int stride, float slope,
                                       cudaStream_t stream);
/* fots_b200_instnorm_nhwc_bf16 without its statistics pass: `stats` [B, C, 2] comes from the call above. */
int fots_b200_instnorm_apply_nhwc_bf16(const void* x, void* y, const float* gamma, const float* beta,
                                       const void* residual, const double* stats, int B, int HW, int C,
                                       float eps, float slope, int crelu, cudaStream_t stream);
/* fots_b200_instnorm_nhwc_bf16 takes ONE launch (a cluster per instance keeps it in shared memory, x read once) for small
 * instances (<= 40 KB per CTA: several instances per SM; measured slower than the two-pass form above that).
 * mode 0 = never (A/B timing), 1 = automatic (default), 2 = whenever the instance fits 160 KB per CTA (tests, sweeps). */
int fots_b200_instnorm_set_single_pass(int mode);
/*
 * The feature extractor's first layer (tools/models.py:250-251: Conv2d(3, 16, 3, stride 1, pad 1, bias=False)) fused
 * with the statistics pass of the CReLU_IN that follows it (tools/models.py:41-48); csrc/stem_conv.cu.
 *   x      fp32 [B, H, W, 3]   (channels-last image; rounded to bf16 on load, as autocast does for the library call)
 *   w      bf16 [16, 3, 3, 3]  = [cout][r][s][cin] (the storage of the channels-last weight)
 *   y      bf16 [B, H, W, 16]
 *   stats  fp64 [B, 16, 2] sum / sum of squares of the bf16 values stored (cleared by the call), or NULL
 * HBM-bound (27 MACs per output value); the arithmetic runs on mma.sync fragments built from a shared-memory tile.
 */
int fots_b200_stem_conv3x3_c3_c16(const float* x, const void* w, void* y, double* stats, int B, int H, int W,
                                  cudaStream_t stream);
/*
 * Same layer fed with the RAW image: x = uint8 [B, H, W, 3] as cv2.imread returns it; the reference's host-side
 * preprocessing (test.py:80-83: images /= 128; images -= 1) is applied on load -- exact in fp32, so y and stats are
 * bit-identical to the fp32 entry point on the preprocessed image, and the upload is a quarter of the bytes.
 */
int fots_b200_stem_conv3x3_c3_c16_u8(const unsigned char* x, const void* w, void* y, double* stats, int B, int H, int W,
                                     cudaStream_t stream);
/*
 * Consumer B's recurrent half (tools/models.py:17-33 BidirectionalLSTM = nn.LSTM(bidirectional) + nn.Linear; :898-909).
 * csrc/lstm_kernels.cu.  One layer = gemm (input projections of all time steps, both directions) -> recurrent kernel
 * (the whole time loop of both directions in one launch) -> gemm (the embedding Linear).
 *
 * fots_b200_gemm_bf16w:  C[M, N] fp32 = A[M, K] * W[N, K]^T + bias[N]
 *   A  bf16 (a_is_f32 = 0) or fp32 (a_is_f32 = 1: split into bf16 hi + lo parts, two tensor-core products, so nothing
 *      of an fp32 activation is lost); W bf16 [N, K] as nn.Linear / nn.LSTM store their weights; bias fp32 or NULL;
 *      K % 32 == 0; A and W 16-byte aligned.
 * fots_b200_bilstm_recurrent:  H must be 256 (the CRNN's hidden size).
 *   G    fp32 [T, N, 2, 4H]  x_t * W_ih^T + b_ih + b_hh for (forward, reverse), gate order i, f, g, o (nn.LSTM)
 *   Whh  bf16 [2, 4H, H]     weight_hh_l0, weight_hh_l0_reverse
 *   Y    fp32 [T, N, 2H]     h_t, forward direction in [.., :H], reverse in [.., H:]  (what nn.LSTM returns), h_0 = c_0 = 0
 */
int fots_b200_gemm_bf16w(const void* A, int a_is_f32, const void* W, const float* bias, float* C, int M, int N, int K,
                         cudaStream_t stream);
int fots_b200_bilstm_recurrent(const float* G, const void* Whh, float* Y, int T, int N, int H, cudaStream_t stream);
/*
 * Depthwise 3x3 convolution, pad 1, stride 1 or 2, no bias (the first halves of BasicBlockSepIn and of upconv1/2,
 * tools/models.py:59-111, :300-301): x bf16 [N, H, W, C], w bf16 [C, 3, 3] (the storage of a [C, 1, 3, 3] weight),
 * y bf16 [N, (H - 1) / stride + 1, (W - 1) / stride + 1, C]; C % 64 == 0.  HBM-bound; csrc/dwconv_kernels.cu.
 */
int fots_b200_dwconv3x3_nhwc_bf16(const void* x, const void* w, void* y, int N, int H, int W, int C, int stride,
                                  cudaStream_t stream);
/*
 * The detection heads in one pass (tools/models.py:440-456): seg = sigmoid(act(x)), rbox = sigmoid(rbox(x)) * 128,
 * angle = normalise(sigmoid(angle(x)) * 2 - 1), three 1x1 convolutions with 1 + 4 + 2 output channels.
 *   x     bf16 [B, H, W, C], C in {128, 256, 512}
 *   wq    bf16 [8, C]: rows {act, zeros, rbox0, rbox1, rbox2, rbox3, angle0, angle1}   bias fp32 [8], same order
 *   seg fp32 [B, 1, H, W]   rbox fp32 [B, 4, H, W]   angle fp32 [B, 2, H, W]   (NCHW, as the reference returns them)
 * HBM-bound: x is read once (the library path reads it three times and launches ~15 element-wise kernels).
 */
int fots_b200_heads_nhwc_bf16(const void* x, const void* wq, const float* bias, float* seg, float* rbox, float* angle,
                              int B, int H, int W, int C, cudaStream_t stream);
/* Same with the InstanceNorms either side of it fused:
 *   stats != NULL: computes dw(act(IN(x))) -- the InstanceNorm (+ affine gamma / beta, or both NULL) + leaky-ReLU(slope)
 *     between the pointwise and the depthwise half of a separable block is applied to the staged tile in shared memory, from
 *     `stats` [N, C, 2] (the sums of x, fots_b200_instnorm_stats_nhwc_bf16).  The normalised tensor never exists in HBM.
 *   stats_out != NULL: [N, C, 2] (cleared by the call) receives the sums of the bf16 outputs -- the statistics pass of the
 *     InstanceNorm that FOLLOWS, accumulated in the epilogue (feeds fots_b200_instnorm_apply_nhwc_bf16). */
int fots_b200_dwconv3x3_norm_nhwc_bf16(const void* x, const void* w, void* y, const double* stats, const float* gamma,
                                       const float* beta, float eps, float slope, double* stats_out, int N, int H, int W,
                                       int C, int stride, cudaStream_t stream);
/* dw(upsample(x_lo)): bilinear (align_corners = True) upsampling of x_lo [N, h, w, C] to H x W computed while the depthwise
 * kernel stages its tile (tools/models.py:418-436: upconv(F.interpolate(x))): the upsampled map is never written. */
int fots_b200_dwconv3x3_up_nhwc_bf16(const void* x_lo, const void* w, void* y, int N, int h, int wlo, int H, int W, int C,
                                     cudaStream_t stream);
/* The statistics pass of fots_b200_instnorm_nhwc_bf16 on its own: workspace [B, C, 2] fp64 (cleared by the call). */
int fots_b200_instnorm_stats_nhwc_bf16(const void* x, double* workspace, int B, int HW, int C, cudaStream_t stream);
/* A 1x1 convolution to ONE output channel + bias (the attention gate conv_attenton of the top-down merge,
 * tools/models.py:405-438) -> bf16 logits [B, 1, H, W] (what fots_b200_fpn_merge_nhwc_bf16 takes as gate_logits), or with
 * sigmoid != 0 sigmoid(logit) (what fots_b200_fpn_merge_prob_nhwc_bf16 takes):
 * x bf16 [B, H, W, C], wq bf16 [8, C] with the filter in row 0 and zeros elsewhere, bias fp32 [8].  Same kernel as the heads. */
int fots_b200_conv1x1_to1_nhwc_bf16(const void* x, const void* wq, const float* bias, void* out, int B, int H, int W, int C,
                                    int sigmoid, cudaStream_t stream);
/* Backward of fots_b200_instnorm_nhwc_bf16 (crelu == 0) for the training step (train.py:79-123 runs the same InstanceNorm
 * layers under autograd; torch's path copies channels-last tensors to NCHW and back around batch-norm kernels, forward and
 * backward).  y = act(IN(x) * gamma + beta [+ residual]) as stored by the forward, dy its gradient (bf16 [B, HW, C]):
 *     g = dy * act'(.)  (1 where y > 0, else slope),  xh = (x - mean) * rstd
 *     dx = gamma * rstd * (g - mean_hw(g) - xh * mean_hw(g * xh)),   dres = g (optional, NULL to skip)
 * stats = the forward's fp64 [B, C, 2] sums of x; stats_bwd fp64 [B, C, 2] is cleared and filled with (sum g, sum g * xh) per
 * image and channel -- dbeta[c] = sum_b stats_bwd[b, c, 0], dgamma[c] = sum_b stats_bwd[b, c, 1].  Two HBM passes. */
int fots_b200_instnorm_bwd_nhwc_bf16(const void* x, const void* y, const void* dy, const float* gamma, const double* stats,
                                     double* stats_bwd, void* dx, void* dres, int B, int HW, int C, float eps, float slope,
                                     cudaStream_t stream);
/* The same for the CReLU form (crelu != 0 in the forward: y = act(IN(concat(x, -x)) * gamma + beta), 2C output channels):
 * y / dy bf16 [B, HW, 2C], gamma fp32 [2C] or NULL, stats_bwd fp64 [B, 2C, 2] (dbeta[j] = sum_b [b, j, 0], dgamma[j] = sum_b
 * [b, j, 1] for all 2C channels). */
int fots_b200_instnorm_crelu_bwd_nhwc_bf16(const void* x, const void* y, const void* dy, const float* gamma, const double* stats,
                                           double* stats_bwd, void* dx, int B, int HW, int C, float eps, float slope,
                                           cudaStream_t stream);
/* Backward of the align_corners bilinear upsampling of the top-down merge (the a_lo-only form of fots_b200_fpn_merge_nhwc_bf16;
 * tools/models.py:411-438 F.interpolate under autograd): dlo bf16 [B, h, w, C] = U^T dhi bf16 [B, H, W, C].  Gather form, no atomics. */
int fots_b200_upsample_bilinear_bwd_nhwc_bf16(const void* dhi, void* dlo, int B, int h, int w, int H, int W, int C, cudaStream_t stream);
/* The last level of the top-down merge collapsed into the detection heads (inference; tools/models.py:430-456).  There
 *     x = upconv2_pw(d) + feature1(s) * upsample(sigmoid(conv_attenton(f2))),   heads = squash(Wh x + bh)
 * and x feeds nothing but the heads, so with the products Wh Wpw (8 x 256) and Wh Wf1 (8 x 64) folded by the caller
 *     logits = w1 . x1 + upsample(gate_prob) * (w2 . x2) + bias
 * x1 = d bf16 [B, H, W, 256] (the depthwise half of upconv2), x2 = s bf16 [B, H, W, 64] (the stage-1 output), gate_prob bf16
 * [B, gh, gw]; w1 / w2 in the 8-row head layout of fots_b200_heads_nhwc_bf16, outputs as there.  The pointwise 256 -> 256 and
 * lateral 64 -> 256 convolutions, the merge kernel and the 256-channel map x are never computed. */
int fots_b200_heads_merged_nhwc_bf16(const void* x1, const void* w1, const void* x2, const void* w2, const void* gate_prob,
                                     const float* bias, float* seg, float* rbox, float* angle, int B, int H, int W, int C1, int C2,
                                     int gh, int gw, cudaStream_t stream);
/* ... with the depthwise half of upconv2 and its 2x upsampling folded in as well (both linear; the bilinear upsampling
 * commutes with a per-pixel channel mix):  (Wh Wpw) dw3x3(up(f2)) (p) = sum_tap up(T_tap)(p + tap),  T_tap = (Wh Wpw) diag(w_dw[:, tap]) f2,
 * zero where p + tap leaves the map.  T bf16 [B, h, w, TC] (channel tap * 8 + head column, TC >= 72 and a multiple of 8) is one
 * 1x1 convolution of the low-resolution map f2 with the caller-folded weight (fots_b200_conv2d_nhwc_bf16, output channels
 * padded with zero filters); this entry point gathers 9 taps x 4 bilinear samples (align_corners) per full-resolution pixel.
 * The upsample-on-load depthwise convolution and its 256-channel output at 1/4 scale are never computed.  x2, w2, gate_prob
 * [B, h, w], bias, outputs: as above. */
int fots_b200_heads_gather_nhwc_bf16(const void* T, int TC, const void* x2, const void* w2, const void* gate_prob, const float* bias,
                                     float* seg, float* rbox, float* angle, int B, int H, int W, int h, int w, int C2,
                                     cudaStream_t stream);
/* Consumer B's first layer (tools/models.py:853-897, CRNN.cnn conv0 + relu0 + pooling0): 3 input channels cannot fill a
 * k-block of the tcgen05 kernel.  x fp32 NCHW [N, 3, H, W] (RoIRotate of the raw image, src/utils.py:430-436), w bf16
 * [Cout, 3, 3, 3] contiguous, bias fp32 [Cout] or NULL -> y bf16 NHWC = maxpool2x2(relu(conv3x3_pad1(x) + bias)) when
 * pool2x2 != 0 ([N, H/2, W/2, Cout]), else relu(conv + bias) ([N, H, W, Cout]).  Cout % 8 == 0, Cout <= 256. */
int fots_b200_conv3x3_c3_pool_nhwc_bf16(const float* x, const void* w, const float* bias, void* y, int N, int H, int W, int Cout,
                                        int pool2x2, cudaStream_t stream);
/* MaxPool2d((kh, kw), stride (sh, sw), padding (ph, pw)), floor mode, of a channels-last bf16 tensor (CRNN.cnn pooling0-3):
 * x [N, H, W, C] -> y [N, (H + 2 ph - kh) / sh + 1, (W + 2 pw - kw) / sw + 1, C]; padding counts as -inf, NaNs propagate. */
int fots_b200_maxpool_nhwc_bf16(const void* x, void* y, int N, int H, int W, int C, int kh, int kw, int sh, int sw, int ph, int pw,
                                cudaStream_t stream);
/* The three fots_b200_*_set_* switches below are process-wide A/B switches for sweeps and parity tests (not thread-safe,
 * never needed for correct results); everything a production caller selects travels with the call.
 * Output-channel tile of fots_b200_conv2d_nhwc_bf16: 0 = automatic, or 64 / 128 / 256 (for sweeps). */
int fots_b200_conv_set_tile(int bn);
/* Halo reuse of the A operand (3x3 stride-1 convolutions, 64- / 128-wide cout tiles: one TMA load of the tile + halo rows
 * feeds three filter taps; 64 -> 64 channels: one box feeds all nine, weights resident): -1 = automatic (maps of at least
 * 32 x 16 pixels), 0 = never, 1 = whenever the shape allows, 2 = same but always the three-copy form. */
int fots_b200_conv_set_halo(int mode);

#ifdef __cplusplus
}
#endif
#endif /* FOTS_B200_PIPELINE_H_ */
