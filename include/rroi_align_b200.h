/*
 * rroi_align_b200.h -- C ABI of the B200-native RoIRotate (rotated-RoI align) library,
 * fots/pytorch_b200/lib/librroi_b200.so.  Plain pointers and sizes; no torch types.
 *
 * All pointers are DEVICE pointers (fp32) owned by the caller.  Every entry point only enqueues
 * work on `stream` (no synchronisation), is re-entrant, and keeps no state besides the tuning
 * knobs.  RoI rows are [batch_idx, cx, cy, h, w, angle_deg] in input-image pixels
 * (reference: rroi_align/src/rroi_align_kernel.cu:58-65).
 *
 * Reference interfaces replaced (file:line relative to chenjun2hao/FOTS.pytorch):
 *   RROIAlignForwardLaucher / RROIAlignBackwardLaucher   rroi_align/src/rroi_align_kernel.h:8-18
 *       -- exported here with the SAME names and signatures (the reference spells "Laucher").
 *   rroi_align_forward_cuda / rroi_align_backward_cuda   rroi_align/src/rroi_align_cuda.h:2-8
 *       -- THC tensor unpacking; replaced by rroi_b200_forward / rroi_b200_backward, which take the
 *          raw pointers + sizes that wrapper extracted (rroi_align_cuda.c:12-33, :56-77).
 */
#ifndef RROI_ALIGN_B200_H_
#define RROI_ALIGN_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __DRIVER_TYPES_H__
typedef struct CUstream_st* cudaStream_t;   /* same definition as the CUDA runtime's */
#endif

/* ---- status codes of the rroi_b200_* entry points (0 = success) ---- */
#define RROI_B200_OK                0
#define RROI_B200_ERR_INVALID_ARG  -1   /* null pointer, non-positive size, unknown layout/flag */
#define RROI_B200_ERR_TOO_LARGE    -2   /* launch grid would exceed 2^31-1 CTAs */
#define RROI_B200_ERR_CUDA         -3   /* launch failed; see rroi_b200_last_cuda_error() */

/* ---- feature-map / pooled-output memory layouts ---- */
#define RROI_B200_LAYOUT_NCHW 0   /* features [B,C,H,W], pooled [N,C,PH,PW]  (reference layout) */
#define RROI_B200_LAYOUT_NHWC 1   /* features [B,H,W,C], pooled [N,PH,PW,C]  (channels-last)     */

/* ---- tuning keys for rroi_b200_set_tuning / rroi_b200_get_tuning ---- */
#define RROI_B200_TUNE_NCHW_CG      0   /* channels per CTA in the NCHW kernels: 1,2,4,8,16 (0 = default 8) */
#define RROI_B200_TUNE_NHWC_UNROLL  1   /* NHWC forward variant: 0 = auto, 1..6 = fixed tile size x loads in flight  */
#define RROI_B200_TUNE_USE_PDL      2   /* 1: launch with programmatic dependent launch                       */
#define RROI_B200_TUNE_BWD_DEDUPE   3   /* NCHW: 1 (default) warp-merge equal sample points, 0 off; 2 = generic NHWC kernel */
#define RROI_B200_TUNE_BWD_ZERO_FUSED 5 /* channels-last backward with zero_fill: 0 (default) memset + scatter; 1 = one-pass zero + scatter for maps >= 96 MB (measured slower) */
#define RROI_B200_TUNE_NCHW_TMA     4   /* NCHW forward: 0 (default) gather kernel; 1 = stage each patch's footprint with TMA box loads; 2..5 = same with a minimum box size (sweeps) */

/*
 * Drop-in for rroi_align/src/rroi_align_kernel.h:8-12.  NCHW.  top_data / con_idx_x / con_idx_y are
 * [num_rois, channels, pooled_height, pooled_width].  The reference accumulates into caller-zeroed
 * buffers; this overwrites every element (zeros for pw > roi_pooled_width), so the result on zeroed
 * buffers is identical and pre-zeroing is no longer required.  con_idx_x and con_idx_y may both be
 * NULL.  Returns 1 on success like the reference; 0 on invalid arguments; a negative RROI_B200_ERR_*
 * on a CUDA error (the reference prints and calls exit(-1), rroi_align_kernel.cu:179-184).
 * RoIs whose batch index is outside [0, batch) cannot be detected here (no batch argument): the
 * caller must guarantee them, as with the reference.
 */
int RROIAlignForwardLaucher(const float* bottom_data, const float spatial_scale, const int num_rois,
                            const int height, const int width, const int channels,
                            const int pooled_height, const int pooled_width, const float* bottom_rois,
                            float* top_data, float* con_idx_x, float* con_idx_y, cudaStream_t stream);

/*
 * Drop-in for rroi_align/src/rroi_align_kernel.h:14-18.  NCHW.  Accumulates into bottom_diff
 * [batch_size, channels, height, width], which the caller zero-fills (functions/rroi_align.py:35).
 * con_idx_x / con_idx_y are the [N,C,PH,PW] tensors the forward wrote and are read per element.
 */
int RROIAlignBackwardLaucher(const float* top_diff, const float spatial_scale, const int batch_size,
                             const int num_rois, const int height, const int width, const int channels,
                             const int pooled_height, const int pooled_width, const float* bottom_rois,
                             float* bottom_diff, const float* con_idx_x, const float* con_idx_y,
                             cudaStream_t stream);

/*
 * v2 forward (replaces rroi_align_forward_cuda, rroi_align_cuda.h:2-4).
 *   features  [batch,channels,height,width] (NCHW) or [batch,height,width,channels] (NHWC)
 *   rois      [num_rois,6]
 *   pooled    [num_rois,channels,PH,PW] or [num_rois,PH,PW,channels]; every element is written
 *   idx_x/y   optional COMPACT sample centres [num_rois,PH,PW] (they do not depend on the channel;
 *             0 where pw > roi_pooled_width); pass NULL for both to skip them
 * RoIs with a batch index outside [0,batch) produce zeros instead of the reference's out-of-bounds read.
 */
int rroi_b200_forward(const float* features, const float* rois, float* pooled, float* idx_x, float* idx_y,
                      int num_rois, int batch, int channels, int height, int width,
                      int pooled_height, int pooled_width, float spatial_scale, int layout,
                      cudaStream_t stream);

/*
 * bf16 forward for the inference pipeline (hand the recogniser's first convolution the
 * dtype and layout it wants).  Channels-last only:
 *   features  bf16 [batch,height,width,channels]     pooled  bf16 [num_rois,PH,PW,channels]
 *   channels in {32, 64, 128, 256}; both pointers 16-byte aligned; rois / idx_x / idx_y fp32 as above.
 * Same geometry and the same fp32 4-product sum as rroi_b200_forward (rroi_align_kernel.cu:86-141) on the
 * widened taps; the fp32 result is rounded to bf16 once (RNE): pooled == bf16(rroi_b200_forward(float(features))).
 */
int rroi_b200_forward_bf16(const void* features, const float* rois, void* pooled, float* idx_x, float* idx_y,
                           int num_rois, int batch, int channels, int height, int width,
                           int pooled_height, int pooled_width, float spatial_scale, cudaStream_t stream);

/*
 * v2 backward (replaces rroi_align_backward_cuda, rroi_align_cuda.h:6-8).
 *   top_diff     pooled-shaped gradient, in `layout`
 *   idx_x/y      the compact centres saved by rroi_b200_forward, or NULL/NULL to recompute them
 *   bottom_diff  feature-shaped gradient, in `layout`; if zero_fill != 0 it is cleared on `stream`
 *                first (cudaMemsetAsync), otherwise it is accumulated into
 */
int rroi_b200_backward(const float* top_diff, const float* rois, const float* idx_x, const float* idx_y,
                       float* bottom_diff, int num_rois, int batch, int channels, int height, int width,
                       int pooled_height, int pooled_width, float spatial_scale, int layout,
                       int zero_fill, cudaStream_t stream);

/* Expand compact centres [N,PH,PW] to the reference's [N,C,PH,PW] (ctx.idx_x / ctx.idx_y attribute parity). */
int rroi_b200_expand_idx(const float* idx_compact, float* idx_full, int num_rois, int channels,
                         int pooled_height, int pooled_width, cudaStream_t stream);

int         rroi_b200_set_tuning(int key, int value);   /* returns RROI_B200_OK or ERR_INVALID_ARG */
int         rroi_b200_get_tuning(int key);              /* current value, or -1 for an unknown key  */
int         rroi_b200_last_cuda_error(void);            /* cudaError_t of the last failed launch (per process) */
const char* rroi_b200_strerror(int status);
int         rroi_b200_abi_version(void);                /* bumped on any signature change */
const char* rroi_b200_build_info(void);                 /* e.g. "sm_100a nvcc 12.9" */

#ifdef __cplusplus
}
#endif
#endif /* RROI_ALIGN_B200_H_ */
