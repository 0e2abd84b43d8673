/*
 * rroi_align_b200.h -- C ABI of the B200-native RoIRotate (rotated-RoI align) library,
 * fots/pytorch_b200/lib/librroi_b200.so.  Plain pointers and sizes; no torch types.
 *
 * All pointers are DEVICE pointers (fp32) owned by the caller.  Every entry point only enqueues
 * work on `stream` (no synchronisation), is re-entrant and thread-safe, and keeps no mutable state:
 * kernel variants are selected per call (rroi_b200_opts).  RoI rows are [batch_idx, cx, cy, h, w, angle_deg] in input-image pixels
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

/*
 * ---- per-call launch options (the *_opt entry points; NULL = all defaults) ----
 * The library keeps NO mutable state: everything that selects a kernel variant travels with the call, so threads
 * and streams that want different variants cannot race.  `size` must be sizeof(rroi_b200_opts) of the caller's
 * header (the struct may grow at the end; missing fields read as 0).
 */
#define RROI_B200_FLAG_NO_PDL      1u   /* launch without programmatic dependent launch (default: with)            */
#define RROI_B200_FLAG_ROIS_READY  2u   /* the RoI rows were complete in device memory before the kernel that      */
                                        /* PRECEDES this launch on `stream` was enqueued (uploaded by a copy, or   */
                                        /* written two or more kernels ago): the forward may then read them and    */
                                        /* compute the per-RoI transform and the bin geometry while that kernel    */
                                        /* still drains (before griddepcontrol.wait).  Never set it when the       */
                                        /* immediately preceding kernel writes the RoI rows.                       */
typedef struct rroi_b200_opts {
    unsigned int size;         /* sizeof(rroi_b200_opts)                                                            */
    unsigned int flags;        /* RROI_B200_FLAG_*                                                                  */
    int concurrency;           /* independent RoIRotate launches the caller keeps in flight on OTHER streams /      */
                               /* graph branches (0 or 1: this launch has the GPU to itself).  Picks the tile size: */
                               /* one small launch alone wants many small CTAs, overlapping launches want fewer,    */
                               /* larger ones.                                                                      */
    int variant;               /* 0 = automatic (grid size + concurrency); > 0 forces a forward kernel variant      */
                               /* (sweeps and parity tests; see launch_fwd_nhwc / launch_fwd_nchw).  NCHW: 1 = gather */
                               /* kernel, 2 = row segments staged in a double buffer (what automatic takes for        */
                               /* launches that fill the GPU), 3 / 4 = the same in a 3- / 4-deep ring (no gain).      */
    int nchw_cg;               /* NCHW kernels: channels in flight per lane / per CTA: 1,2,4,8,16 (0 = default)     */
    int bwd_mode;              /* backward: 0 = automatic; 2 = generic (non-packed) channels-last kernel;            */
                               /* NCHW: 4 = row segments + per-granule gather (= automatic), 1 = one reduction per  */
                               /* run of equal centres, 3 = one per tap (A/B measurements)                          */
    int nchw_tma;              /* NCHW forward: 0 gather through L1 (default); 1 = stage the patch footprint with    */
                               /* TMA box loads; 2..5 = same with a minimum box index (sweeps)                      */
    int zero_chunk_images;     /* backward with zero_fill: 0 (default) = one memset of the whole map, then one      */
                               /* scatter; k > 0 = clear k images at a time on a side stream and scatter each chunk  */
                               /* as soon as it is clear (measured slower on B200, kept for maps that thrash L2)     */
} rroi_b200_opts;

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

/*
 * The same three operations with per-call options (opts may be NULL = defaults = the entry points above).
 *   xform   optional [num_rois,8] table written by rroi_b200_roi_xform for THESE rois / pooled_height / scale
 *           (the forward then skips the per-RoI transform: fp64 divide, sinf, cosf, three fp32 divides); NULL = compute.
 */
int rroi_b200_forward_opt(const float* features, const float* rois, const float* xform, float* pooled,
                          float* idx_x, float* idx_y, int num_rois, int batch, int channels, int height, int width,
                          int pooled_height, int pooled_width, float spatial_scale, int layout,
                          const rroi_b200_opts* opts, cudaStream_t stream);
int rroi_b200_forward_bf16_opt(const void* features, const float* rois, const float* xform, void* pooled,
                               float* idx_x, float* idx_y, int num_rois, int batch, int channels, int height,
                               int width, int pooled_height, int pooled_width, float spatial_scale,
                               const rroi_b200_opts* opts, cudaStream_t stream);
int rroi_b200_backward_opt(const float* top_diff, const float* rois, const float* idx_x, const float* idx_y,
                           float* bottom_diff, int num_rois, int batch, int channels, int height, int width,
                           int pooled_height, int pooled_width, float spatial_scale, int layout, int zero_fill,
                           const rroi_b200_opts* opts, cudaStream_t stream);

/*
 * Per-RoI affine transform table (rroi_align_kernel.cu:58-84 evaluated once per RoI instead of once per output
 * element): xform [num_rois,8] = {M00, M01, M02, M10, M11, M12, roi_pooled_width, batch_idx (int bits)}.
 * Bit-identical to what the forward computes itself; a producer of RoI rows (fots_b200_boxes_to_rois) can emit
 * it so that the forward's CTAs start at the bin geometry.
 */
int rroi_b200_roi_xform(const float* rois, float* xform, int num_rois, int pooled_height, float spatial_scale,
                        cudaStream_t stream);

/* Expand compact centres [N,PH,PW] to the reference's [N,C,PH,PW] (ctx.idx_x / ctx.idx_y attribute parity). */
int rroi_b200_expand_idx(const float* idx_compact, float* idx_full, int num_rois, int channels,
                         int pooled_height, int pooled_width, cudaStream_t stream);

int         rroi_b200_last_cuda_error(void);            /* cudaError_t of the last failed launch (per process) */
const char* rroi_b200_strerror(int status);
int         rroi_b200_abi_version(void);                /* bumped on any signature change */
const char* rroi_b200_build_info(void);                 /* e.g. "sm_100a nvcc 12.9" */

#ifdef __cplusplus
}
#endif
#endif /* RROI_ALIGN_B200_H_ */
