// rroi_abi.cu -- the extern "C" boundary declared in include/rroi_align_b200.h.
// Argument checking and pointer/size plumbing only; the kernels live in rroi_fwd.cu / rroi_bwd.cu.
#include "../../../include/rroi_align_b200.h"
#include "rroi_kernels.cuh"

#include <atomic>
#include <cstring>

namespace {

std::atomic<int> g_last_cuda_error{0};

int cuda_status(cudaError_t e) {
    if (e == cudaSuccess) return RROI_B200_OK;
    g_last_cuda_error.store((int)e);
    (void)cudaGetLastError();   // clear the sticky launch error so the caller's next launch is not poisoned
    return RROI_B200_ERR_CUDA;
}

bool dims_ok(int n, int b, int c, int h, int w, int ph, int pw) {
    return n >= 0 && b > 0 && c > 0 && h > 0 && w > 0 && ph > 0 && pw > 0 &&
           (long long)ph * pw <= 0x7fffffffLL && (long long)b * h * w <= 0x7fffffffLL;
}

// rroi_b200_opts -> rroi::Opts.  false = malformed (unknown flag, value out of range).
bool parse_opts(const rroi_b200_opts* in, rroi::Opts* out) {
    *out = rroi::Opts();
    if (!in) return true;
    rroi_b200_opts o = {};
    const size_t have = in->size < sizeof(o) ? in->size : sizeof(o);
    if (have < 2 * sizeof(unsigned int)) return false;
    memcpy(&o, in, have);                                  // fields the caller's (older) header does not have read as 0
    if (o.flags & ~(RROI_B200_FLAG_NO_PDL | RROI_B200_FLAG_ROIS_READY)) return false;
    if (o.concurrency < 0 || o.variant < 0 || o.variant > 32) return false;
    if (o.nchw_cg != 0 && o.nchw_cg != 1 && o.nchw_cg != 2 && o.nchw_cg != 4 && o.nchw_cg != 8 && o.nchw_cg != 16) return false;
    if (o.bwd_mode < 0 || o.bwd_mode > 4 || o.nchw_tma < 0 || o.nchw_tma > 5 || o.zero_chunk_images < -1) return false;
    out->pdl = !(o.flags & RROI_B200_FLAG_NO_PDL);
    out->rois_ready = (o.flags & RROI_B200_FLAG_ROIS_READY) != 0 && out->pdl;
    out->concurrency = o.concurrency; out->variant = o.variant; out->nchw_cg = o.nchw_cg;
    out->bwd_mode = o.bwd_mode; out->nchw_tma = o.nchw_tma; out->zero_chunk_images = o.zero_chunk_images;
    return true;
}

__global__ void expand_idx_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t bins, int C, size_t total) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t bin = i % bins;
        const size_t n = i / (bins * (size_t)C);
        dst[i] = __ldg(src + n * bins + bin);
    }
}

}  // namespace

extern "C" {

int RROIAlignForwardLaucher(const float* bottom_data, const float spatial_scale, const int num_rois,
                            const int height, const int width, const int channels,
                            const int pooled_height, const int pooled_width, const float* bottom_rois,
                            float* top_data, float* con_idx_x, float* con_idx_y, cudaStream_t stream) {
    if (!bottom_data || !bottom_rois || !top_data || ((con_idx_x == nullptr) != (con_idx_y == nullptr))) return 0;
    if (!dims_ok(num_rois, 1, channels, height, width, pooled_height, pooled_width)) return 0;   // batch unknown here
    if (num_rois == 0) return 1;
    rroi::FwdParams p = {};
    p.feat = bottom_data; p.rois = bottom_rois; p.out = top_data; p.idx_x = con_idx_x; p.idx_y = con_idx_y;
    p.N = num_rois; p.B = 0x7fffffff;   // the legacy signature carries no batch size: trust the caller
    p.C = channels; p.H = height; p.W = width; p.PH = pooled_height; p.PW = pooled_width;
    p.scale = spatial_scale;
    p.idx_mode = con_idx_x ? rroi::IDX_FULL : rroi::IDX_NONE;
    const int st = cuda_status(rroi::launch_fwd_nchw(p, rroi::Opts(), stream));
    return st == RROI_B200_OK ? 1 : st;
}

int RROIAlignBackwardLaucher(const float* top_diff, const float spatial_scale, const int batch_size,
                             const int num_rois, const int height, const int width, const int channels,
                             const int pooled_height, const int pooled_width, const float* bottom_rois,
                             float* bottom_diff, const float* con_idx_x, const float* con_idx_y,
                             cudaStream_t stream) {
    if (!top_diff || !bottom_rois || !bottom_diff || !con_idx_x || !con_idx_y) return 0;
    if (!dims_ok(num_rois, batch_size, channels, height, width, pooled_height, pooled_width)) return 0;
    if (num_rois == 0) return 1;
    rroi::BwdParams p = {};
    p.top_diff = top_diff; p.rois = bottom_rois; p.bottom_diff = bottom_diff; p.idx_x = con_idx_x; p.idx_y = con_idx_y;
    p.N = num_rois; p.B = batch_size; p.C = channels; p.H = height; p.W = width;
    p.PH = pooled_height; p.PW = pooled_width; p.scale = spatial_scale; p.idx_mode = rroi::IDX_FULL;
    p.img_lo = -0x7fffffff; p.img_hi = 0x7fffffff;
    const int st = cuda_status(rroi::launch_bwd_legacy(p, rroi::Opts(), stream));
    return st == RROI_B200_OK ? 1 : st;
}

int rroi_b200_forward_opt(const float* features, const float* rois, const float* xform, float* pooled,
                          float* idx_x, float* idx_y, int num_rois, int batch, int channels, int height, int width,
                          int pooled_height, int pooled_width, float spatial_scale, int layout,
                          const rroi_b200_opts* opts, cudaStream_t stream) {
    rroi::Opts o;
    if (!parse_opts(opts, &o)) return RROI_B200_ERR_INVALID_ARG;
    if (!features || !pooled || (num_rois > 0 && !rois) || ((idx_x == nullptr) != (idx_y == nullptr)))
        return RROI_B200_ERR_INVALID_ARG;
    if (!dims_ok(num_rois, batch, channels, height, width, pooled_height, pooled_width)) return RROI_B200_ERR_INVALID_ARG;
    if (layout != RROI_B200_LAYOUT_NCHW && layout != RROI_B200_LAYOUT_NHWC) return RROI_B200_ERR_INVALID_ARG;
    if (xform && (reinterpret_cast<uintptr_t>(xform) & 15)) return RROI_B200_ERR_INVALID_ARG;
    if (num_rois == 0) return RROI_B200_OK;
    rroi::FwdParams p = {};
    p.feat = features; p.rois = rois; p.out = pooled; p.idx_x = idx_x; p.idx_y = idx_y;
    p.N = num_rois; p.B = batch; p.C = channels; p.H = height; p.W = width;
    p.PH = pooled_height; p.PW = pooled_width; p.scale = spatial_scale;
    p.idx_mode = idx_x ? rroi::IDX_COMPACT : rroi::IDX_NONE;
    p.xform = xform; p.early = o.rois_ready ? 1 : 0;
    const cudaError_t e = layout == RROI_B200_LAYOUT_NCHW ? rroi::launch_fwd_nchw(p, o, stream) : rroi::launch_fwd_nhwc(p, o, stream);
    if (e == cudaErrorInvalidConfiguration) { (void)cudaGetLastError(); return RROI_B200_ERR_TOO_LARGE; }
    if (e == cudaErrorInvalidValue && o.variant != 0) { (void)cudaGetLastError(); return RROI_B200_ERR_INVALID_ARG; }   // unknown variant
    return cuda_status(e);
}

int rroi_b200_forward(const float* features, const float* rois, float* pooled, float* idx_x, float* idx_y,
                      int num_rois, int batch, int channels, int height, int width,
                      int pooled_height, int pooled_width, float spatial_scale, int layout,
                      cudaStream_t stream) {
    return rroi_b200_forward_opt(features, rois, nullptr, pooled, idx_x, idx_y, num_rois, batch, channels, height, width,
                                 pooled_height, pooled_width, spatial_scale, layout, nullptr, stream);
}

int rroi_b200_forward_bf16_opt(const void* features, const float* rois, const float* xform, void* pooled,
                               float* idx_x, float* idx_y, int num_rois, int batch, int channels, int height,
                               int width, int pooled_height, int pooled_width, float spatial_scale,
                               const rroi_b200_opts* opts, cudaStream_t stream) {
    rroi::Opts o;
    if (!parse_opts(opts, &o)) return RROI_B200_ERR_INVALID_ARG;
    if (!features || !pooled || (num_rois > 0 && !rois) || ((idx_x == nullptr) != (idx_y == nullptr)))
        return RROI_B200_ERR_INVALID_ARG;
    if (!dims_ok(num_rois, batch, channels, height, width, pooled_height, pooled_width)) return RROI_B200_ERR_INVALID_ARG;
    if (channels != 32 && channels != 64 && channels != 128 && channels != 256) return RROI_B200_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(features) | reinterpret_cast<uintptr_t>(pooled)) & 15) return RROI_B200_ERR_INVALID_ARG;
    if (xform && (reinterpret_cast<uintptr_t>(xform) & 15)) return RROI_B200_ERR_INVALID_ARG;
    if (num_rois == 0) return RROI_B200_OK;
    rroi::FwdParams p = {};
    p.feat = static_cast<const float*>(features); p.rois = rois; p.out = static_cast<float*>(pooled);
    p.idx_x = idx_x; p.idx_y = idx_y;
    p.N = num_rois; p.B = batch; p.C = channels; p.H = height; p.W = width;
    p.PH = pooled_height; p.PW = pooled_width; p.scale = spatial_scale;
    p.idx_mode = idx_x ? rroi::IDX_COMPACT : rroi::IDX_NONE;
    p.xform = xform; p.early = o.rois_ready ? 1 : 0;
    const cudaError_t e = rroi::launch_fwd_nhwc_bf16(p, o, stream);
    if (e == cudaErrorInvalidConfiguration) { (void)cudaGetLastError(); return RROI_B200_ERR_TOO_LARGE; }
    return cuda_status(e);
}

int rroi_b200_forward_bf16(const void* features, const float* rois, void* pooled, float* idx_x, float* idx_y,
                           int num_rois, int batch, int channels, int height, int width,
                           int pooled_height, int pooled_width, float spatial_scale, cudaStream_t stream) {
    return rroi_b200_forward_bf16_opt(features, rois, nullptr, pooled, idx_x, idx_y, num_rois, batch, channels, height,
                                      width, pooled_height, pooled_width, spatial_scale, nullptr, stream);
}

int rroi_b200_backward_opt(const float* top_diff, const float* rois, const float* idx_x, const float* idx_y,
                           float* bottom_diff, int num_rois, int batch, int channels, int height, int width,
                           int pooled_height, int pooled_width, float spatial_scale, int layout, int zero_fill,
                           const rroi_b200_opts* opts, cudaStream_t stream) {
    rroi::Opts o;
    if (!parse_opts(opts, &o)) return RROI_B200_ERR_INVALID_ARG;
    if (!bottom_diff || (num_rois > 0 && (!rois || !top_diff)) || ((idx_x == nullptr) != (idx_y == nullptr)))
        return RROI_B200_ERR_INVALID_ARG;
    if (!dims_ok(num_rois, batch, channels, height, width, pooled_height, pooled_width)) return RROI_B200_ERR_INVALID_ARG;
    if (layout != RROI_B200_LAYOUT_NCHW && layout != RROI_B200_LAYOUT_NHWC) return RROI_B200_ERR_INVALID_ARG;
    rroi::BwdParams p = {};
    p.top_diff = top_diff; p.rois = rois; p.bottom_diff = bottom_diff; p.idx_x = idx_x; p.idx_y = idx_y;
    p.N = num_rois; p.B = batch; p.C = channels; p.H = height; p.W = width;
    p.PH = pooled_height; p.PW = pooled_width; p.scale = spatial_scale;
    p.idx_mode = idx_x ? rroi::IDX_COMPACT : rroi::IDX_NONE;
    p.img_lo = -0x7fffffff; p.img_hi = 0x7fffffff;
    const bool nhwc = layout == RROI_B200_LAYOUT_NHWC;
    cudaError_t e;
    if (zero_fill) e = rroi::launch_bwd_zero_scatter(p, o, nhwc, stream);
    else if (num_rois == 0) return RROI_B200_OK;
    else e = nhwc ? rroi::launch_bwd_nhwc(p, o, stream) : rroi::launch_bwd_nchw(p, o, stream);
    if (e == cudaErrorInvalidConfiguration) { (void)cudaGetLastError(); return RROI_B200_ERR_TOO_LARGE; }
    return cuda_status(e);
}

int rroi_b200_backward(const float* top_diff, const float* rois, const float* idx_x, const float* idx_y,
                       float* bottom_diff, int num_rois, int batch, int channels, int height, int width,
                       int pooled_height, int pooled_width, float spatial_scale, int layout,
                       int zero_fill, cudaStream_t stream) {
    return rroi_b200_backward_opt(top_diff, rois, idx_x, idx_y, bottom_diff, num_rois, batch, channels, height, width,
                                  pooled_height, pooled_width, spatial_scale, layout, zero_fill, nullptr, stream);
}

int rroi_b200_roi_xform(const float* rois, float* xform, int num_rois, int pooled_height, float spatial_scale,
                        cudaStream_t stream) {
    if (num_rois < 0 || pooled_height <= 0 || (num_rois > 0 && (!rois || !xform))) return RROI_B200_ERR_INVALID_ARG;
    if (reinterpret_cast<uintptr_t>(xform) & 15) return RROI_B200_ERR_INVALID_ARG;
    return cuda_status(rroi::launch_roi_xform(rois, xform, num_rois, pooled_height, spatial_scale, stream));
}

int rroi_b200_expand_idx(const float* idx_compact, float* idx_full, int num_rois, int channels,
                         int pooled_height, int pooled_width, cudaStream_t stream) {
    if (!idx_compact || !idx_full || num_rois < 0 || channels <= 0 || pooled_height <= 0 || pooled_width <= 0)
        return RROI_B200_ERR_INVALID_ARG;
    const size_t bins = (size_t)pooled_height * pooled_width;
    const size_t total = (size_t)num_rois * channels * bins;
    if (total == 0) return RROI_B200_OK;
    size_t grid = (total + 255) / 256;
    if (grid > 148u * 64u) grid = 148u * 64u;
    expand_idx_kernel<<<(unsigned)grid, 256, 0, stream>>>(idx_compact, idx_full, bins, channels, total);
    return cuda_status(cudaGetLastError());
}

int rroi_b200_last_cuda_error(void) { return g_last_cuda_error.load(); }

const char* rroi_b200_strerror(int status) {
    switch (status) {
        case RROI_B200_OK:              return "ok";
        case RROI_B200_ERR_INVALID_ARG: return "invalid argument (null pointer, non-positive size, unknown layout or key)";
        case RROI_B200_ERR_TOO_LARGE:   return "problem too large for one launch (more than 2^31-1 CTAs)";
        case RROI_B200_ERR_CUDA:        return "CUDA error at launch (see rroi_b200_last_cuda_error)";
        default:                        return "unknown status";
    }
}

int rroi_b200_abi_version(void) { return 2; }

const char* rroi_b200_build_info(void) {
#define RROI_STR2(x) #x
#define RROI_STR(x) RROI_STR2(x)
    return "librroi_b200 sm_100a nvcc " RROI_STR(__CUDACC_VER_MAJOR__) "." RROI_STR(__CUDACC_VER_MINOR__) "." RROI_STR(__CUDACC_VER_BUILD__);
}

}  // extern "C"
