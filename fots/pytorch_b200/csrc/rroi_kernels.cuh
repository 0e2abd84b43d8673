// rroi_kernels.cuh -- launch parameter block and host-side launch prototypes shared by the
// forward/backward translation units and the C-ABI (rroi_abi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rroi {

enum IdxMode : int {
    IDX_NONE = 0,     // do not write / read the sample centres
    IDX_COMPACT = 1,  // [N, PH, PW]   -- the centres do not depend on the channel
    IDX_FULL = 2      // [N, C, PH, PW] -- the reference's C-fold redundant layout (functions/rroi_align.py:19-20)
};

struct FwdParams {
    const float* feat;   // [B,C,H,W] (NCHW kernels) or [B,H,W,C] (NHWC kernels), fp32
    const float* rois;   // [N,6]
    float* out;          // [N,C,PH,PW] or [N,PH,PW,C]
    float* idx_x;        // see IdxMode; may be null when idx_mode == IDX_NONE
    float* idx_y;
    int N, B, C, H, W, PH, PW;
    float scale;
    int idx_mode;
    int tiles;           // bin tiles per (RoI[, channel group])
    int cgroups;         // channel groups per RoI (NCHW kernels)
    const float* xform;  // optional [N,8] RoiXform table (rroi_b200_roi_xform); null = compute from the RoI row
    int early;           // 1: the RoI rows / xform table may be read before griddepcontrol.wait (Opts::rois_ready)
};

struct BwdParams {
    const float* top_diff;  // [N,C,PH,PW] or [N,PH,PW,C]
    const float* rois;      // [N,6]
    float* bottom_diff;     // [B,C,H,W] or [B,H,W,C]; accumulated into (caller / ABI zero-fills)
    const float* idx_x;     // saved centres (IdxMode) or null -> recompute from the RoI
    const float* idx_y;
    int N, B, C, H, W, PH, PW;
    float scale;
    int idx_mode;
    int tiles;
    int cgroups;
    int img_lo, img_hi;     // only RoIs whose image index is in [img_lo, img_hi) scatter (chunked zero-fill + scatter)
};

// Per-call launch options (the C ABI's rroi_b200_opts, validated and with defaults filled in by rroi_abi.cu).
// There is no process-global tuning state: two threads / streams that want different variants cannot race.
struct Opts {
    bool pdl = true;         // launch with programmatic stream serialization
    bool rois_ready = false; // RROI_B200_FLAG_ROIS_READY: prologue (RoI rows -> transform -> geometry) may run before griddepcontrol.wait
    int concurrency = 0;     // independent launches the caller keeps in flight (0/1 = alone)
    int variant = 0;         // 0 = automatic; > 0 forces a forward kernel variant
    int nchw_cg = 0;         // channels per lane / per CTA in the NCHW kernels: 1,2,4,8,16 (0 = default)
    int bwd_mode = 0;        // 0 auto; 1 per-tap reductions (NCHW: warp-merged runs); 2 generic channels-last kernel; 3 NCHW without the warp merge; 4 NCHW row segments + gather (= auto)
    int nchw_tma = 0;        // 0 = NCHW forward gathers through L1; 1 = TMA box staging; 2..5 = same, box index >= value - 2
    int zero_chunk_images = 0; // backward zero_fill: images per zero/scatter chunk (0 auto, -1 whole map)
};

cudaError_t launch_fwd_nchw(const FwdParams& p, const Opts& o, cudaStream_t s);
cudaError_t launch_fwd_nhwc(const FwdParams& p, const Opts& o, cudaStream_t s);
// bf16 features / bf16 pooled, channels-last, C in {32,64,128,256}; p.feat / p.out alias the bf16 buffers
cudaError_t launch_fwd_nhwc_bf16(const FwdParams& p, const Opts& o, cudaStream_t s);
cudaError_t launch_bwd_nchw(const BwdParams& p, const Opts& o, cudaStream_t s);
cudaError_t launch_bwd_nhwc(const BwdParams& p, const Opts& o, cudaStream_t s);
// zero-fill of the whole gradient map + scatter, chunked by image and overlapped on a side stream when the map is much
// larger than L2 (rroi_bwd.cu); plain cudaMemsetAsync + one scatter otherwise
cudaError_t launch_bwd_zero_scatter(const BwdParams& p, const Opts& o, bool nhwc, cudaStream_t s);
// the reference-layout backward that honours caller-supplied [N,C,PH,PW] centres element by element
cudaError_t launch_bwd_legacy(const BwdParams& p, const Opts& o, cudaStream_t s);
// [N,8] transform table (rroi_b200_roi_xform)
cudaError_t launch_roi_xform(const float* rois, float* xform, int N, int PH, float scale, cudaStream_t s);

template <typename K, typename P>
inline cudaError_t launch_1d(K kernel, long long grid, int block, const P& p, cudaStream_t s, bool pdl) {
    if (grid <= 0) return cudaSuccess;
    if (grid > 2147483647LL) return cudaErrorInvalidConfiguration;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, p);
}

// Programmatic dependent launch (sm_90+): let the next kernel in the stream start its prologue
// while this one drains, and do not touch global memory before the previous one has flushed.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

}  // namespace rroi
