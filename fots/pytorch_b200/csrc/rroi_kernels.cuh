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
    // fused zero-fill (channels-last packed kernel only; see rroi_bwd.cu "zero + scatter in one pass"): per-RoI rank
    // inside its image, per-image RoI count, per-image arrival counter, list of images without RoIs, {ok, n_empty}
    const int* zf_rank;
    const int* zf_count;
    int* zf_arrived;
    const int* zf_empty;
    const int* zf_meta;
    const int* zf_order;
    const int* zf_pos;
};

// Tunables a caller (bench sweeps, tests) may override through rroi_b200_set_tuning(); 0 = default.
struct Tuning {
    int nchw_cg;       // channels per CTA in the NCHW kernels: 1,2,4,8,16
    int nhwc_unroll;   // NHWC forward variant 0..5 (bins per warp x bins in flight), see launch_fwd_nhwc_vec
    int use_pdl;       // launch with programmatic stream serialization
    int bwd_dedupe;    // warp-level merge of equal sample points before the atomics (NCHW backward)
    int bwd_zero_fused; // 1 = one-pass zero + scatter backward for maps >= 96 MB (measured slower than memset + scatter: opt-in)
    int nchw_tma;      // 0 = NCHW forward gathers through L1 (default); 1 = stages its footprint with TMA box loads; 2..5 = same, box index >= value - 2
};
extern Tuning g_tuning;

cudaError_t launch_fwd_nchw(const FwdParams& p, cudaStream_t s);
cudaError_t launch_fwd_nhwc(const FwdParams& p, cudaStream_t s);
// bf16 features / bf16 pooled, channels-last, C in {32,64,128,256}; p.feat / p.out alias the bf16 buffers
cudaError_t launch_fwd_nhwc_bf16(const FwdParams& p, cudaStream_t s);
cudaError_t launch_bwd_nchw(const BwdParams& p, cudaStream_t s);
cudaError_t launch_bwd_nhwc(const BwdParams& p, cudaStream_t s);
// channels-last backward that also defines the whole gradient map (replaces cudaMemsetAsync + launch_bwd_nhwc when the
// map is much larger than L2); returns cudaErrorNotSupported when the shape is not eligible (caller falls back)
cudaError_t launch_bwd_nhwc_zero_fused(const BwdParams& p, cudaStream_t s);
// the reference-layout backward that honours caller-supplied [N,C,PH,PW] centres element by element
cudaError_t launch_bwd_legacy(const BwdParams& p, cudaStream_t s);

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
