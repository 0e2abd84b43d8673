// rroi_fwd.cu -- fused RoIRotate forward for sm_100a: per-RoI affine parameters + bilinear gather in
// one kernel.  Replaces RROIAlignForward (/root/reference/rroi_align/src/rroi_align_kernel.cu:28-162)
// and the three zero-fills its caller needs (rroi_align/functions/rroi_align.py:17-20).
//
// What changed relative to the reference's one-thread-per-output-element kernel:
//   * the RoI transform (fp64 divide, sinf, cosf, 3 fp32 divides) is computed once per CTA, not once
//     per output element (C*PH*PW-fold redundancy removed);
//   * the bin geometry (corner projection, round, clamp, weights) is computed once per (n,ph,pw) and
//     reused for every channel;
//   * outputs are written with plain stores, zero tail included, so nothing has to be pre-zeroed and
//     the three float atomics per element are gone;
//   * the sample centres can be kept compact ([N,PH,PW]) instead of C-fold replicated;
//   * two data layouts: reference NCHW (drop-in) and channels-last, where one tap of one bin is a
//     contiguous C*4-byte vector and every global access is a full 128-bit coalesced transaction.
// HBM-bound gather: no tensor cores, no shared-memory staging needed for the features (each tap is
// read once per CTA; reuse between neighbouring bins is served by L1/L2).
#include "rroi_geom.cuh"
#include "rroi_kernels.cuh"
#include <cuda.h>
#include <climits>
#include <mutex>

namespace rroi {

// The per-RoI transform: from the caller's table (rroi_b200_roi_xform) when there is one, else computed here.
__device__ __forceinline__ RoiXform get_xform(const FwdParams& p, int n) {
    if (p.xform) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p.xform) + 2 * (size_t)n);
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.xform) + 2 * (size_t)n + 1);
        RoiXform X;
        X.M00 = a.x; X.M01 = a.y; X.M02 = a.z; X.M10 = a.w; X.M11 = b.x; X.M12 = b.y; X.rpw = b.z; X.batch = __float_as_int(b.w);
        return X;
    }
    return roi_xform(p.rois + (size_t)n * 6, p.scale, p.PH);
}

// CTA prologue of the block-level kernels: warp 0 fetches the transform into shared memory.  With p.early (the caller
// vouches that the RoI rows are older than the preceding kernel) that happens before the grid dependency resolves.
__device__ __forceinline__ void cta_xform_prologue(const FwdParams& p, int n, RoiXform* sX) {
    if (!p.early) pdl_wait();
    pdl_launch_dependents();
    if (threadIdx.x < 32) {
        const RoiXform X = get_xform(p, n);
        if (threadIdx.x == 0) *sX = X;
    }
    if (p.early) pdl_wait();
    __syncthreads();
}

__global__ void __launch_bounds__(128) roi_xform_kernel(const float* __restrict__ rois, float* __restrict__ xform, int N, int PH, float scale) {
    pdl_wait();
    pdl_launch_dependents();
    const int n = blockIdx.x * 128 + threadIdx.x;
    if (n >= N) return;
    const RoiXform X = roi_xform(rois + (size_t)n * 6, scale, PH);
    float4* o = reinterpret_cast<float4*>(xform) + 2 * (size_t)n;
    o[0] = make_float4(X.M00, X.M01, X.M02, X.M10);
    o[1] = make_float4(X.M11, X.M12, X.rpw, __int_as_float(X.batch));
}

cudaError_t launch_roi_xform(const float* rois, float* xform, int N, int PH, float scale, cudaStream_t s) {
    if (N <= 0) return cudaSuccess;
    roi_xform_kernel<<<(N + 127) / 128, 128, 0, s>>>(rois, xform, N, PH, scale);
    return cudaGetLastError();
}

// code word of a bin
enum : uint32_t {
    C_LT = 1u, C_RT = 2u, C_LB = 4u, C_RB = 8u,  // tap must be loaded (valid, distinct)
    C_HX = 16u, C_HY = 32u,                       // rx == 0.5 / ry == 0.5
    C_IN = 64u,                                   // bin inside the RoI
    C_NAN = 128u,                                 // centre not finite: the reference's weights are NaN
    C_LIVE = 256u                                 // bin index < PH*PW
};

// ------------------------------------------------------------------------------------------ NCHW
// Reference layout: feat [B,C,H,W], out [N,C,PH,PW].  A tap is one 4-byte word of plane (b,c); the 32
// lanes of a warp are 32 BINS and the warp walks the channels, so stores are coalesced along pw and the
// loads are a 32-lane gather inside one plane.
//
// grid = N * tiles; CTA = 8 warps = one RoI x an 8x8 patch of bins (8 ph x 8 pw) x all channels.  The 2-D
// patch keeps the CTA's footprint in the feature plane compact for any rotation, so the 32-byte sectors
// a gather touches are reused out of L1 by the neighbouring bins of the same CTA instead of being
// re-fetched from L2 by another CTA.  Warp w owns half of the patch (4 ph x 8 pw: four full 32-byte
// sectors per store) and every 4th channel; its lanes keep their bin's record in registers.
// Bin geometry: one thread per bin, once per CTA (the reference recomputes it per channel).
constexpr int kNchwBlock = 256;
constexpr int kPatch = 8;                 // patch edge in bins

// 32-bit load under a predicate, zero otherwise (not .nc: see ldg_pred_v4)
template <int kByteOff>
__device__ __forceinline__ float ldg_pred_f32(const float* ptr, uint32_t pred) {
    float r;
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\tmov.b32 %0, 0;\n\t@q ld.global.f32 %0, [%1+%3];\n\t}"
        : "=f"(r) : "l"(ptr), "r"(pred), "n"(kByteOff));
    return r;
}

struct __align__(16) BinRecP {
    int off;                       // t*W + l inside the plane
    uint32_t code;                 // C_LT|C_RT|C_LB|C_RB load bits, C_LIVE
    float wlt, wrt, wrb, wlb;
    float cx, cy;                  // for the reference's [N,C,PH,PW] centre tensors (IDX_FULL)
};

template <int UN>
__global__ void __launch_bounds__(kNchwBlock) rroi_fwd_nchw_kernel(const FwdParams p) {
    __shared__ RoiXform sX;
    __shared__ BinRecP rec[kPatch * kPatch];
    const int n = blockIdx.x / p.tiles;
    const int tile = blockIdx.x - n * p.tiles;
    const int tiles_w = (p.PW + kPatch - 1) / kPatch;
    const int ph0 = (tile / tiles_w) * kPatch, pw0 = (tile % tiles_w) * kPatch;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bins = p.PH * p.PW;

    cta_xform_prologue(p, n, &sX);
    const RoiXform X = sX;
    const bool batch_ok = (X.batch >= 0) & (X.batch < p.B);
    if (threadIdx.x < kPatch * kPatch) {
        BinRecP r;
        r.off = 0; r.code = 0; r.wlt = r.wrt = r.wrb = r.wlb = 0.0f; r.cx = r.cy = 0.0f;
        const int ph = ph0 + (int)threadIdx.x / kPatch, pw = pw0 + (int)threadIdx.x % kPatch;
        if (ph < p.PH && pw < p.PW) {
            const BinTaps g = bin_taps(X, ph, pw, p.H, p.W, (float)(p.W - 1), (float)(p.H - 1), batch_ok);
            const bool in = g.flags & BIN_IN, two_c = g.flags & TWO_COLS, two_r = g.flags & TWO_ROWS;
            r.off = (int)((unsigned)g.t * (unsigned)p.W + (unsigned)g.l);
            r.code = C_LIVE;
            if (in) {
                const bool nanw = !(fabsf(g.cx) < INFINITY) || !(fabsf(g.cy) < INFINITY);
                const bool l_lt = g.flags & TAP_LT;
                const bool l_rt = (g.flags & TAP_RT) && two_c;
                const bool l_lb = (g.flags & TAP_LB) && two_r;
                const bool l_rb = (g.flags & TAP_RB) && two_c && two_r;
                r.code |= (l_lt ? C_LT : 0u) | (l_rt ? C_RT : 0u) | (l_lb ? C_LB : 0u) | (l_rb ? C_RB : 0u);
                r.wlt = (l_lt || nanw) ? g.wlt : 0.0f;
                r.wrt = (l_rt || nanw) ? g.wrt : 0.0f;
                r.wrb = (l_rb || nanw) ? g.wrb : 0.0f;
                r.wlb = (l_lb || nanw) ? g.wlb : 0.0f;
                r.cx = g.cx; r.cy = g.cy;
            }
            if (p.idx_mode == IDX_COMPACT) {
                p.idx_x[(size_t)n * bins + ph * p.PW + pw] = r.cx;
                p.idx_y[(size_t)n * bins + ph * p.PW + pw] = r.cy;
            }
        }
        rec[threadIdx.x] = r;
    }
    __syncthreads();

    // lane -> bin of this warp's half patch; warp -> channel phase
    const int half = warp & 1, cq = warp >> 1;                  // 2 halves x 4 channel phases
    const int slot = half * 32 + lane;                          // row-major inside the 8x8 patch
    const BinRecP r = rec[slot];
    if (!(r.code & C_LIVE)) return;
    const int ph = ph0 + slot / kPatch, pw = pw0 + slot % kPatch;
    const size_t HW = (size_t)p.H * p.W;
    const float* top = p.feat + ((size_t)(batch_ok ? X.batch : 0) * p.C) * HW + r.off;   // lt; rt = top + 1
    const float* bot = top + p.W;                                                         // lb; rb = bot + 1
    const size_t obin = (size_t)ph * p.PW + pw;
    float* dst = p.out + (size_t)n * p.C * bins + obin;
    const bool full_idx = p.idx_mode == IDX_FULL;

#pragma unroll 1
    for (int c0 = cq; c0 < p.C; c0 += 4 * UN) {
        float lt[UN], rt[UN], lb[UN], rb[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int c = c0 + 4 * u;
            const uint32_t ok = c < p.C;
            const size_t po = (size_t)c * HW;
            lt[u] = ldg_pred_f32<0>(top + po, ok ? (r.code & C_LT) : 0u);
            rt[u] = ldg_pred_f32<4>(top + po, ok ? (r.code & C_RT) : 0u);
            lb[u] = ldg_pred_f32<0>(bot + po, ok ? (r.code & C_LB) : 0u);
            rb[u] = ldg_pred_f32<4>(bot + po, ok ? (r.code & C_RB) : 0u);
        }
        if (r.code & C_LIVE) {     // own basic block: keeps the UN channels' loads batched (see NHWC kernel)
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int c = c0 + 4 * u;
                if (c < p.C) {
                    float v = __fmaf_rn(lt[u], r.wlt, 0.0f);
                    v = __fmaf_rn(rt[u], r.wrt, v);
                    v = __fmaf_rn(r.wrb, rb[u], v);
                    v = __fmaf_rn(lb[u], r.wlb, v);
                    dst[(size_t)c * bins] = v;
                    if (full_idx) {
                        p.idx_x[((size_t)n * p.C + c) * bins + obin] = r.cx;
                        p.idx_y[((size_t)n * p.C + c) * bins + obin] = r.cy;
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------ NCHW, TMA-staged
// The gather kernel above is bound by L1 wavefronts: a warp-wide tap load of 32 bins of one plane touches ~8 different
// 128-byte lines (ncu: 7.8 sectors per request, L1 throughput 76 %).  Here the feature data does not go through the
// load/store unit at all: for every group of CG channels the CTA's footprint in the planes -- the bounding box of
// the taps of its 8x8 bin patch, at most 32x32 pixels -- is fetched by ONE 3-D TMA box load {BX, BY, CG planes}
// (cp.async.bulk.tensor.3d, out-of-image parts zero-filled) into shared memory, double-buffered over the channel
// groups, and the four taps of every bin are read from there.  Four tensor maps with boxes 12x8x32, 20x16x16, 28x24x8
// and 36x32x4 (12-21 KB per stage; four pixels wider than tall because the box must start on a 16-byte boundary of the
// row) are encoded per call; each CTA picks the smallest that holds its footprint.  CTAs
// whose footprint is larger (huge RoIs) or empty fall back to the gather loop.  Same records, same FFMA chain, same
// stores as the gather kernel: bit-identical results.
struct NchwTmaMaps { CUtensorMap m[4]; };
constexpr int kTmaStageBytes = 28 * 24 * 8 * 4;          // the largest of the four boxes (21 504 B)

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(kNchwBlock)
rroi_fwd_nchw_tma_kernel(const FwdParams p, const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                         const __grid_constant__ CUtensorMap map2, const __grid_constant__ CUtensorMap map3) {
    extern __shared__ __align__(128) uint8_t stage_mem[];    // 2 x kTmaStageBytes
    __shared__ RoiXform sX;
    __shared__ BinRecP rec[kPatch * kPatch];
    __shared__ int sL[kPatch * kPatch], sT[kPatch * kPatch];
    __shared__ int bb[4];                                   // x_min, y_min, x_max, y_max of the patch's taps
    __shared__ __align__(8) uint64_t full_bar[2];
    const int n = blockIdx.x / p.tiles;
    const int tile = blockIdx.x - n * p.tiles;
    const int tiles_w = (p.PW + kPatch - 1) / kPatch;
    const int ph0 = (tile / tiles_w) * kPatch, pw0 = (tile % tiles_w) * kPatch;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bins = p.PH * p.PW;

    if (!p.early) pdl_wait();
    pdl_launch_dependents();
    if (warp == 0) {
        const RoiXform X = get_xform(p, n);
        if (p.early) pdl_wait();
        if (lane == 0) {
            sX = X;
            bb[0] = INT_MAX; bb[1] = INT_MAX; bb[2] = INT_MIN; bb[3] = INT_MIN;
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&full_bar[0])) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&full_bar[1])) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    __syncthreads();
    const RoiXform X = sX;
    const bool batch_ok = (X.batch >= 0) & (X.batch < p.B);
    if (threadIdx.x < kPatch * kPatch) {
        BinRecP r;
        r.off = 0; r.code = 0; r.wlt = r.wrt = r.wrb = r.wlb = 0.0f; r.cx = r.cy = 0.0f;
        int xl = INT_MAX, yt = INT_MAX, xr = INT_MIN, yb = INT_MIN, gl = 0, gt = 0;
        const int ph = ph0 + (int)threadIdx.x / kPatch, pw = pw0 + (int)threadIdx.x % kPatch;
        if (ph < p.PH && pw < p.PW) {
            const BinTaps g = bin_taps(X, ph, pw, p.H, p.W, (float)(p.W - 1), (float)(p.H - 1), batch_ok);
            const bool in = g.flags & BIN_IN, two_c = g.flags & TWO_COLS, two_r = g.flags & TWO_ROWS;
            r.off = (int)((unsigned)g.t * (unsigned)p.W + (unsigned)g.l);
            r.code = C_LIVE;
            if (in) {
                const bool nanw = !(fabsf(g.cx) < INFINITY) || !(fabsf(g.cy) < INFINITY);
                const bool l_lt = g.flags & TAP_LT;
                const bool l_rt = (g.flags & TAP_RT) && two_c;
                const bool l_lb = (g.flags & TAP_LB) && two_r;
                const bool l_rb = (g.flags & TAP_RB) && two_c && two_r;
                r.code |= (l_lt ? C_LT : 0u) | (l_rt ? C_RT : 0u) | (l_lb ? C_LB : 0u) | (l_rb ? C_RB : 0u);
                r.wlt = (l_lt || nanw) ? g.wlt : 0.0f;
                r.wrt = (l_rt || nanw) ? g.wrt : 0.0f;
                r.wrb = (l_rb || nanw) ? g.wrb : 0.0f;
                r.wlb = (l_lb || nanw) ? g.wlb : 0.0f;
                r.cx = g.cx; r.cy = g.cy;
                if (r.code & (C_LT | C_RT | C_LB | C_RB)) {       // loaded taps are inside the image: 0 <= l,t and l+1 < W, t+1 < H when used
                    gl = g.l; gt = g.t;
                    xl = g.l; yt = g.t; xr = g.l + (two_c ? 1 : 0); yb = g.t + (two_r ? 1 : 0);
                }
            }
            if (p.idx_mode == IDX_COMPACT) {
                p.idx_x[(size_t)n * bins + ph * p.PW + pw] = r.cx;
                p.idx_y[(size_t)n * bins + ph * p.PW + pw] = r.cy;
            }
        }
        rec[threadIdx.x] = r;
        sL[threadIdx.x] = gl; sT[threadIdx.x] = gt;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            xl = min(xl, __shfl_xor_sync(0xffffffffu, xl, m)); yt = min(yt, __shfl_xor_sync(0xffffffffu, yt, m));
            xr = max(xr, __shfl_xor_sync(0xffffffffu, xr, m)); yb = max(yb, __shfl_xor_sync(0xffffffffu, yb, m));
        }
        if (lane == 0) { atomicMin(&bb[0], xl); atomicMin(&bb[1], yt); atomicMax(&bb[2], xr); atomicMax(&bb[3], yb); }
    }
    __syncthreads();
    const bool any_tap = bb[2] >= bb[0];                    // false when the patch loads nothing
    // TMA wants the box to start on a 16-byte boundary of the innermost dimension: x0 is rounded down to a multiple of
    // four pixels and every box is four pixels wider than it is tall.
    const int x0 = any_tap ? (bb[0] & ~3) : 0, y0 = any_tap ? bb[1] : 0;
    const int need_w = any_tap ? bb[2] - x0 + 1 : 0, need_h = any_tap ? bb[3] - y0 + 1 : 0;
    const int side = max(need_w - 4, need_h);
    // box index: 0: 12x8x32, 1: 20x16x16, 2: 28x24x8, 3: 36x32x4  (width x height x channels)
    int k = side <= 8 ? 0 : side <= 16 ? 1 : side <= 24 ? 2 : 3;
    if (p.cgroups >= 2 && p.cgroups - 2 >= k) k = p.cgroups - 2;       // tuning/debug: force a (larger) box
    const bool use_tma = any_tap && side <= 32;
    const int BY = 8 * (k + 1), BX = BY + 4;
    const int CG = k == 0 ? 32 : k == 1 ? 16 : k == 2 ? 8 : 4;

    // lane -> bin of this warp's half patch; warp -> channel phase (as in the gather kernel)
    const int half = warp & 1, cq = warp >> 1;
    const int slot = half * 32 + lane;
    const BinRecP r = rec[slot];
    const bool live = r.code & C_LIVE;
    const int ph = ph0 + slot / kPatch, pw = pw0 + slot % kPatch;
    const size_t HW = (size_t)p.H * p.W;
    const size_t obin = (size_t)ph * p.PW + pw;
    float* dst = p.out + (size_t)n * p.C * bins + obin;
    const bool full_idx = p.idx_mode == IDX_FULL;

    if (use_tma) {
        const int soff = (sT[slot] - y0) * BX + (sL[slot] - x0);         // top-left tap inside one staged plane
        const int plane_elems = BX * BY;
        const uint32_t stage_bytes = (uint32_t)(plane_elems * CG * 4);
        const int nst = (p.C + CG - 1) / CG;
        const int plane0 = (batch_ok ? X.batch : 0) * p.C;
        auto tma3d = [&](const CUtensorMap* map, uint32_t dsts, uint32_t bar, int cz) {
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(dsts), "l"(map), "r"(bar), "r"(x0), "r"(y0), "r"(cz) : "memory");
        };
        auto issue = [&](int st) {
            const uint32_t bar = smem_addr(&full_bar[st & 1]);
            const uint32_t dsts = smem_addr(stage_mem + (st & 1) * kTmaStageBytes);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(stage_bytes) : "memory");
            const int cz = plane0 + st * CG;
            // each tensor map is addressed as the kernel parameter it is (no run-time indexing of parameter space)
            if (k == 0) tma3d(&map0, dsts, bar, cz);
            else if (k == 1) tma3d(&map1, dsts, bar, cz);
            else if (k == 2) tma3d(&map2, dsts, bar, cz);
            else tma3d(&map3, dsts, bar, cz);
        };
        if (threadIdx.x == 0) {
            issue(0);
            if (nst > 1) issue(1);
        }
        for (int st = 0; st < nst; ++st) {
            const uint32_t bar = smem_addr(&full_bar[st & 1]);
            const uint32_t parity = (uint32_t)(st >> 1) & 1u;
            uint32_t ok = 0;
            while (!ok) {
                asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                             : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
            }
            const float* tl = reinterpret_cast<const float*>(stage_mem + (st & 1) * kTmaStageBytes) + soff;
            if (live) {
#pragma unroll 4
                for (int cl = cq; cl < CG; cl += 4) {
                    const int c = st * CG + cl;
                    if (c >= p.C) break;
                    const float* s0 = tl + cl * plane_elems;
                    const float lt = (r.code & C_LT) ? s0[0] : 0.0f;
                    const float rt = (r.code & C_RT) ? s0[1] : 0.0f;
                    const float lb = (r.code & C_LB) ? s0[BX] : 0.0f;
                    const float rb = (r.code & C_RB) ? s0[BX + 1] : 0.0f;
                    float v = __fmaf_rn(lt, r.wlt, 0.0f);
                    v = __fmaf_rn(rt, r.wrt, v);
                    v = __fmaf_rn(r.wrb, rb, v);
                    v = __fmaf_rn(lb, r.wlb, v);
                    dst[(size_t)c * bins] = v;
                    if (full_idx) {
                        p.idx_x[((size_t)n * p.C + c) * bins + obin] = r.cx;
                        p.idx_y[((size_t)n * p.C + c) * bins + obin] = r.cy;
                    }
                }
            }
            __syncthreads();                                  // everybody is done with this stage's buffer
            if (threadIdx.x == 0 && st + 2 < nst) issue(st + 2);
        }
        return;
    }

    // ---- fallback: footprint larger than 32x32 pixels, or nothing to load -- the gather loop ----
    if (!live) return;
    const float* top = p.feat + ((size_t)(batch_ok ? X.batch : 0) * p.C) * HW + r.off;
    const float* bot = top + p.W;
    constexpr int UN = 4;
#pragma unroll 1
    for (int c0 = cq; c0 < p.C; c0 += 4 * UN) {
        float lt[UN], rt[UN], lb[UN], rb[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int c = c0 + 4 * u;
            const uint32_t okc = c < p.C;
            const size_t po = (size_t)c * HW;
            lt[u] = ldg_pred_f32<0>(top + po, okc ? (r.code & C_LT) : 0u);
            rt[u] = ldg_pred_f32<4>(top + po, okc ? (r.code & C_RT) : 0u);
            lb[u] = ldg_pred_f32<0>(bot + po, okc ? (r.code & C_LB) : 0u);
            rb[u] = ldg_pred_f32<4>(bot + po, okc ? (r.code & C_RB) : 0u);
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int c = c0 + 4 * u;
            if (c < p.C) {
                float v = __fmaf_rn(lt[u], r.wlt, 0.0f);
                v = __fmaf_rn(rt[u], r.wrt, v);
                v = __fmaf_rn(r.wrb, rb[u], v);
                v = __fmaf_rn(lb[u], r.wlb, v);
                dst[(size_t)c * bins] = v;
                if (full_idx) {
                    p.idx_x[((size_t)n * p.C + c) * bins + obin] = r.cx;
                    p.idx_y[((size_t)n * p.C + c) * bins + obin] = r.cy;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------- NCHW, row segments staged
// The bounding-box staging above over-fetches 3-4x (a rotated footprint fills half of its box, plus alignment slack and
// box-size quantisation); the gather kernel is bound by L1 wavefronts (7.8 sectors per 32-lane request).  This kernel
// stages exactly the ROW SEGMENTS the tile touches: CTA = one RoI x an 8 (ph) x 32 (pw) block of bins, thread = bin.
//   1. every thread computes its bin's geometry (registers);
//   2. the CTA builds, once, the list of 16-byte granules (4 pixels) of its footprint: per image row the span
//      [min x, max x] of the taps that are loaded, widened to granule boundaries (shared-memory atomicMin/Max, a warp
//      scan for the offsets, a table plane-offset-of-granule);
//   3. for every group of channels the granules are copied plane by plane with cp.async (16 bytes per thread, consecutive
//      threads = consecutive granules of a row: fully coalesced, every fetched byte within one granule of a tap),
//      double-buffered over the channel groups;
//   4. the blend reads its four taps from shared memory (predicated: an unloaded tap is 0, as in the gather kernel),
//      same 4-FFMA chain, and stores 32 consecutive pw per warp (128-byte coalesced).
// Tiles whose footprint does not fit (more than 96 rows or 20 KB per channel) or planes whose rows are not 16-byte
// aligned take the gather loop inside the same kernel.  Bit-identical to the gather kernel (and to the reference kernel, see tests/).
// Template: S stages of SF floats each (dynamic shared memory); RING = S-deep ring with ONE barrier per stage (stage
// st + S - 1 is issued right after the barrier that publishes stage st) instead of the two-barrier double buffer.
constexpr int kRowsMax = 96;                      // image rows a tile's footprint may span

__device__ __forceinline__ void cp_async_16(uint32_t dst_smem, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct RowsStageArgs {
    float* stage0;            // S stages of SF floats
    const int* gsrc;          // plane offset (floats) of every granule of the footprint
    const float* plane0;      // channel 0 of the RoI's image
    float* dst;               // this thread's bin in channel 0 of the RoI's output
    size_t HW, bins;
    int C, gtot;
    bool live;
    int o_lt, o_rt, o_lb, o_rb;       // float offsets of the four taps inside a channel slot (gtot * 4 = the zero granule)
    float wlt, wrt, wrb, wlb;
};

// The staged channel loop for CGS channels per stage.  A stage is CGS slots of SF / CGS floats, so every shared-memory
// address of the blend is (per-thread tap address of the stage) + (compile-time channel offset): LDS with an immediate,
// no address arithmetic per channel.  Copies: warp w owns channel slot w % CGS and, with the 8 / CGS - 1 other warps of
// that slot, walks the slot's granules 32 * (8 / CGS) apart -- consecutive lanes copy consecutive granules of one plane
// (coalesced), the plane offsets of a thread's <= KP granules are loaded once, the destination of granule k is the
// thread's first destination + a compile-time step, and the source plane advances by CGS planes per stage.
template <int S, int SF, bool RING, int CGS>
__device__ __forceinline__ void rows_stage_loop(const RowsStageArgs& a) {
    constexpr int STRIDE = SF / CGS;                          // floats per channel slot
    constexpr int WPC = 8 / CGS;                              // warps that share a slot
    constexpr int KP = (SF / 4 - CGS + 255) / 256;            // granules per thread: ceil((STRIDE / 4 - 1) / (32 * WPC))
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nst = (a.C + CGS - 1) / CGS;
    for (int i = threadIdx.x; i < S * CGS * 4; i += 256)      // the zero granule of every slot of every stage
        a.stage0[(i >> 2) / CGS * SF + ((i >> 2) % CGS) * STRIDE + a.gtot * 4 + (i & 3)] = 0.0f;
    const int ch = warp % CGS, g0 = (warp / CGS) * 32 + lane;
    uint32_t soff[KP];
    int nk = 0;                                               // granules this thread copies per stage
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        const int gi = g0 + k * 32 * WPC;
        soff[k] = gi < a.gtot ? (uint32_t)a.gsrc[gi] : 0u;
        nk += gi < a.gtot;
    }
    const uint32_t sbase = smem_addr(a.stage0);
    const uint32_t cdst = sbase + (uint32_t)(ch * STRIDE + g0 * 4) * 4u;
    const float* csrc = a.plane0 + (size_t)ch * a.HW;         // this warp's plane of stage 0; + CGS planes per issued stage
    int issued = 0;
    auto issue = [&]() {                                      // stages are issued in order: stage `issued` into buffer issued % S
        if (issued * CGS + ch < a.C) {                        // warp-uniform (the last stage may hold fewer channels)
            const uint32_t db = cdst + (uint32_t)(issued % S) * (uint32_t)(SF * 4);
#pragma unroll
            for (int k = 0; k < KP; ++k)
                if (k < nk) cp_async_16(db + k * (32 * WPC * 16), csrc + soff[k]);
        }
        csrc += (size_t)CGS * a.HW;
        ++issued;
    };
    float* d = a.dst;
    auto blend = [&](int st) {
        const float* sb = a.stage0 + (st % S) * SF;
        const float* q_lt = sb + a.o_lt; const float* q_rt = sb + a.o_rt;
        const float* q_rb = sb + a.o_rb; const float* q_lb = sb + a.o_lb;
        const int cn = min(CGS, a.C - st * CGS);
        if (a.live) {
            if (cn == CGS) {
                float x[CGS][4];
#pragma unroll
                for (int ci = 0; ci < CGS; ++ci) {
                    x[ci][0] = q_lt[ci * STRIDE]; x[ci][1] = q_rt[ci * STRIDE];
                    x[ci][2] = q_rb[ci * STRIDE]; x[ci][3] = q_lb[ci * STRIDE];
                }
#pragma unroll
                for (int ci = 0; ci < CGS; ++ci) {
                    float v = __fmaf_rn(x[ci][0], a.wlt, 0.0f);
                    v = __fmaf_rn(x[ci][1], a.wrt, v);
                    v = __fmaf_rn(a.wrb, x[ci][2], v);
                    v = __fmaf_rn(x[ci][3], a.wlb, v);
                    d[ci * a.bins] = v;
                }
            } else {
                for (int ci = 0; ci < cn; ++ci) {
                    float v = __fmaf_rn(q_lt[ci * STRIDE], a.wlt, 0.0f);
                    v = __fmaf_rn(q_rt[ci * STRIDE], a.wrt, v);
                    v = __fmaf_rn(a.wrb, q_rb[ci * STRIDE], v);
                    v = __fmaf_rn(q_lb[ci * STRIDE], a.wlb, v);
                    d[ci * a.bins] = v;
                }
            }
        }
        d += (size_t)cn * a.bins;
    };
    __syncthreads();                                          // zero granules written before anybody blends
    if (RING) {
        // S-deep ring, one barrier per stage: the barrier of stage st says "everybody's copies of stage st have landed"
        // AND "everybody has finished blending stage st - 1", so buffer (st - 1) % S = (st + S - 1) % S may be refilled.
#pragma unroll
        for (int k = 0; k < S - 1; ++k) {
            if (k < nst) issue();
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        for (int st = 0; st < nst; ++st) {
            cp_async_wait<S - 2>();
            __syncthreads();
            if (st + S - 1 < nst) issue();
            asm volatile("cp.async.commit_group;" ::: "memory");   // (possibly empty: keeps the group count uniform)
            blend(st);
        }
    } else {
        issue();
        asm volatile("cp.async.commit_group;" ::: "memory");
        for (int st = 0; st < nst; ++st) {
            if (st + 1 < nst) { issue(); asm volatile("cp.async.commit_group;" ::: "memory"); cp_async_wait<1>(); }
            else cp_async_wait<0>();
            __syncthreads();                                  // everybody's copies of this stage have landed
            blend(st);
            __syncthreads();                                  // the buffer may be refilled (stage st + 2)
        }
    }
}

template <int S, int SF, bool RING>
__global__ void __launch_bounds__(256) rroi_fwd_nchw_rows_kernel(const FwdParams p) {
    constexpr int kStageFloats = SF;              // floats per stage
    constexpr int kGranMax = SF / 4;              // granules of one channel
    static_assert(SF % 1024 == 0 && (RING || S == 2), "stage = whole rounds of 256 granules; the double buffer has two stages");
    extern __shared__ __align__(16) unsigned char rows_smem[];
    float (*stage)[kStageFloats] = reinterpret_cast<float (*)[kStageFloats]>(rows_smem);
    int* gsrc = reinterpret_cast<int*>(rows_smem + (size_t)S * SF * 4);   // plane offset (in floats) of every granule of the footprint
    __shared__ int rlo[kRowsMax], rhi[kRowsMax], goff[kRowsMax + 1];
    __shared__ int ylim[2];                       // first / last image row of the footprint
    __shared__ RoiXform sX;
    const int n = blockIdx.x / p.tiles;
    const int tile = blockIdx.x - n * p.tiles;
    const int tiles_w = (p.PW + 31) / 32;
    const int ph0 = (tile / tiles_w) * 8, pw0 = (tile % tiles_w) * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bins = p.PH * p.PW;

    if (threadIdx.x < kRowsMax) { rlo[threadIdx.x] = INT_MAX; rhi[threadIdx.x] = INT_MIN; }
    if (threadIdx.x == 0) { ylim[0] = INT_MAX; ylim[1] = INT_MIN; }
    cta_xform_prologue(p, n, &sX);                // includes the grid-dependency wait and a __syncthreads()
    const RoiXform X = sX;
    const bool batch_ok = (X.batch >= 0) & (X.batch < p.B);

    // ---- 1. geometry: thread = bin (warp = one ph row of the block, lanes = 32 consecutive pw)
    const int ph = ph0 + warp, pw = pw0 + lane;
    const bool live = ph < p.PH && pw < p.PW;
    uint32_t code = 0;
    float wlt = 0.f, wrt = 0.f, wrb = 0.f, wlb = 0.f, ccx = 0.f, ccy = 0.f;
    int gl = 0, gt = 0;
    bool two_c = false, two_r = false;
    if (live) {
        const BinTaps g = bin_taps(X, ph, pw, p.H, p.W, (float)(p.W - 1), (float)(p.H - 1), batch_ok);
        const bool in = g.flags & BIN_IN;
        two_c = g.flags & TWO_COLS; two_r = g.flags & TWO_ROWS;
        code = C_LIVE;
        gl = g.l; gt = g.t;
        if (in) {
            const bool nanw = !(fabsf(g.cx) < INFINITY) || !(fabsf(g.cy) < INFINITY);
            const bool l_lt = g.flags & TAP_LT;
            const bool l_rt = (g.flags & TAP_RT) && two_c;
            const bool l_lb = (g.flags & TAP_LB) && two_r;
            const bool l_rb = (g.flags & TAP_RB) && two_c && two_r;
            code |= (l_lt ? C_LT : 0u) | (l_rt ? C_RT : 0u) | (l_lb ? C_LB : 0u) | (l_rb ? C_RB : 0u);
            wlt = (l_lt || nanw) ? g.wlt : 0.0f;
            wrt = (l_rt || nanw) ? g.wrt : 0.0f;
            wrb = (l_rb || nanw) ? g.wrb : 0.0f;
            wlb = (l_lb || nanw) ? g.wlb : 0.0f;
            ccx = g.cx; ccy = g.cy;
        }
        if (p.idx_mode == IDX_COMPACT) {
            p.idx_x[(size_t)n * bins + ph * p.PW + pw] = ccx;
            p.idx_y[(size_t)n * bins + ph * p.PW + pw] = ccy;
        }
    }
    const bool top_used = code & (C_LT | C_RT), bot_used = code & (C_LB | C_RB);
    // loaded taps lie inside the image (0 < x < W, 0 < y < H), so rows / columns below are valid indices
    {
        int ya = INT_MAX, yb = INT_MIN;
        if (top_used) { ya = gt; yb = gt; }
        if (bot_used) { ya = min(ya, gt + 1); yb = max(yb, gt + 1); }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            ya = min(ya, __shfl_xor_sync(0xffffffffu, ya, m));
            yb = max(yb, __shfl_xor_sync(0xffffffffu, yb, m));
        }
        if (lane == 0 && yb >= ya) { atomicMin(&ylim[0], ya); atomicMax(&ylim[1], yb); }
    }
    __syncthreads();
    const int y0 = ylim[0], nrows = ylim[1] >= ylim[0] ? ylim[1] - ylim[0] + 1 : 0;
    const bool aligned = (p.W % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.feat) & 15) == 0);
    bool staged = aligned && nrows > 0 && nrows <= kRowsMax && (size_t)p.H * p.W <= (1u << 28);   // CTA-uniform (32-bit copy offsets)
    if (staged) {
        // ---- 2. per-row spans of the loaded taps
        if (top_used) {
            const int xa = (code & C_LT) ? gl : gl + 1, xb = (code & C_RT) ? gl + 1 : gl;
            atomicMin(&rlo[gt - y0], xa); atomicMax(&rhi[gt - y0], xb);
        }
        if (bot_used) {
            const int xa = (code & C_LB) ? gl : gl + 1, xb = (code & C_RB) ? gl + 1 : gl;
            atomicMin(&rlo[gt + 1 - y0], xa); atomicMax(&rhi[gt + 1 - y0], xb);
        }
    }
    __syncthreads();
    if (staged && warp == 0) {                    // exclusive scan of the rows' granule counts (<= 96 rows: 3 per lane)
        int len[3], sum = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int r = lane * 3 + k;
            len[k] = (r < nrows && rhi[r] >= rlo[r]) ? ((rhi[r] | 3) - (rlo[r] & ~3) + 1) >> 2 : 0;
            sum += len[k];
        }
        int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        int run = incl - sum;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int r = lane * 3 + k;
            if (r < kRowsMax) goff[r] = run;
            run += len[k];
        }
        if (lane == 31) goff[kRowsMax] = incl;    // total granules of one channel
    }
    __syncthreads();
    const int gtot = staged ? goff[kRowsMax] : 0;
    staged = staged && gtot > 0 && gtot < kGranMax;        // + one zero granule per channel
    if (staged) {
        // table: plane offset of every granule (row r owns granules goff[r] .. goff[r+1])
        for (int r = threadIdx.x; r < nrows; r += 256) {
            if (rhi[r] < rlo[r]) continue;
            const int a = rlo[r] & ~3, cnt = ((rhi[r] | 3) - a + 1) >> 2, base = (y0 + r) * p.W + a;
            for (int k = 0; k < cnt; ++k) gsrc[goff[r] + k] = base + 4 * k;
        }
    }
    __syncthreads();

    const size_t HW = (size_t)p.H * p.W;
    const size_t obin = (size_t)ph * p.PW + pw;
    float* dst = p.out + (size_t)n * p.C * bins + obin;
    const bool full_idx = p.idx_mode == IDX_FULL;
    const float* plane0 = p.feat + (size_t)(batch_ok ? X.batch : 0) * p.C * HW;

    if (staged) {
        // Every staged channel is followed by one granule of zeros; a tap that is not loaded (border test, coinciding tap,
        // bin outside the RoI) reads that zero, so the blend is branch-free: 4 LDS + 4 FFMA + 1 STG per channel.
        const int zero = gtot * 4;
        int o_lt = zero, o_rt = zero, o_lb = zero, o_rb = zero;
        if (top_used) {
            const int o = goff[gt - y0] * 4 + (gl - (rlo[gt - y0] & ~3));
            if (code & C_LT) o_lt = o;
            if (code & C_RT) o_rt = o + 1;
        }
        if (bot_used) {
            const int o = goff[gt + 1 - y0] * 4 + (gl - (rlo[gt + 1 - y0] & ~3));
            if (code & C_LB) o_lb = o;
            if (code & C_RB) o_rb = o + 1;
        }
        RowsStageArgs a;
        a.stage0 = &stage[0][0]; a.gsrc = gsrc; a.plane0 = plane0; a.dst = dst; a.HW = HW; a.bins = (size_t)bins;
        a.C = p.C; a.gtot = gtot; a.live = live;
        a.o_lt = o_lt; a.o_rt = o_rt; a.o_lb = o_lb; a.o_rb = o_rb; a.wlt = wlt; a.wrt = wrt; a.wrb = wrb; a.wlb = wlb;
        // channels per stage: the largest power of two whose slot (SF / cgs floats) holds the footprint + its zero granule
        const int need = gtot * 4 + 4;                                             // <= SF: gtot < kGranMax
        if (need <= SF / 8) rows_stage_loop<S, SF, RING, 8>(a);
        else if (need <= SF / 4) rows_stage_loop<S, SF, RING, 4>(a);
        else if (need <= SF / 2) rows_stage_loop<S, SF, RING, 2>(a);
        else rows_stage_loop<S, SF, RING, 1>(a);
        if (full_idx && live) {                                                    // legacy [N,C,PH,PW] centre tensors
            for (int c = 0; c < p.C; ++c) {
                p.idx_x[((size_t)n * p.C + c) * bins + obin] = ccx;
                p.idx_y[((size_t)n * p.C + c) * bins + obin] = ccy;
            }
        }
        return;
    }

    // ---- fallback: nothing to load, footprint too large, or unaligned rows ----
    if (!live) return;
    if (nrows == 0) {
        // no bin of this block loads anything (the zero tail pw > roi_pooled_width, or a RoI off the image): the reference's
        // sum of four 0-weight products, i.e. weights * 0 -- NaN weights (non-finite centre) still give NaN
        const float v = __fmaf_rn(0.0f, wlb, __fmaf_rn(wrb, 0.0f, __fmaf_rn(0.0f, wrt, __fmaf_rn(0.0f, wlt, 0.0f))));
        float* d = dst;
        for (int c = 0; c < p.C; ++c, d += bins) *d = v;
        if (full_idx)
            for (int c = 0; c < p.C; ++c) {
                p.idx_x[((size_t)n * p.C + c) * bins + obin] = ccx;
                p.idx_y[((size_t)n * p.C + c) * bins + obin] = ccy;
            }
        return;
    }
    const float* top = plane0 + (unsigned)gt * (unsigned)p.W + (unsigned)gl;
    const float* bot = top + p.W;
#pragma unroll 1
    for (int c = 0; c < p.C; ++c) {
        const size_t po = (size_t)c * HW;
        const float lt = ldg_pred_f32<0>(top + po, code & C_LT), rt = ldg_pred_f32<4>(top + po, code & C_RT);
        const float lb = ldg_pred_f32<0>(bot + po, code & C_LB), rb = ldg_pred_f32<4>(bot + po, code & C_RB);
        float v = __fmaf_rn(lt, wlt, 0.0f);
        v = __fmaf_rn(rt, wrt, v);
        v = __fmaf_rn(wrb, rb, v);
        v = __fmaf_rn(lb, wlb, v);
        dst[(size_t)c * bins] = v;
        if (full_idx) {
            p.idx_x[((size_t)n * p.C + c) * bins + obin] = ccx;
            p.idx_y[((size_t)n * p.C + c) * bins + obin] = ccy;
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn tma_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* q = nullptr;
        cudaDriverEntryPointQueryResult res;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &res) == cudaSuccess &&
            res == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(q);
    });
    return fn;
}

// true when the four maps could be built: planes [B*C][H][W] fp32, rows a multiple of 16 bytes, 16-byte aligned base
static bool make_nchw_maps(const FwdParams& p, NchwTmaMaps* out) {
    if (p.B <= 0 || p.B > (1 << 20) || (p.W % 4) != 0 || (reinterpret_cast<uintptr_t>(p.feat) & 15)) return false;
    if (!tma_encode_fn()) return false;
    const cuuint64_t planes = (cuuint64_t)p.B * (cuuint64_t)p.C;
    if (planes > 0xffffffffull) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)p.W, (cuuint64_t)p.H, planes};
    const cuuint64_t strides[2] = {(cuuint64_t)p.W * 4, (cuuint64_t)p.H * p.W * 4};
    const cuuint32_t es[3] = {1, 1, 1};
    const int side[4] = {8, 16, 24, 32}, cg[4] = {32, 16, 8, 4};
    for (int k = 0; k < 4; ++k) {
        const cuuint32_t box[3] = {(cuuint32_t)side[k] + 4, (cuuint32_t)side[k], (cuuint32_t)cg[k]};
        if (tma_encode_fn()(&out->m[k], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(p.feat), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    return true;
}

// cudaFuncSetAttribute is per device: remember per device ordinal which devices have been opted in (the attribute
// call itself is idempotent, so a benign race between threads only repeats it)
static cudaError_t tma_kernel_smem_optin() {
    static bool done[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(rroi_fwd_nchw_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kTmaStageBytes);
    if (e == cudaSuccess && dev >= 0 && dev < 64) done[dev] = true;
    return e;
}

// Row-segment kernel: dynamic shared memory (stages + granule table), opted in per device ordinal like the TMA kernel.
template <int S, int SF, bool RING>
static cudaError_t launch_rows(long long grid, const FwdParams& p, cudaStream_t s, bool pdl) {
    constexpr int kSmem = S * SF * 4 + SF;                                          // stages + int gsrc[SF / 4]
    static bool done[64] = {};
    if (grid <= 0) return cudaSuccess;
    if (grid > 2147483647LL) return cudaErrorInvalidConfiguration;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (!(dev >= 0 && dev < 64 && done[dev])) {
        e = cudaFuncSetAttribute(rroi_fwd_nchw_rows_kernel<S, SF, RING>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) done[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = kSmem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, rroi_fwd_nchw_rows_kernel<S, SF, RING>, p);
}

cudaError_t launch_fwd_nchw(const FwdParams& p0, const Opts& o, cudaStream_t s) {
    FwdParams p = p0;
    p.tiles = ((p.PH + kPatch - 1) / kPatch) * ((p.PW + kPatch - 1) / kPatch);
    p.cgroups = 1;
    const long long grid = (long long)p.N * p.tiles;
    const bool pdl = o.pdl;
    // Measured on B200: the TMA-staged kernel is bit-identical but not faster than the gather kernel
    // (5.46 vs 5.54 us on cfg1, 179.8 vs 183.2 us on cfg4's per-GPU batch) -- the boxes over-fetch 3-4x from L2, which
    // trades the L1-wavefront bound for an L2-bandwidth bound -- and it costs four tensor-map encodes per call on the
    // host, so it is opt-in (opts.nchw_tma >= 1).
    if (o.nchw_tma >= 1 && p.B != 0x7fffffff) {      // the legacy launcher does not know the batch size
        p.cgroups = o.nchw_tma >= 2 ? o.nchw_tma : 1;
        NchwTmaMaps maps;
        if (make_nchw_maps(p, &maps)) {
            if (grid <= 0) return cudaSuccess;
            if (grid > 2147483647LL) return cudaErrorInvalidConfiguration;
            const cudaError_t ea = tma_kernel_smem_optin();
            if (ea != cudaSuccess) return ea;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)grid);
            cfg.blockDim = dim3(kNchwBlock);
            cfg.dynamicSmemBytes = 2 * kTmaStageBytes;
            cfg.stream = s;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at;
            cfg.numAttrs = pdl ? 1 : 0;
            return cudaLaunchKernelEx(&cfg, rroi_fwd_nchw_tma_kernel, p, maps.m[0], maps.m[1], maps.m[2], maps.m[3]);
        }
    }
    // Row-segment staging (rroi_fwd_nchw_rows_kernel) vs the gather kernel, measured on B200 (profiles/r02_sweep_nchw.txt):
    // cfg4's per-GPU batch 130.6 vs 180.5 us, cfg1 on 8 streams 3.91 vs 5.58 us per launch -- but one cfg1 launch ALONE
    // 20.2 vs 9.7 us (128 CTAs, each a chain of 8 dependent stage loads).  So: staged when the launch, together with the
    // launches the caller overlaps with it, fills the machine; gather otherwise.  variant 1 / 2 force gather / staged.
    {
        const long long rtiles = (long long)p.N * ((p.PH + 7) / 8) * ((p.PW + 31) / 32);
        const bool staged = o.variant == 2 || (o.variant == 0 && o.nchw_cg == 0 && p.idx_mode != IDX_FULL &&
                                               rtiles * (o.concurrency > 1 ? o.concurrency : 1) >= 148 * 4);
        if (o.variant == 3 || o.variant == 4) {       // ring forms kept for the sweep (profiles/r02_sweep_nchw3.txt): no gain
            p.tiles = ((p.PH + 7) / 8) * ((p.PW + 31) / 32);
            return o.variant == 3 ? launch_rows<3, 4096, true>(rtiles, p, s, pdl) : launch_rows<4, 3072, true>(rtiles, p, s, pdl);
        }
        if (staged) {
            p.tiles = ((p.PH + 7) / 8) * ((p.PW + 31) / 32);
            return launch_rows<2, 5120, false>(rtiles, p, s, pdl);
        }
    }
    switch (o.nchw_cg) {     // channels in flight per lane
        case 1:  return launch_1d(rroi_fwd_nchw_kernel<1>, grid, kNchwBlock, p, s, pdl);
        case 2:  return launch_1d(rroi_fwd_nchw_kernel<2>, grid, kNchwBlock, p, s, pdl);
        case 8:  return launch_1d(rroi_fwd_nchw_kernel<8>, grid, kNchwBlock, p, s, pdl);
        case 16: return launch_1d(rroi_fwd_nchw_kernel<16>, grid, kNchwBlock, p, s, pdl);
        default: return launch_1d(rroi_fwd_nchw_kernel<4>, grid, kNchwBlock, p, s, pdl);
    }
}

// ------------------------------------------------------------------------------------------ NHWC
// Channels-last: feat [B,H,W,C], out [N,PH,PW,C].  One tap of one bin is a contiguous C*4-byte vector.
//
// grid = N * tiles; CTA = 8 warps = one RoI x (8*PPW) consecutive bins.  Warp 0 computes the RoI
// transform once (shared memory broadcast).  Each warp then owns PPW bins:
//   geometry -- lane j computes bin j (lane-parallel, no redundancy), keeping only a float offset of
//               the top-left tap and a small code word;
//   gather   -- the WHOLE warp walks its bins one at a time, lane = channel vector, so every tap is
//               one fully coalesced 32*VEC*4-byte request, every branch is warp-uniform, and there
//               is no per-lane select/predicate arithmetic.  UN bins are in flight at once
//               (loads of UN bins are issued before the first blend).
// Per bin this is ~25 warp instructions for C=64 (3 SHFL, <=4 LDG, 8 FFMA, 1 STG, addressing), against
// ~40 thread-instructions per output float in a one-thread-per-vector formulation.
constexpr int kNhwcWarps = 8;


template <int VEC> struct VecT;
template <> struct VecT<1> { using T = float; };
template <> struct VecT<2> { using T = float2; };
template <> struct VecT<4> { using T = float4; };

template <int VEC> __device__ __forceinline__ void vzero(float (&v)[VEC]) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) v[i] = 0.0f;
}
template <int VEC> __device__ __forceinline__ void vload(float (&v)[VEC], const float* p) {
    using T = typename VecT<VEC>::T;
    const T t = __ldg(reinterpret_cast<const T*>(p));
    const float* f = reinterpret_cast<const float*>(&t);
#pragma unroll
    for (int i = 0; i < VEC; ++i) v[i] = f[i];
}
template <int VEC> __device__ __forceinline__ void vstore(float* p, const float (&v)[VEC]) {
    using T = typename VecT<VEC>::T;
    T t;
    float* f = reinterpret_cast<float*>(&t);
#pragma unroll
    for (int i = 0; i < VEC; ++i) f[i] = v[i];
    *reinterpret_cast<T*>(p) = t;
}

// kernel.cu:136-141 with the weights of the four half-grid cases as literals.  Taps that coincide
// (r == l and/or b == t) reuse the loaded register, exactly as the reference re-reads the same pixel.
template <int VEC>
__device__ __forceinline__ void blend_case(float (&o)[VEC], const float (&lt)[VEC], const float (&rt)[VEC],
                                           const float (&lb)[VEC], const float (&rb)[VEC], uint32_t code) {
    const bool hx = code & C_HX, hy = code & C_HY;
    if (!hx && !hy) {            // weights 1,0,0,0 ; rt = rb = lb = lt
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            float v = __fmaf_rn(lt[i], 1.0f, 0.0f);
            v = __fmaf_rn(lt[i], 0.0f, v); v = __fmaf_rn(0.0f, lt[i], v); v = __fmaf_rn(lt[i], 0.0f, v);
            o[i] = v;
        }
    } else if (hx && !hy) {      // weights .5,.5,0,0 ; lb = lt, rb = rt
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            float v = __fmaf_rn(lt[i], 0.5f, 0.0f);
            v = __fmaf_rn(rt[i], 0.5f, v); v = __fmaf_rn(0.0f, rt[i], v); v = __fmaf_rn(lt[i], 0.0f, v);
            o[i] = v;
        }
    } else if (!hx && hy) {      // weights .5,0,0,.5 ; rt = lt, rb = lb
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            float v = __fmaf_rn(lt[i], 0.5f, 0.0f);
            v = __fmaf_rn(lt[i], 0.0f, v); v = __fmaf_rn(0.0f, lb[i], v); v = __fmaf_rn(lb[i], 0.5f, v);
            o[i] = v;
        }
    } else {                     // weights .25 x4
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            float v = __fmaf_rn(lt[i], 0.25f, 0.0f);
            v = __fmaf_rn(rt[i], 0.25f, v); v = __fmaf_rn(0.25f, rb[i], v); v = __fmaf_rn(lb[i], 0.25f, v);
            o[i] = v;
        }
    }
}

// CT = compile-time channel count (64/128/256: tap strides become LDG immediates and a warp spans the
// channels exactly) or 0 = run-time p.C with vector width VEC and a ragged last channel chunk.
template <int CT, int VEC, int PPW, int UN>
__global__ void __launch_bounds__(kNhwcWarps * 32) rroi_fwd_nhwc_kernel(const FwdParams p) {
    __shared__ RoiXform sX;
    const int C = CT ? CT : p.C;
    const int n = blockIdx.x / p.tiles;
    const int tile = blockIdx.x - n * p.tiles;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bins = p.PH * p.PW;

    cta_xform_prologue(p, n, &sX);

    // ---- geometry: lane j <-> bin bin0 + j
    const int bin0 = (tile * kNhwcWarps + warp) * PPW;
    if (bin0 >= bins) return;                         // warp-uniform
    int pix = 0;                                      // pixel index (batch*H + t)*W + l of the top-left tap
    uint32_t code = 0;
    if (lane < PPW && bin0 + lane < bins) {
        const RoiXform X = sX;
        const int bin = bin0 + lane;
        const int ph = bin / p.PW, pw = bin - ph * p.PW;
        const bool batch_ok = (X.batch >= 0) & (X.batch < p.B);
        const BinTaps g = bin_taps(X, ph, pw, p.H, p.W, (float)(p.W - 1), (float)(p.H - 1), batch_ok);
        const bool in = g.flags & BIN_IN, two_c = g.flags & TWO_COLS, two_r = g.flags & TWO_ROWS;
        // only dereferenced under a tap bit, and then 0 <= t < H, 0 <= l < W (wrap-around is harmless otherwise)
        pix = (int)(((unsigned)(batch_ok ? X.batch : 0) * (unsigned)p.H + (unsigned)g.t) * (unsigned)p.W + (unsigned)g.l);
        code = C_LIVE;
        if (in) {
            code |= C_IN;
            if (g.flags & TAP_LT) code |= C_LT;
            if ((g.flags & TAP_RT) && two_c) code |= C_RT;
            if ((g.flags & TAP_LB) && two_r) code |= C_LB;
            if ((g.flags & TAP_RB) && two_c && two_r) code |= C_RB;
            if (two_c) code |= C_HX;                  // r != l  <=>  rx == 0.5 for a finite centre
            if (two_r) code |= C_HY;
            if (!(fabsf(g.cx) < INFINITY) || !(fabsf(g.cy) < INFINITY)) code |= C_NAN;
        }
        if (p.idx_mode == IDX_COMPACT) {
            p.idx_x[(size_t)n * bins + bin] = in ? g.cx : 0.0f;
            p.idx_y[(size_t)n * bins + bin] = in ? g.cy : 0.0f;
        }
    }

    // ---- gather: whole warp per bin, lane = channel vector
    const int CV = C / VEC;
    const long long rowC = (long long)p.W * C;
    constexpr bool kExact = CT != 0;                 // CT % (32*VEC) == 0: no ragged chunk
    for (int cv = lane; cv < (kExact ? CV : ((CV + 31) & ~31)); cv += 32) {
        const bool act = kExact || cv < CV;
        const float* fbase = p.feat + (size_t)cv * VEC;
        float* outp = p.out + ((size_t)n * bins + bin0) * C + (size_t)cv * VEC;
#pragma unroll 1
        for (int j0 = 0; j0 < PPW; j0 += UN) {
            float lt[UN][VEC], rt[UN][VEC], lb[UN][VEC], rb[UN][VEC];
            uint32_t cd[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int px = __shfl_sync(0xffffffffu, pix, j0 + u);
                cd[u] = __shfl_sync(0xffffffffu, code, j0 + u);
                vzero<VEC>(lt[u]); vzero<VEC>(rt[u]); vzero<VEC>(lb[u]); vzero<VEC>(rb[u]);
                if (act) {
                    const float* s = fbase + (long long)px * C;
                    const float* s2 = s + rowC;
                    if (cd[u] & C_LT) vload<VEC>(lt[u], s);
                    if (cd[u] & C_RT) vload<VEC>(rt[u], s + C);
                    if (cd[u] & C_LB) vload<VEC>(lb[u], s2);
                    if (cd[u] & C_RB) vload<VEC>(rb[u], s2 + C);
                }
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                if (!(cd[u] & C_LIVE) || !act) continue;
                float o[VEC];
                if (!(cd[u] & C_IN)) {
                    vzero<VEC>(o);
                } else if (cd[u] & C_NAN) {          // reference: every tap invalid, weights NaN -> 0*NaN
#pragma unroll
                    for (int i = 0; i < VEC; ++i) o[i] = __int_as_float(0x7fffffff);
                } else {
                    blend_case<VEC>(o, lt[u], rt[u], lb[u], rb[u], cd[u]);
                }
                vstore<VEC>(outp + (j0 + u) * C, o);
            }
        }
    }
}

// ------------------------------------------------------------------------------------ NHWC, packed
// Specialisation for C in {32, 64, 128, 256} (C*4 bytes per tap = 128 B .. 1 KB): every lane moves one
// float4, LPP = C/4 lanes span a pixel, so a warp iteration covers 32/LPP bins (C <= 128) or one bin's
// 32-lane channel chunk (C = 256).  The record is per lane, so the body is branch-free: predicated
// 128-bit loads and the reference's 4-FFMA chain with per-lane weights.
// Per 512 output bytes: 2 LDS.128 + <=4 LDG.128 + 16 FFMA + 1 STG.128 + ~15 integer/zeroing.
// Geometry uses one THREAD per bin of the CTA tile (shared-memory records), so it costs the same
// ~6 warp-instructions per bin whatever the tile shape.
// 128-bit load under a predicate; the destination is zero when the predicate is off.  Deliberately not
// ld.global.nc: ptxas moves .nc loads across barriers, which defeats the load batching below.
__device__ __forceinline__ float4 ldg_pred_v4(const float* ptr, uint32_t pred) {
    float4 r;
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t"
        "mov.b32 %0, 0;\n\tmov.b32 %1, 0;\n\tmov.b32 %2, 0;\n\tmov.b32 %3, 0;\n\t"
        "@q ld.global.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
        : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
        : "l"(ptr), "r"(pred));
    return r;
}

// Shared-memory record of one bin.  The weights already fold in everything the blend needs to know:
// a tap that is not loaded (outside the border test, coinciding with another tap, or bin outside the
// RoI) has weight 0 and its register holds 0, so the reference's four-product sum comes out of the same
// four FFMAs in the same order without any per-lane select.  (For finite features this is bit-identical
// to the reference; a zero-weight tap whose value is inf/NaN makes the reference return NaN where this
// returns the finite/inf bilinear value -- documented in DESIGN.md.)
struct __align__(16) BinRec {
    int pix;                       // (batch*H + t)*W + l of the top-left tap
    uint32_t code;                 // C_LT|C_RT|C_LB|C_RB load bits, C_LIVE
    float wlt, wrt, wrb, wlb;      // NaN when the centre is not finite (reference: 0 * NaN)
    int pad[2];
};

template <int CT, int TILE, int UN>
__global__ void __launch_bounds__(kNhwcWarps * 32) rroi_fwd_nhwc_packed_kernel(const FwdParams p) {
    constexpr int LPP = CT / 4;                          // lanes per pixel
    constexpr int PPI = LPP >= 32 ? 1 : 32 / LPP;        // pixels per warp iteration
    constexpr int NCH = LPP > 32 ? LPP / 32 : 1;         // 32-lane channel chunks per pixel
    constexpr int PIXW = TILE / kNhwcWarps;              // pixels per warp
    constexpr int ITERS = PIXW * NCH / PPI;              // warp iterations
    static_assert(PIXW * kNhwcWarps == TILE && ITERS % UN == 0 && ITERS > 0, "tile shape");
    __shared__ RoiXform sX;
    __shared__ BinRec rec[TILE];

    const int n = blockIdx.x / p.tiles;
    const int tile = blockIdx.x - n * p.tiles;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bins = p.PH * p.PW;
    const int bin0 = tile * TILE;

    cta_xform_prologue(p, n, &sX);
    for (int t = threadIdx.x; t < TILE; t += kNhwcWarps * 32) {
        BinRec r;
        r.pix = 0; r.code = 0; r.wlt = r.wrt = r.wrb = r.wlb = 0.0f; r.pad[0] = r.pad[1] = 0;
        const int bin = bin0 + t;
        if (bin < bins) {
            const RoiXform X = sX;
            const int ph = bin / p.PW, pw = bin - ph * p.PW;
            const bool batch_ok = (X.batch >= 0) & (X.batch < p.B);
            const BinTaps g = bin_taps(X, ph, pw, p.H, p.W, (float)(p.W - 1), (float)(p.H - 1), batch_ok);
            const bool in = g.flags & BIN_IN, two_c = g.flags & TWO_COLS, two_r = g.flags & TWO_ROWS;
            r.pix = (int)(((unsigned)(batch_ok ? X.batch : 0) * (unsigned)p.H + (unsigned)g.t) * (unsigned)p.W + (unsigned)g.l);
            r.code = C_LIVE;
            if (in) {
                const bool nanw = !(fabsf(g.cx) < INFINITY) || !(fabsf(g.cy) < INFINITY);
                const bool l_lt = g.flags & TAP_LT;
                const bool l_rt = (g.flags & TAP_RT) && two_c;
                const bool l_lb = (g.flags & TAP_LB) && two_r;
                const bool l_rb = (g.flags & TAP_RB) && two_c && two_r;
                r.code |= (l_lt ? C_LT : 0u) | (l_rt ? C_RT : 0u) | (l_lb ? C_LB : 0u) | (l_rb ? C_RB : 0u);
                // Coinciding taps: the reference adds w*x once per tap NAME; here the duplicate names'
                // weights are zero by construction (rx == 0 -> wrt = wrb = 0, ry == 0 -> wrb = wlb = 0),
                // so only border-invalid taps need their weight cleared.
                r.wlt = (l_lt || nanw) ? g.wlt : 0.0f;
                r.wrt = (l_rt || nanw) ? g.wrt : 0.0f;
                r.wrb = (l_rb || nanw) ? g.wrb : 0.0f;
                r.wlb = (l_lb || nanw) ? g.wlb : 0.0f;
            }
            if (p.idx_mode == IDX_COMPACT) {
                p.idx_x[(size_t)n * bins + bin] = in ? g.cx : 0.0f;
                p.idx_y[(size_t)n * bins + bin] = in ? g.cy : 0.0f;
            }
        }
        rec[t] = r;
    }
    __syncthreads();

    const int sub = LPP >= 32 ? 0 : lane / LPP;          // which of the iteration's pixels this lane serves
    const int cvl = LPP >= 32 ? lane : lane % LPP;       // float4 index inside the 32-lane chunk
    const long long rowC = (long long)p.W * CT;
    const float* fbase = p.feat + cvl * 4;
    const int pw0 = warp * PIXW;
    // per-lane output pointer of iteration 0; later iterations are compile-time offsets from it
    float* obase = p.out + ((size_t)n * bins + bin0 + pw0 + (NCH > 1 ? 0 : sub)) * CT + cvl * 4;
    const BinRec* rbase = rec + pw0 + (NCH > 1 ? 0 : sub);

#pragma unroll 1
    for (int it0 = 0; it0 < ITERS; it0 += UN) {
        float4 lt[UN], rt[UN], lb[UN], rb[UN];
        BinRec r[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int it = it0 + u;
            r[u] = rbase[NCH > 1 ? it / NCH : it * PPI];
        }
        // all loads of the UN iterations are issued before the first blend (volatile asm keeps them
        // together and in order: ptxas otherwise serialises the iterations to save registers)
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int it = it0 + u;
            const int ch = NCH > 1 ? (it % NCH) * 128 : 0;            // float offset of the channel chunk
            const float* s = fbase + (long long)r[u].pix * CT + ch;
            const float* s2 = s + rowC;
            lt[u] = ldg_pred_v4(s, r[u].code & C_LT);
            rt[u] = ldg_pred_v4(s + CT, r[u].code & C_RT);
            lb[u] = ldg_pred_v4(s2, r[u].code & C_LB);
            rb[u] = ldg_pred_v4(s2 + CT, r[u].code & C_RB);
        }
        // The blend lives in its own basic block (a branch the optimiser cannot fold): ptxas otherwise
        // interleaves iteration u's FFMAs between the loads of iterations u and u+1 to save registers,
        // which serialises the DRAM latencies of the UN iterations.
        uint32_t any_live = 0;
#pragma unroll
        for (int u = 0; u < UN; ++u) any_live |= r[u].code;
        if (any_live & C_LIVE) {
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int it = it0 + u;
                const int dpx = NCH > 1 ? it / NCH : it * PPI;        // pixel offset from the lane's base
                const int ch = NCH > 1 ? (it % NCH) * 128 : 0;
                const float wlt = r[u].wlt, wrt = r[u].wrt, wrb = r[u].wrb, wlb = r[u].wlb;
                float4 o;
                float v;
                v = __fmaf_rn(lt[u].x, wlt, 0.0f); v = __fmaf_rn(rt[u].x, wrt, v); v = __fmaf_rn(wrb, rb[u].x, v); o.x = __fmaf_rn(lb[u].x, wlb, v);
                v = __fmaf_rn(lt[u].y, wlt, 0.0f); v = __fmaf_rn(rt[u].y, wrt, v); v = __fmaf_rn(wrb, rb[u].y, v); o.y = __fmaf_rn(lb[u].y, wlb, v);
                v = __fmaf_rn(lt[u].z, wlt, 0.0f); v = __fmaf_rn(rt[u].z, wrt, v); v = __fmaf_rn(wrb, rb[u].z, v); o.z = __fmaf_rn(lb[u].z, wlb, v);
                v = __fmaf_rn(lt[u].w, wlt, 0.0f); v = __fmaf_rn(rt[u].w, wrt, v); v = __fmaf_rn(wrb, rb[u].w, v); o.w = __fmaf_rn(lb[u].w, wlb, v);
                if (r[u].code & C_LIVE) *reinterpret_cast<float4*>(obase + dpx * CT + ch) = o;
            }
        }
    }
}

// ------------------------------------------------------------------------- NHWC, packed, warp-autonomous
// Same lanes, records and blend as the block-level kernel above, but every WARP is its own unit of work: it owns
// BPW consecutive bins of one RoI, evaluates the RoI transform itself (all lanes redundantly -- no shared-memory
// broadcast), computes its bins' geometry lane-parallel into its private slice of shared memory and starts loading
// after a __syncwarp().  No __syncthreads() anywhere: a single small launch (cfg1: 64 RoIs) is a pure latency chain
// RoI row -> transform -> geometry -> taps -> store, and the two block barriers of the block-level kernel made
// every warp wait for the slowest one twice (ncu: barrier = 5.1 of 17 stall cycles per issued instruction).
// With p.early (RROI_B200_FLAG_ROIS_READY) the whole prologue runs before griddepcontrol.wait, i.e. while the
// previous kernel in the stream drains; only the feature reads and the stores wait for the grid dependency.
template <int CT, int BPW, int UN>
__global__ void __launch_bounds__(kNhwcWarps * 32) rroi_fwd_nhwc_warp_kernel(const FwdParams p) {
    constexpr int LPP = CT / 4;
    constexpr int PPI = LPP >= 32 ? 1 : 32 / LPP;
    constexpr int NCH = LPP > 32 ? LPP / 32 : 1;
    constexpr int ITERS = BPW * NCH / PPI;
    static_assert(ITERS > 0 && ITERS % UN == 0 && BPW % PPI == 0, "segment shape");
    __shared__ BinRec rec[kNhwcWarps][BPW];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * kNhwcWarps + warp;       // global warp = (RoI, segment)
    const int n = (int)(gw / p.tiles);
    const int bins = p.PH * p.PW;
    const int bin0 = (int)(gw - (long long)n * p.tiles) * BPW;
    if (!p.early) pdl_wait();
    pdl_launch_dependents();
    if (n >= p.N) return;                                                  // warp-uniform (ragged last CTA)

    const RoiXform X = get_xform(p, n);                                    // every lane: same addresses, same result
    const bool batch_ok = (X.batch >= 0) & (X.batch < p.B);
    float ccx[(BPW + 31) / 32], ccy[(BPW + 31) / 32];
#pragma unroll
    for (int k = 0; k < (BPW + 31) / 32; ++k) {
        const int t = k * 32 + lane;
        ccx[k] = 0.0f; ccy[k] = 0.0f;
        if (t < BPW) {
            BinRec r;
            r.pix = 0; r.code = 0; r.wlt = r.wrt = r.wrb = r.wlb = 0.0f; r.pad[0] = r.pad[1] = 0;
            const int bin = bin0 + t;
            if (bin < bins) {
                const int ph = bin / p.PW, pw = bin - ph * p.PW;
                const BinTaps g = bin_taps(X, ph, pw, p.H, p.W, (float)(p.W - 1), (float)(p.H - 1), batch_ok);
                const bool in = g.flags & BIN_IN, two_c = g.flags & TWO_COLS, two_r = g.flags & TWO_ROWS;
                r.pix = (int)(((unsigned)(batch_ok ? X.batch : 0) * (unsigned)p.H + (unsigned)g.t) * (unsigned)p.W + (unsigned)g.l);
                r.code = C_LIVE;
                if (in) {
                    const bool nanw = !(fabsf(g.cx) < INFINITY) || !(fabsf(g.cy) < INFINITY);
                    const bool l_lt = g.flags & TAP_LT;
                    const bool l_rt = (g.flags & TAP_RT) && two_c;
                    const bool l_lb = (g.flags & TAP_LB) && two_r;
                    const bool l_rb = (g.flags & TAP_RB) && two_c && two_r;
                    r.code |= (l_lt ? C_LT : 0u) | (l_rt ? C_RT : 0u) | (l_lb ? C_LB : 0u) | (l_rb ? C_RB : 0u);
                    r.wlt = (l_lt || nanw) ? g.wlt : 0.0f;
                    r.wrt = (l_rt || nanw) ? g.wrt : 0.0f;
                    r.wrb = (l_rb || nanw) ? g.wrb : 0.0f;
                    r.wlb = (l_lb || nanw) ? g.wlb : 0.0f;
                    ccx[k] = g.cx; ccy[k] = g.cy;
                }
            }
            rec[warp][t] = r;
        }
    }
    __syncwarp();
    if (p.early) pdl_wait();                                               // nothing above touched the caller's outputs
    if (p.idx_mode == IDX_COMPACT) {
#pragma unroll
        for (int k = 0; k < (BPW + 31) / 32; ++k) {
            const int t = k * 32 + lane;
            if (t < BPW && bin0 + t < bins) {
                p.idx_x[(size_t)n * bins + bin0 + t] = ccx[k];
                p.idx_y[(size_t)n * bins + bin0 + t] = ccy[k];
            }
        }
    }

    const int sub = LPP >= 32 ? 0 : lane / LPP;
    const int cvl = LPP >= 32 ? lane : lane % LPP;
    const long long rowC = (long long)p.W * CT;
    const float* fbase = p.feat + cvl * 4;
    float* obase = p.out + ((size_t)n * bins + bin0 + (NCH > 1 ? 0 : sub)) * CT + cvl * 4;
    const BinRec* rbase = rec[warp] + (NCH > 1 ? 0 : sub);

#pragma unroll 1
    for (int it0 = 0; it0 < ITERS; it0 += UN) {
        float4 lt[UN], rt[UN], lb[UN], rb[UN];
        BinRec r[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int it = it0 + u;
            r[u] = rbase[NCH > 1 ? it / NCH : it * PPI];
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int it = it0 + u;
            const int ch = NCH > 1 ? (it % NCH) * 128 : 0;
            const float* s = fbase + (long long)r[u].pix * CT + ch;
            const float* s2 = s + rowC;
            lt[u] = ldg_pred_v4(s, r[u].code & C_LT);
            rt[u] = ldg_pred_v4(s + CT, r[u].code & C_RT);
            lb[u] = ldg_pred_v4(s2, r[u].code & C_LB);
            rb[u] = ldg_pred_v4(s2 + CT, r[u].code & C_RB);
        }
        uint32_t any_live = 0;
#pragma unroll
        for (int u = 0; u < UN; ++u) any_live |= r[u].code;
        if (any_live & C_LIVE) {                         // own basic block: keeps the UN iterations' loads batched
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int it = it0 + u;
                const int dpx = NCH > 1 ? it / NCH : it * PPI;
                const int ch = NCH > 1 ? (it % NCH) * 128 : 0;
                const float wlt = r[u].wlt, wrt = r[u].wrt, wrb = r[u].wrb, wlb = r[u].wlb;
                float4 o;
                float v;
                v = __fmaf_rn(lt[u].x, wlt, 0.0f); v = __fmaf_rn(rt[u].x, wrt, v); v = __fmaf_rn(wrb, rb[u].x, v); o.x = __fmaf_rn(lb[u].x, wlb, v);
                v = __fmaf_rn(lt[u].y, wlt, 0.0f); v = __fmaf_rn(rt[u].y, wrt, v); v = __fmaf_rn(wrb, rb[u].y, v); o.y = __fmaf_rn(lb[u].y, wlb, v);
                v = __fmaf_rn(lt[u].z, wlt, 0.0f); v = __fmaf_rn(rt[u].z, wrt, v); v = __fmaf_rn(wrb, rb[u].z, v); o.z = __fmaf_rn(lb[u].z, wlb, v);
                v = __fmaf_rn(lt[u].w, wlt, 0.0f); v = __fmaf_rn(rt[u].w, wrt, v); v = __fmaf_rn(wrb, rb[u].w, v); o.w = __fmaf_rn(lb[u].w, wlb, v);
                if (r[u].code & C_LIVE) *reinterpret_cast<float4*>(obase + dpx * CT + ch) = o;
            }
        }
    }
}

// EVALUATED AND REMOVED (round 2, profiles/r02_sweep_fwd2.txt): staging every tap of a warp's bins in shared memory with
// cp.async (LDGSTS, zero-fill for taps that are not loaded) so that a warp needs ONE DRAM round trip and no registers
// for loads in flight.  Bit-identical, but slower everywhere: cfg1 alone 5.2-7.2 us (register-staged: 4.2-5.2), on 8
// streams 3.6-4.3 us (2.2), cfg4's per-GPU batch 118-137 us (76).  LDGSTS costs ~8 issue cycles per warp-level copy plus
// the shared-memory round trip, and 8 KB of staging per warp caps occupancy at 26 warps per SM.

template <int CT>
static cudaError_t launch_fwd_nhwc_packed(FwdParams& p, cudaStream_t s, const Opts& o) {
    const int bins = p.PH * p.PW;
    const bool pdl = o.pdl;
    auto go = [&](auto kernel, int tile) {               // block-level kernels: one CTA per (RoI, tile)
        p.tiles = (bins + tile - 1) / tile;
        return launch_1d(kernel, (long long)p.N * p.tiles, kNhwcWarps * 32, p, s, pdl);
    };
    auto gow = [&](auto kernel, int bpw) {               // warp-autonomous kernels: p.tiles = segments (warps) per RoI
        p.tiles = (bins + bpw - 1) / bpw;
        const long long warps = (long long)p.N * p.tiles;
        return launch_1d(kernel, (warps + kNhwcWarps - 1) / kNhwcWarps, kNhwcWarps * 32, p, s, pdl);
    };
    constexpr int PPI = CT >= 128 ? 1 : 128 / CT;   // pixels per warp iteration
    constexpr int NCH = CT > 128 ? CT / 128 : 1;
    // UN chosen so that a 64-bin tile's per-warp iterations (8*NCH/PPI) are a multiple of it
    constexpr int I64 = 8 * NCH / PPI;
    constexpr int I8 = 8 * NCH / PPI, I16 = 16 * NCH / PPI;          // iterations of an 8- / 16-bin warp segment
    int variant = o.variant;
    if (variant == 0) {
        // auto, measured on B200 (profiles/r02_sweep_fwd.txt, cfg1 = 128 CTAs of 256 bins; us per launch):
        //   in flight        1 (alone)   1 + RoIs ready   2      4      8
        //   block,  64 bins    5.14          4.40        3.06   --     2.95        (variant 1)
        //   warp,    8 bins    5.62          4.17        3.17   --     3.12        (variant 11)
        //   warp,   16 bins    6.42          5.23        3.31   2.48   2.48        (variant 12)
        //   block, 256 bins    8.66          8.66        4.71   3.01   2.22        (variant 5)
        // A launch that has the GPU to itself is a latency chain and wants many small units of work (and, when the RoI
        // rows are declared ready, no block barrier between the prologue and the grid dependency); large launches and
        // launches the caller overlaps with others amortise the per-CTA prologue over 256 bins.
        const long long ctas256 = (long long)p.N * ((bins + 255) / 256);
        const long long load = ctas256 * (o.concurrency > 1 ? o.concurrency : 1);
        if (load >= 148 * 4) variant = 5;
        else if (o.concurrency >= 3) variant = 12;
        else variant = (p.early && o.concurrency <= 1) ? 11 : 1;
    }
    switch (variant) {
        case 1:  return go(rroi_fwd_nhwc_packed_kernel<CT, 64, (I64 >= 2 ? 2 : 1)>, 64);
        case 2:  return go(rroi_fwd_nhwc_packed_kernel<CT, 128, 4>, 128);
        case 3:  return go(rroi_fwd_nhwc_packed_kernel<CT, 128, 2>, 128);
        case 4:  return go(rroi_fwd_nhwc_packed_kernel<CT, 256, 4>, 256);
        case 5:  return go(rroi_fwd_nhwc_packed_kernel<CT, 256, 2>, 256);
        case 6:  return go(rroi_fwd_nhwc_packed_kernel<CT, 64, (I64 >= 4 ? 4 : I64)>, 64);
        case 11: return gow(rroi_fwd_nhwc_warp_kernel<CT, 8, (I8 >= 2 ? 2 : 1)>, 8);
        case 12: return gow(rroi_fwd_nhwc_warp_kernel<CT, 16, (I16 >= 2 ? 2 : 1)>, 16);
        case 13: return gow(rroi_fwd_nhwc_warp_kernel<CT, 16, (I16 >= 4 ? 4 : I16)>, 16);
        case 14: return gow(rroi_fwd_nhwc_warp_kernel<CT, 32, 4>, 32);
        case 15: return gow(rroi_fwd_nhwc_warp_kernel<CT, 64, 4>, 64);
        case 16: return gow(rroi_fwd_nhwc_warp_kernel<CT, 32, 2>, 32);
        case 17: return gow(rroi_fwd_nhwc_warp_kernel<CT, 8, (I8 >= 4 ? 4 : I8)>, 8);
        default: return cudaErrorInvalidValue;
    }
}

template <int CT, int VEC>
static cudaError_t launch_fwd_nhwc_vec(FwdParams& p, cudaStream_t s, const Opts& o) {
    const int bins = p.PH * p.PW;
    auto go = [&](auto kernel, int ppw) {
        p.tiles = (bins + kNhwcWarps * ppw - 1) / (kNhwcWarps * ppw);
        return launch_1d(kernel, (long long)p.N * p.tiles, kNhwcWarps * 32, p, s, o.pdl);
    };
    switch (o.variant) {
        case 1:  return go(rroi_fwd_nhwc_kernel<CT, VEC, 8, 4>, 8);
        case 3:  return go(rroi_fwd_nhwc_kernel<CT, VEC, 32, 4>, 32);
        default: return go(rroi_fwd_nhwc_kernel<CT, VEC, 16, 4>, 16);
    }
}

cudaError_t launch_fwd_nhwc(const FwdParams& p0, const Opts& o, cudaStream_t s) {
    FwdParams p = p0;
    p.cgroups = 1;
    const bool al16 = (reinterpret_cast<uintptr_t>(p.feat) | reinterpret_cast<uintptr_t>(p.out)) % 16 == 0;
    if (al16 && p.C == 32)  return launch_fwd_nhwc_packed<32>(p, s, o);
    if (al16 && p.C == 64)  return launch_fwd_nhwc_packed<64>(p, s, o);
    if (al16 && p.C == 128) return launch_fwd_nhwc_packed<128>(p, s, o);
    if (al16 && p.C == 256) return launch_fwd_nhwc_packed<256>(p, s, o);
    // any other C: warp-per-bin kernel with run-time C, widest vector that divides it
    if (al16 && p.C % 4 == 0 && p.C >= 128) return launch_fwd_nhwc_vec<0, 4>(p, s, o);
    if (al16 && p.C % 2 == 0 && p.C >= 64)  return launch_fwd_nhwc_vec<0, 2>(p, s, o);
    return launch_fwd_nhwc_vec<0, 1>(p, s, o);
}

// ------------------------------------------------------------------------------ NHWC, bf16 -> bf16
// Inference variant for the end-to-end path: the feeder hands over bf16 channels-last
// features and the recogniser's first convolution wants bf16 channels-last input, so sampling bf16 -> bf16
// removes the fp32 copy of the map, the fp32 pooled tensor and the cast back to bf16 (half the bytes of the
// fp32 kernel on both sides).  Arithmetic is unchanged: the four taps are widened to fp32 (exact), blended by
// the reference's 4-FFMA chain (kernel.cu:136-141; weights in {0, 1/4, 1/2, 1}) and the fp32 result is rounded
// to bf16 once (round-to-nearest-even), i.e. out = bf16(reference(float(features))).
// Same shape as the packed fp32 kernel: every lane moves 16 bytes = 8 channels, LPP = C/8 lanes span a pixel.
__device__ __forceinline__ uint4 ldg_pred_u4(const uint4* ptr, uint32_t pred) {
    uint4 r;
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t"
        "mov.b32 %0, 0;\n\tmov.b32 %1, 0;\n\tmov.b32 %2, 0;\n\tmov.b32 %3, 0;\n\t"
        "@q ld.global.v4.b32 {%0, %1, %2, %3}, [%4];\n\t}"
        : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
        : "l"(ptr), "r"(pred));
    return r;
}

// Two fp32 lanes per instruction (sm_100 FFMA2, PTX fma.rn.f32x2): each half is an ordinary IEEE fp32 FMA, so the
// result is bit-identical to the scalar chain; it halves the FFMA issue slots of this (issue-limited) kernel.
__device__ __forceinline__ uint64_t pair_f32(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ uint64_t widen_bf16x2(uint32_t w) {          // {bf16 lo, bf16 hi} -> {f32, f32}, exact
    return pair_f32(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// wXX2 = the tap weight duplicated into both halves
__device__ __forceinline__ uint32_t blend_bf16x2(uint32_t lt, uint32_t rt, uint32_t rb, uint32_t lb, uint64_t wlt2,
                                                 uint64_t wrt2, uint64_t wrb2, uint64_t wlb2) {
    uint64_t v = fma2(widen_bf16x2(lt), wlt2, 0ull);                    // {+0.0f, +0.0f}: the reference's sum starts at 0
    v = fma2(widen_bf16x2(rt), wrt2, v);
    v = fma2(wrb2, widen_bf16x2(rb), v);
    v = fma2(widen_bf16x2(lb), wlb2, v);
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
    uint32_t out;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(out) : "f"(hi), "f"(lo));
    return out;
}

template <int CT, int TILE, int UN, int MINB>
__global__ void __launch_bounds__(kNhwcWarps * 32, MINB) rroi_fwd_nhwc_bf16_kernel(const FwdParams p) {
    constexpr int LPP = CT / 8;                          // lanes per pixel (16 bytes = 8 bf16 per lane)
    constexpr int PPI = 32 / LPP;                        // pixels per warp iteration
    constexpr int PIXW = TILE / kNhwcWarps;              // pixels per warp
    constexpr int ITERS = PIXW / PPI;
    static_assert(LPP >= 1 && LPP <= 32 && PIXW * kNhwcWarps == TILE && ITERS > 0 && ITERS % UN == 0, "tile shape");
    __shared__ RoiXform sX;
    __shared__ BinRec rec[TILE];

    const int n = blockIdx.x / p.tiles;
    const int tile = blockIdx.x - n * p.tiles;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bins = p.PH * p.PW;
    const int bin0 = tile * TILE;

    cta_xform_prologue(p, n, &sX);
    for (int t = threadIdx.x; t < TILE; t += kNhwcWarps * 32) {
        BinRec r;
        r.pix = 0; r.code = 0; r.wlt = r.wrt = r.wrb = r.wlb = 0.0f; r.pad[0] = r.pad[1] = 0;
        const int bin = bin0 + t;
        if (bin < bins) {
            const RoiXform X = sX;
            const int ph = bin / p.PW, pw = bin - ph * p.PW;
            const bool batch_ok = (X.batch >= 0) & (X.batch < p.B);
            const BinTaps g = bin_taps(X, ph, pw, p.H, p.W, (float)(p.W - 1), (float)(p.H - 1), batch_ok);
            const bool in = g.flags & BIN_IN, two_c = g.flags & TWO_COLS, two_r = g.flags & TWO_ROWS;
            r.pix = (int)(((unsigned)(batch_ok ? X.batch : 0) * (unsigned)p.H + (unsigned)g.t) * (unsigned)p.W + (unsigned)g.l);
            r.code = C_LIVE;
            if (in) {
                const bool nanw = !(fabsf(g.cx) < INFINITY) || !(fabsf(g.cy) < INFINITY);
                const bool l_lt = g.flags & TAP_LT;
                const bool l_rt = (g.flags & TAP_RT) && two_c;
                const bool l_lb = (g.flags & TAP_LB) && two_r;
                const bool l_rb = (g.flags & TAP_RB) && two_c && two_r;
                r.code |= (l_lt ? C_LT : 0u) | (l_rt ? C_RT : 0u) | (l_lb ? C_LB : 0u) | (l_rb ? C_RB : 0u);
                r.wlt = (l_lt || nanw) ? g.wlt : 0.0f;
                r.wrt = (l_rt || nanw) ? g.wrt : 0.0f;
                r.wrb = (l_rb || nanw) ? g.wrb : 0.0f;
                r.wlb = (l_lb || nanw) ? g.wlb : 0.0f;
            }
            if (p.idx_mode == IDX_COMPACT) {
                p.idx_x[(size_t)n * bins + bin] = in ? g.cx : 0.0f;
                p.idx_y[(size_t)n * bins + bin] = in ? g.cy : 0.0f;
            }
        }
        rec[t] = r;
    }
    __syncthreads();

    const int sub = lane / LPP, cvl = lane % LPP;
    const long long rowU = (long long)p.W * LPP;         // one feature row in 16-byte units
    const uint4* fbase = reinterpret_cast<const uint4*>(p.feat) + cvl;
    const int pw0 = warp * PIXW;
    uint4* obase = reinterpret_cast<uint4*>(p.out) + ((size_t)n * bins + bin0 + pw0 + sub) * LPP + cvl;
    const BinRec* rbase = rec + pw0 + sub;

#pragma unroll 1
    for (int it0 = 0; it0 < ITERS; it0 += UN) {
        uint4 lt[UN], rt[UN], lb[UN], rb[UN];
        BinRec r[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) r[u] = rbase[(it0 + u) * PPI];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const uint4* s = fbase + (long long)r[u].pix * LPP;
            const uint4* s2 = s + rowU;
            lt[u] = ldg_pred_u4(s, r[u].code & C_LT);
            rt[u] = ldg_pred_u4(s + LPP, r[u].code & C_RT);
            lb[u] = ldg_pred_u4(s2, r[u].code & C_LB);
            rb[u] = ldg_pred_u4(s2 + LPP, r[u].code & C_RB);
        }
        uint32_t any_live = 0;
#pragma unroll
        for (int u = 0; u < UN; ++u) any_live |= r[u].code;
        if (any_live & C_LIVE) {                         // own basic block: keeps the UN iterations' loads batched
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const uint64_t wlt = pair_f32(r[u].wlt, r[u].wlt), wrt = pair_f32(r[u].wrt, r[u].wrt);
                const uint64_t wrb = pair_f32(r[u].wrb, r[u].wrb), wlb = pair_f32(r[u].wlb, r[u].wlb);
                uint4 o;
                o.x = blend_bf16x2(lt[u].x, rt[u].x, rb[u].x, lb[u].x, wlt, wrt, wrb, wlb);
                o.y = blend_bf16x2(lt[u].y, rt[u].y, rb[u].y, lb[u].y, wlt, wrt, wrb, wlb);
                o.z = blend_bf16x2(lt[u].z, rt[u].z, rb[u].z, lb[u].z, wlt, wrt, wrb, wlb);
                o.w = blend_bf16x2(lt[u].w, rt[u].w, rb[u].w, lb[u].w, wlt, wrt, wrb, wlb);
                if (r[u].code & C_LIVE) obase[(size_t)(it0 + u) * PPI * LPP] = o;
            }
        }
    }
}

template <int CT>
static cudaError_t launch_fwd_nhwc_bf16_ct(FwdParams& p, cudaStream_t s, const Opts& o) {
    const bool pdl = o.pdl;
    const int bins = p.PH * p.PW;
    auto go = [&](auto kernel, int tile) {
        p.tiles = (bins + tile - 1) / tile;
        return launch_1d(kernel, (long long)p.N * p.tiles, kNhwcWarps * 32, p, s, pdl);
    };
    constexpr int PPI = 256 / CT;                        // pixels per warp iteration
    constexpr int I64 = 8 / PPI, I128 = 16 / PPI, I256 = 32 / PPI;   // per-warp iterations of a 64/128/256-bin tile
    constexpr int U64 = I64 >= 2 ? 2 : 1, U128 = I128 >= 2 ? 2 : 1, U256 = I256 >= 4 ? 4 : 2;
    const long long ctas256 = (long long)p.N * ((bins + 255) / 256);
    int variant = o.variant;                             // 0 = by grid size, like the fp32 kernel
    // one small launch: 64-bin tiles; large grids: 256-bin tiles; very large grids: whole-RoI 512-bin tiles, which halve
    // the per-CTA prologues (measured +7 % at 2 048 RoIs, but -20 % on a 64-RoI launch)
    const long long load = ctas256 * (o.concurrency > 1 ? o.concurrency : 1);     // launches the caller overlaps with this one
    if (variant == 0) variant = load < 148 * 4 ? 1 : ctas256 < 148 * 16 ? 5 : 7;
    // Measured on B200 (profiles/r01_sweep_bf16.txt): occupancy is what matters once the bytes per bin are halved --
    // 256-bin tiles, 2 iterations in flight, 64 registers (4 CTAs/SM) beat every wider-unrolled shape.
    switch (variant) {
        case 1:  return go(rroi_fwd_nhwc_bf16_kernel<CT, 64, U64, 1>, 64);
        case 2:  return go(rroi_fwd_nhwc_bf16_kernel<CT, 128, U128, 3>, 128);
        case 3:  return go(rroi_fwd_nhwc_bf16_kernel<CT, 128, (I128 >= 4 ? 4 : U128), 2>, 128);
        case 4:  return go(rroi_fwd_nhwc_bf16_kernel<CT, 256, U256, 2>, 256);
        case 6:  return go(rroi_fwd_nhwc_bf16_kernel<CT, 256, 2, 3>, 256);
        case 7:  return go(rroi_fwd_nhwc_bf16_kernel<CT, 512, 2, 4>, 512);
        default: return go(rroi_fwd_nhwc_bf16_kernel<CT, 256, 2, 4>, 256);
    }
}

cudaError_t launch_fwd_nhwc_bf16(const FwdParams& p0, const Opts& o, cudaStream_t s) {
    FwdParams p = p0;
    p.cgroups = 1;
    switch (p.C) {
        case 32:  return launch_fwd_nhwc_bf16_ct<32>(p, s, o);
        case 64:  return launch_fwd_nhwc_bf16_ct<64>(p, s, o);
        case 128: return launch_fwd_nhwc_bf16_ct<128>(p, s, o);
        case 256: return launch_fwd_nhwc_bf16_ct<256>(p, s, o);
        default:  return cudaErrorInvalidValue;
    }
}

}  // namespace rroi
