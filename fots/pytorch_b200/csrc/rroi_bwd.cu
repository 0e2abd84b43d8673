// rroi_bwd.cu -- RoIRotate backward for sm_100a.  Replaces RROIAlignBackward
// (/root/reference/rroi_align/src/rroi_align_kernel.cu:193-278): every valid output element scatters
// top_diff * {wlt, wrt, wrb, wlb} to the <=4 pixels around its saved sample centre, each only if
// 0 < y < H-1 and 0 < x < W-1 (kernel.cu:267-274 -- stricter than the forward's border test).
//
// What changed relative to the reference's four float atomics per output element:
//   * bin geometry once per (n,ph,pw), reused over channels; centres read from the compact
//     [N,PH,PW] tensor the forward saved, or recomputed from the RoI row when none was kept;
//   * NCHW: lanes of a warp hold consecutive bins along pw; neighbours that landed on the same
//     sample centre (bin pitch < 1 feature pixel) are summed with a segmented warp-shuffle
//     reduction and only the run head issues RED.ADD -- fewer, less contended L2 atomics;
//   * zero-weight taps (rx == 0 or ry == 0: the reference adds 0.0 to the same pixel) are skipped;
//   * channels-last: one 128-bit vector reduction (red.global.add.v4.f32, sm_90+) per tap and
//     4-channel group instead of four scalar atomics.
// fp32 accumulation order is unspecified exactly as in the reference, so parity is to 1e-4 rel.
#include "rroi_geom.cuh"
#include "rroi_kernels.cuh"
#include <limits.h>

namespace rroi {

// What a thread needs to scatter one bin.
struct ScatterGeom {
    float wlt, wrt, wrb, wlb;
    int   l, t, r, b;
    bool  in;         // bin inside the RoI (kernel.cu:238: skipped only if rpw < pw)
    bool  p_lt, p_rt, p_rb, p_lb;
    float cx, cy;
};

// kernel.cu:236-274 for one (n, ph, pw), centre either loaded or recomputed.
template <bool kF64Weights>
__device__ __forceinline__ ScatterGeom scatter_geom(float cx, float cy, bool in, int H, int W) {
    ScatterGeom g;
    g.cx = cx; g.cy = cy; g.in = in;
    if (kF64Weights) weights_f64(cx, cy, g.wlt, g.wrt, g.wrb, g.wlb);
    else             weights_half_grid(cx, cy, g.wlt, g.wrt, g.wrb, g.wlb);
    g.l = __float2int_rz(floorf(cx)); g.r = __float2int_rz(ceilf(cx));
    g.t = __float2int_rz(floorf(cy)); g.b = __float2int_rz(ceilf(cy));
    const bool xl = (g.l > 0) & (g.l < W - 1), xr = (g.r > 0) & (g.r < W - 1);
    const bool yt = (g.t > 0) & (g.t < H - 1), yb = (g.b > 0) & (g.b < H - 1);
    g.p_lt = in & xl & yt;
    g.p_rt = in & xr & yt;
    g.p_rb = in & xr & yb;
    g.p_lb = in & xl & yb;
    return g;
}

__device__ __forceinline__ void red_add(float* addr, float v) { atomicAdd(addr, v); }

// Chunked zero-fill + scatter (launch_bwd_zero_scatter): one scatter launch only handles the RoIs whose image lies in
// [img_lo, img_hi); CTAs of other RoIs leave at once (CTA-uniform: every thread reads the same word).
__device__ __forceinline__ bool in_image_window(const BwdParams& p, int n) {
    const int b = __float2int_rz(__ldg(p.rois + (size_t)n * 6));
    return b >= p.img_lo && b < p.img_hi;
}

// ------------------------------------------------------------------------------------------ NCHW
constexpr int kBlock = 256;

template <int CG, bool kDedupe>
__global__ void __launch_bounds__(kBlock) rroi_bwd_nchw_kernel(const BwdParams p) {
    __shared__ RoiXform sX;
    int item = blockIdx.x;
    const int tile = item % p.tiles;  item /= p.tiles;
    const int cg = item % p.cgroups;
    const int n = item / p.cgroups;
    const int bins = p.PH * p.PW;

    pdl_wait();
    pdl_launch_dependents();
    if (!in_image_window(p, n)) return;
    if (threadIdx.x < 32) {
        RoiXform X;
        if (p.idx_mode == IDX_NONE) {
            X = roi_xform(p.rois + (size_t)n * 6, p.scale, p.PH);
        } else {  // centres are loaded: only the batch index and the width mask are needed
            const float* roi = p.rois + (size_t)n * 6;
            X.batch = __float2int_rz(__ldg(roi));
            X.rpw = __fdiv_rn(__fmul_rn(__ldg(roi + 4), (float)p.PH), __ldg(roi + 3));
        }
        if (threadIdx.x == 0) sX = X;
    }
    __syncthreads();
    const RoiXform X = sX;

    const int bin = tile * kBlock + threadIdx.x;
    const bool live = bin < bins;
    const int ph = live ? bin / p.PW : 0, pw = live ? bin - ph * p.PW : 0;
    const bool batch_ok = (X.batch >= 0) & (X.batch < p.B);
    float cx = 0.f, cy = 0.f;
    if (live) {
        if (p.idx_mode == IDX_NONE) bin_center(X, ph, pw, (float)(p.W - 1), (float)(p.H - 1), cx, cy);
        else { cx = __ldg(p.idx_x + (size_t)n * bins + bin); cy = __ldg(p.idx_y + (size_t)n * bins + bin); }
    }
    const bool in = live & batch_ok & !(X.rpw < (float)pw);
    const ScatterGeom g = scatter_geom<false>(cx, cy, in, p.H, p.W);

    // Segmented-reduction plan over the warp: runs of consecutive lanes with the same centre.
    const unsigned lane = threadIdx.x & 31u;
    unsigned step_mask = 0;   // bit k: lane + 2^k is in my run
    bool head = in;
    if (kDedupe) {
        const unsigned kx = __float_as_uint(cx), ky = __float_as_uint(cy);
        const unsigned px = __shfl_up_sync(0xffffffffu, kx, 1), py = __shfl_up_sync(0xffffffffu, ky, 1);
        const bool pin = __shfl_up_sync(0xffffffffu, (int)in, 1);
        const bool same_prev = (lane > 0) & in & pin & (px == kx) & (py == ky);
        if (__ballot_sync(0xffffffffu, same_prev) != 0u) {       // warp-uniform: any run longer than 1?
            const unsigned breaks = __ballot_sync(0xffffffffu, !same_prev);
            const int run = __popc(breaks & (0xffffffffu >> (31u - lane)));
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const int other = __shfl_down_sync(0xffffffffu, run, 1 << k);
                if ((lane + (1u << k)) < 32u && other == run) step_mask |= 1u << k;
            }
            head = in & !same_prev;
        }
    }
    const bool any_merge = kDedupe && (__ballot_sync(0xffffffffu, step_mask != 0u) != 0u);

    const int c0 = cg * CG;
    const size_t HW = (size_t)p.H * p.W;
    const float* gsrc = p.top_diff + ((size_t)n * p.C + c0) * bins + bin;
    float* dst = p.bottom_diff + ((size_t)(batch_ok ? X.batch : 0) * p.C + c0) * HW;
    const unsigned o_lt = (unsigned)g.t * (unsigned)p.W + (unsigned)g.l;
    const unsigned o_rt = (unsigned)g.t * (unsigned)p.W + (unsigned)g.r;
    const unsigned o_rb = (unsigned)g.b * (unsigned)p.W + (unsigned)g.r;
    const unsigned o_lb = (unsigned)g.b * (unsigned)p.W + (unsigned)g.l;
    const bool s_lt = head & g.p_lt & (g.wlt != 0.0f);
    const bool s_rt = head & g.p_rt & (g.wrt != 0.0f);
    const bool s_rb = head & g.p_rb & (g.wrb != 0.0f);
    const bool s_lb = head & g.p_lb & (g.wlb != 0.0f);

    float gv[CG];
#pragma unroll
    for (int k = 0; k < CG; ++k)
        gv[k] = (in && (c0 + k) < p.C) ? __ldg(gsrc + (size_t)k * bins) : 0.0f;
    if (any_merge) {
#pragma unroll
        for (int k = 0; k < CG; ++k) {
#pragma unroll
            for (int s = 0; s < 5; ++s) {
                const float o = __shfl_down_sync(0xffffffffu, gv[k], 1 << s);
                if (step_mask & (1u << s)) gv[k] += o;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < CG; ++k) {
        if ((c0 + k) < p.C) {
            float* d = dst + (size_t)k * HW;
            if (s_lt) red_add(d + o_lt, __fmul_rn(g.wlt, gv[k]));
            if (s_rt) red_add(d + o_rt, __fmul_rn(g.wrt, gv[k]));
            if (s_rb) red_add(d + o_rb, __fmul_rn(g.wrb, gv[k]));
            if (s_lb) red_add(d + o_lb, __fmul_rn(g.wlb, gv[k]));
        }
    }
}

// ------------------------------------------------------------------------------- NCHW, row segments + gather
// The kernel above issues one 4-byte RED per tap, channel and run of bins: ~190 M reductions for cfg4's per-GPU batch, and
// the L2 retires ~0.5 reductions per ns whatever their width (the channels-last kernel moves 16 bytes with each) -- 390 us.
// Here a CTA (one RoI x 8 x 32 bins, as in rroi_fwd_nchw_rows_kernel) first TRANSPOSES its scatter: it builds the list of
// 16-byte granules (4 consecutive pixels of an image row) its taps touch -- per-row spans, a warp scan -- and, with native
// integer shared-memory atomics, a CSR table pixel slot -> (bin, weight) of the taps that land on it.  Then, per group of 8
// channels: the tile's 256 gradients per channel are staged in shared memory (coalesced along pw), and every GRANULE owner
// thread sums its four pixels' contributions in registers (no floating-point atomics in shared memory -- those compile to
// a compare-and-swap loop on sm_100) and issues ONE red.global.add.v4.f32 per channel plane: taps that coincide are merged
// before they leave the SM and a reduction carries up to four pixels.  ~5x fewer L2 reductions.
constexpr int kBRows = 96;                         // image rows a tile's footprint may span
constexpr int kBGran = 512;                        // granules of one channel plane per tile
constexpr int kBSlots = kBGran * 4;                // pixel slots
constexpr int kBCh = 8;                            // channels per stage

__global__ void __launch_bounds__(256, 4) rroi_bwd_nchw_rows_kernel(const BwdParams p) {
    __shared__ int rlo[kBRows], rhi[kBRows], goff[kBRows + 1];
    __shared__ int gsrc[kBGran];                   // plane offset (floats) of every granule
    __shared__ int start[kBSlots];                 // counts -> exclusive starts -> (after the fill) ends of the slots' entry lists
    __shared__ int2 ent[1024];                     // {weight bits, local bin} of every tap, grouped by pixel slot
    // staged gradients [channel][bin]: a warp's gather loads hit random bins of ONE channel row (few bank conflicts); the
    // [bin][channel] form with two LDS.128 per tap was measured slower (320 vs 264 us: conflicts + 82 registers)
    __shared__ __align__(16) float gs[2][kBCh][256];
    __shared__ int ylim[2], wsum[8];
    __shared__ RoiXform sX;
    const int n = blockIdx.x / p.tiles;
    const int tile = blockIdx.x - n * p.tiles;
    const int tiles_w = (p.PW + 31) / 32;
    const int ph0 = (tile / tiles_w) * 8, pw0 = (tile % tiles_w) * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bins = p.PH * p.PW;

    if (threadIdx.x < kBRows) { rlo[threadIdx.x] = INT_MAX; rhi[threadIdx.x] = INT_MIN; }
    if (threadIdx.x == 0) { ylim[0] = INT_MAX; ylim[1] = INT_MIN; }
    pdl_wait();
    pdl_launch_dependents();
    if (!in_image_window(p, n)) return;
    if (threadIdx.x < 32) {
        RoiXform X;
        if (p.idx_mode == IDX_NONE) {
            X = roi_xform(p.rois + (size_t)n * 6, p.scale, p.PH);
        } else {  // centres are loaded: only the batch index and the width mask are needed
            const float* roi = p.rois + (size_t)n * 6;
            X.batch = __float2int_rz(__ldg(roi));
            X.rpw = __fdiv_rn(__fmul_rn(__ldg(roi + 4), (float)p.PH), __ldg(roi + 3));
        }
        if (threadIdx.x == 0) sX = X;
    }
    __syncthreads();
    const RoiXform X = sX;
    const bool batch_ok = (X.batch >= 0) & (X.batch < p.B);

    // ---- 1. geometry: thread = bin (warp = one ph row of the tile, lanes = 32 consecutive pw)
    const int ph = ph0 + warp, pw = pw0 + lane;
    const bool live = ph < p.PH && pw < p.PW;
    const int bin = ph * p.PW + pw;
    float cx = 0.f, cy = 0.f;
    if (live) {
        if (p.idx_mode == IDX_NONE) bin_center(X, ph, pw, (float)(p.W - 1), (float)(p.H - 1), cx, cy);
        else { cx = __ldg(p.idx_x + (size_t)n * bins + bin); cy = __ldg(p.idx_y + (size_t)n * bins + bin); }
    }
    const bool in = live & batch_ok & !(X.rpw < (float)pw);
    const ScatterGeom g = scatter_geom<false>(cx, cy, in, p.H, p.W);
    const bool s_lt = g.p_lt & (g.wlt != 0.0f), s_rt = g.p_rt & (g.wrt != 0.0f);
    const bool s_rb = g.p_rb & (g.wrb != 0.0f), s_lb = g.p_lb & (g.wlb != 0.0f);
    const bool top_used = s_lt | s_rt, bot_used = s_rb | s_lb;
    {
        int ya = INT_MAX, yb = INT_MIN;
        if (top_used) { ya = g.t; yb = g.t; }
        if (bot_used) { ya = min(ya, g.b); yb = max(yb, g.b); }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            ya = min(ya, __shfl_xor_sync(0xffffffffu, ya, m));
            yb = max(yb, __shfl_xor_sync(0xffffffffu, yb, m));
        }
        if (lane == 0 && yb >= ya) { atomicMin(&ylim[0], ya); atomicMax(&ylim[1], yb); }
    }
    __syncthreads();
    const int y0 = ylim[0], nrows = ylim[1] >= ylim[0] ? ylim[1] - ylim[0] + 1 : 0;
    if (nrows == 0) return;                                                        // nothing to scatter (CTA-uniform)
    const bool aligned = (p.W % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.bottom_diff) & 15) == 0);
    bool staged = aligned && nrows <= kBRows;                                      // CTA-uniform
    if (staged) {
        // ---- 2. per-row spans of the scattered taps (scattered taps satisfy 0 < x < W-1, 0 < y < H-1: valid indices)
        if (top_used) {
            atomicMin(&rlo[g.t - y0], s_lt ? g.l : g.r); atomicMax(&rhi[g.t - y0], s_rt ? g.r : g.l);
        }
        if (bot_used) {
            atomicMin(&rlo[g.b - y0], s_lb ? g.l : g.r); atomicMax(&rhi[g.b - y0], s_rb ? g.r : g.l);
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
            if (r < kBRows) goff[r] = run;
            run += len[k];
        }
        if (lane == 31) goff[kBRows] = incl;      // total granules of one channel plane
    }
    __syncthreads();
    const int gtot = staged ? goff[kBRows] : 0;
    staged = staged && gtot > 0 && gtot <= kBGran;

    const size_t HW = (size_t)p.H * p.W;
    const float* gsrc_bin = p.top_diff + (size_t)n * p.C * bins + bin;             // + c * bins
    float* plane0 = p.bottom_diff + (size_t)(batch_ok ? X.batch : 0) * p.C * HW;

    if (!staged) {
        // ---- fallback (footprint too large or unaligned rows): one reduction per tap and channel
        if (!in) return;
        const unsigned o_lt = (unsigned)g.t * (unsigned)p.W + (unsigned)g.l, o_rt = (unsigned)g.t * (unsigned)p.W + (unsigned)g.r;
        const unsigned o_rb = (unsigned)g.b * (unsigned)p.W + (unsigned)g.r, o_lb = (unsigned)g.b * (unsigned)p.W + (unsigned)g.l;
#pragma unroll 1
        for (int c = 0; c < p.C; ++c) {
            const float gv = __ldg(gsrc_bin + (size_t)c * bins);
            float* d = plane0 + (size_t)c * HW;
            if (s_lt) red_add(d + o_lt, __fmul_rn(g.wlt, gv));
            if (s_rt) red_add(d + o_rt, __fmul_rn(g.wrt, gv));
            if (s_rb) red_add(d + o_rb, __fmul_rn(g.wrb, gv));
            if (s_lb) red_add(d + o_lb, __fmul_rn(g.wlb, gv));
        }
        return;
    }

    // ---- 3. granule table and the CSR table slot -> taps
    const int nslots = gtot * 4;
    for (int r = threadIdx.x; r < nrows; r += 256) {
        if (rhi[r] < rlo[r]) continue;
        const int a = rlo[r] & ~3, cnt = ((rhi[r] | 3) - a + 1) >> 2, base = (y0 + r) * p.W + a;
        for (int k = 0; k < cnt; ++k) gsrc[goff[r] + k] = base + 4 * k;
    }
    for (int i = threadIdx.x; i < nslots; i += 256) start[i] = 0;
    __syncthreads();
    int sl_lt = 0, sl_rt = 0, sl_rb = 0, sl_lb = 0;
    if (top_used) {
        const int o = goff[g.t - y0] * 4 - (rlo[g.t - y0] & ~3);
        sl_lt = o + g.l; sl_rt = o + g.r;
    }
    if (bot_used) {
        const int o = goff[g.b - y0] * 4 - (rlo[g.b - y0] & ~3);
        sl_lb = o + g.l; sl_rb = o + g.r;
    }
    if (s_lt) atomicAdd(&start[sl_lt], 1);
    if (s_rt) atomicAdd(&start[sl_rt], 1);
    if (s_rb) atomicAdd(&start[sl_rb], 1);
    if (s_lb) atomicAdd(&start[sl_lb], 1);
    __syncthreads();
    {   // block-wide exclusive scan of start[0 .. nslots): 8 consecutive slots per thread
        int v[8], sum = 0;
        const int i0 = threadIdx.x * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) { v[k] = (i0 + k) < nslots ? start[i0 + k] : 0; sum += v[k]; }
        int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int base = incl - sum;
        for (int w = 0; w < warp; ++w) base += wsum[w];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if ((i0 + k) < nslots) start[i0 + k] = base;
            base += v[k];
        }
    }
    __syncthreads();
    const int tl = threadIdx.x;                                                    // local bin index of this thread
    if (s_lt) ent[atomicAdd(&start[sl_lt], 1)] = make_int2(__float_as_int(g.wlt), tl);
    if (s_rt) ent[atomicAdd(&start[sl_rt], 1)] = make_int2(__float_as_int(g.wrt), tl);
    if (s_rb) ent[atomicAdd(&start[sl_rb], 1)] = make_int2(__float_as_int(g.wrb), tl);
    if (s_lb) ent[atomicAdd(&start[sl_lb], 1)] = make_int2(__float_as_int(g.wlb), tl);
    // start[s] is now the END of slot s; its entries begin at start[s - 1] (0 for s = 0)

    // ---- 4. channel stages: gradients of the tile -> shared memory -> per-granule sums -> one vector reduction per plane
    const int nst = (p.C + kBCh - 1) / kBCh;
    float gv[kBCh];
    auto load_stage = [&](int st) {
        const int c0 = st * kBCh;
#pragma unroll
        for (int ci = 0; ci < kBCh; ++ci) gv[ci] = (in && (c0 + ci) < p.C) ? __ldg(gsrc_bin + (size_t)(c0 + ci) * bins) : 0.0f;
    };
    auto store_stage = [&](int buf) {
#pragma unroll
        for (int ci = 0; ci < kBCh; ++ci) gs[buf][ci][tl] = gv[ci];
    };
    load_stage(0);
    store_stage(0);
    for (int st = 0; st < nst; ++st) {
        __syncthreads();                          // gs[st & 1] is complete (and, first time round, so is the CSR table)
        if (st + 1 < nst) load_stage(st + 1);     // in flight during the gather below
        const int c0 = st * kBCh, cn = min(kBCh, p.C - c0);
        const float (*gb)[256] = gs[st & 1];
        for (int gi = threadIdx.x; gi < gtot; gi += 256) {
            float acc[kBCh][4];
#pragma unroll
            for (int ci = 0; ci < kBCh; ++ci) acc[ci][0] = acc[ci][1] = acc[ci][2] = acc[ci][3] = 0.0f;
            int e = gi > 0 ? start[gi * 4 - 1] : 0;
            bool any = false;
#pragma unroll
            for (int px = 0; px < 4; ++px) {
                const int e1 = start[gi * 4 + px];
                any |= e1 > e;
                for (; e < e1; ++e) {
                    const int2 en = ent[e];
                    const float w = __int_as_float(en.x);
#pragma unroll
                    for (int ci = 0; ci < kBCh; ++ci) acc[ci][px] = __fmaf_rn(w, gb[ci][en.y], acc[ci][px]);
                }
            }
            if (any) {
                float* d = plane0 + (size_t)c0 * HW + gsrc[gi];
#pragma unroll
                for (int ci = 0; ci < kBCh; ++ci)
                    if (ci < cn)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d + (size_t)ci * HW), "f"(acc[ci][0]), "f"(acc[ci][1]),
                                     "f"(acc[ci][2]), "f"(acc[ci][3]) : "memory");
            }
        }
        if (st + 1 < nst) store_stage((st + 1) & 1);
    }
}

template <bool kDedupe>
static cudaError_t launch_bwd_nchw_cg(const BwdParams& p, int cg, long long grid, cudaStream_t s, bool pdl) {
    switch (cg) {
        case 1:  return launch_1d(rroi_bwd_nchw_kernel<1, kDedupe>, grid, kBlock, p, s, pdl);
        case 2:  return launch_1d(rroi_bwd_nchw_kernel<2, kDedupe>, grid, kBlock, p, s, pdl);
        case 4:  return launch_1d(rroi_bwd_nchw_kernel<4, kDedupe>, grid, kBlock, p, s, pdl);
        case 16: return launch_1d(rroi_bwd_nchw_kernel<16, kDedupe>, grid, kBlock, p, s, pdl);
        default: return launch_1d(rroi_bwd_nchw_kernel<8, kDedupe>, grid, kBlock, p, s, pdl);
    }
}

static int pick_cg(int want, int C) {
    int cg = (want == 1 || want == 2 || want == 4 || want == 8 || want == 16) ? want : 8;
    while (cg > 1 && cg / 2 >= C) cg /= 2;
    return cg;
}

cudaError_t launch_bwd_nchw(const BwdParams& p0, const Opts& o, cudaStream_t s) {
    BwdParams p = p0;
    const int cg = pick_cg(o.nchw_cg, p.C);
    const int bins = p.PH * p.PW;
    p.tiles = (bins + kBlock - 1) / kBlock;
    p.cgroups = (p.C + cg - 1) / cg;
    const long long grid = (long long)p.N * p.cgroups * p.tiles;
    const bool pdl = o.pdl;
    // Automatic: row segments + gather (per-tile fallback when rows are unaligned) once its 256-bin CTAs fill the machine a few
    // times over -- measured (profiles/r02_sweep_bwd_nchw.txt) 264 vs 468 us at 2 048 RoIs, 90 vs 120 us at 512, but 41 vs 21 us
    // for 64 RoIs alone, where 128 CTAs each walk a chain of set-up phases and 8 channel stages.
    const long long rows_ctas = (long long)p.N * ((p.PH + 7) / 8) * ((p.PW + 31) / 32);
    if (o.bwd_mode == 4 || (o.bwd_mode == 0 && rows_ctas >= 4 * 148)) {
        p.tiles = ((p.PH + 7) / 8) * ((p.PW + 31) / 32);
        if ((long long)p.H * p.W * 4 < (1LL << 31))                               // granule offsets are 32-bit
            return launch_1d(rroi_bwd_nchw_rows_kernel, (long long)p.N * p.tiles, 256, p, s, pdl);
        p.tiles = (bins + kBlock - 1) / kBlock;
    }
    return o.bwd_mode != 3 ? launch_bwd_nchw_cg<true>(p, cg, grid, s, pdl)
                           : launch_bwd_nchw_cg<false>(p, cg, grid, s, pdl);
}

// ------------------------------------------------------------------------------------------ legacy
// Reference-layout backward that, like kernel.cu:232-233, reads the centre of EVERY (n,c,ph,pw)
// element from caller-supplied [N,C,PH,PW] tensors (fp64 weight products, all four adds issued).
// One thread per output element, pw fastest; only used by RROIAlignBackwardLaucher.
__global__ void __launch_bounds__(kBlock) rroi_bwd_legacy_kernel(const BwdParams p) {
    const size_t bins = (size_t)p.PH * p.PW;
    const size_t total = (size_t)p.N * p.C * bins;
    pdl_wait();
    pdl_launch_dependents();
    for (size_t index = (size_t)blockIdx.x * kBlock + threadIdx.x; index < total;
         index += (size_t)gridDim.x * kBlock) {
        const int pw = (int)(index % p.PW);
        const size_t nc = index / bins;
        const int c = (int)(nc % p.C);
        const size_t n = nc / p.C;
        const float* roi = p.rois + n * 6;
        const int batch = __float2int_rz(__ldg(roi));
        const float rpw = __fdiv_rn(__fmul_rn(__ldg(roi + 4), (float)p.PH), __ldg(roi + 3));
        if (rpw < (float)pw) continue;
        if (batch < 0 || batch >= p.B) continue;
        const ScatterGeom g = scatter_geom<true>(__ldg(p.idx_x + index), __ldg(p.idx_y + index), true, p.H, p.W);
        const float gv = __ldg(p.top_diff + index);
        float* d = p.bottom_diff + ((size_t)batch * p.C + c) * p.H * p.W;
        if (g.p_lt) red_add(d + (size_t)g.t * p.W + g.l, __fmul_rn(g.wlt, gv));
        if (g.p_rt) red_add(d + (size_t)g.t * p.W + g.r, __fmul_rn(g.wrt, gv));
        if (g.p_rb) red_add(d + (size_t)g.b * p.W + g.r, __fmul_rn(g.wrb, gv));
        if (g.p_lb) red_add(d + (size_t)g.b * p.W + g.l, __fmul_rn(g.wlb, gv));
    }
}

cudaError_t launch_bwd_legacy(const BwdParams& p, const Opts& o, cudaStream_t s) {
    const size_t total = (size_t)p.N * p.C * p.PH * p.PW;
    long long grid = (long long)((total + kBlock - 1) / kBlock);
    if (grid > (1LL << 30)) grid = 1LL << 30;
    return launch_1d(rroi_bwd_legacy_kernel, grid, kBlock, p, s, o.pdl);
}

// ------------------------------------------------------------------------------------------ NHWC
// top_diff [N,PH,PW,C], bottom_diff [B,H,W,C].  Same two-phase shape as the NHWC forward.
constexpr int kTilePix = 64;

struct __align__(16) PixScatter {
    long long base;                 // (batch*H)*W pixel index of the image origin
    int l, t, r, b;
    float wlt, wrt, wrb, wlb;
    uint32_t pred;                  // bit0 lt, bit1 rt, bit2 rb, bit3 lb (border test & weight != 0)
};

__device__ __forceinline__ void red_add_v4(float* addr, float w, const float4& g) {
    const float a = __fmul_rn(w, g.x), b = __fmul_rn(w, g.y), c = __fmul_rn(w, g.z), d = __fmul_rn(w, g.w);
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <bool kVec>
__global__ void __launch_bounds__(kBlock) rroi_bwd_nhwc_kernel(const BwdParams p) {
    __shared__ RoiXform sX;
    __shared__ PixScatter rec[kTilePix];
    const int n = blockIdx.x / p.tiles;
    const int tile = blockIdx.x - n * p.tiles;
    const int bins = p.PH * p.PW;
    const int bin0 = tile * kTilePix;
    const int npix = min(kTilePix, bins - bin0);

    pdl_wait();
    pdl_launch_dependents();
    if (!in_image_window(p, n)) return;
    if (threadIdx.x < 32) {
        RoiXform X;
        if (p.idx_mode == IDX_NONE) {
            X = roi_xform(p.rois + (size_t)n * 6, p.scale, p.PH);
        } else {
            const float* roi = p.rois + (size_t)n * 6;
            X.batch = __float2int_rz(__ldg(roi));
            X.rpw = __fdiv_rn(__fmul_rn(__ldg(roi + 4), (float)p.PH), __ldg(roi + 3));
        }
        if (threadIdx.x == 0) sX = X;
    }
    __syncthreads();
    if (threadIdx.x < npix) {
        const RoiXform X = sX;
        const int bin = bin0 + threadIdx.x;
        const int ph = bin / p.PW, pw = bin - ph * p.PW;
        const bool batch_ok = (X.batch >= 0) & (X.batch < p.B);
        float cx, cy;
        if (p.idx_mode == IDX_NONE) bin_center(X, ph, pw, (float)(p.W - 1), (float)(p.H - 1), cx, cy);
        else { cx = __ldg(p.idx_x + (size_t)n * bins + bin); cy = __ldg(p.idx_y + (size_t)n * bins + bin); }
        const bool in = batch_ok & !(X.rpw < (float)pw);
        const ScatterGeom g = scatter_geom<false>(cx, cy, in, p.H, p.W);
        PixScatter r;
        r.base = (long long)(batch_ok ? X.batch : 0) * p.H * p.W;
        r.l = g.l; r.t = g.t; r.r = g.r; r.b = g.b;
        r.wlt = g.wlt; r.wrt = g.wrt; r.wrb = g.wrb; r.wlb = g.wlb;
        r.pred = ((g.p_lt & (g.wlt != 0.f)) ? 1u : 0u) | ((g.p_rt & (g.wrt != 0.f)) ? 2u : 0u) |
                 ((g.p_rb & (g.wrb != 0.f)) ? 4u : 0u) | ((g.p_lb & (g.wlb != 0.f)) ? 8u : 0u);
        rec[threadIdx.x] = r;
    }
    __syncthreads();

    constexpr int VN = kVec ? 4 : 1;
    const int CV = p.C / VN;
    const int units = npix * CV;
    const float* gsrc = p.top_diff + ((size_t)n * bins + bin0) * p.C;
    for (int u = threadIdx.x; u < units; u += kBlock) {
        const int px = u / CV, v = u - px * CV;
        const PixScatter r = rec[px];
        if (r.pred == 0u) continue;
        float* base = p.bottom_diff + (size_t)r.base * p.C + (size_t)v * VN;
        const size_t o_lt = ((size_t)r.t * p.W + r.l) * p.C, o_rt = ((size_t)r.t * p.W + r.r) * p.C;
        const size_t o_rb = ((size_t)r.b * p.W + r.r) * p.C, o_lb = ((size_t)r.b * p.W + r.l) * p.C;
        if (kVec) {
            const float4 gq = __ldg(reinterpret_cast<const float4*>(gsrc + (size_t)u * 4));
            if (r.pred & 1u) red_add_v4(base + o_lt, r.wlt, gq);
            if (r.pred & 2u) red_add_v4(base + o_rt, r.wrt, gq);
            if (r.pred & 4u) red_add_v4(base + o_rb, r.wrb, gq);
            if (r.pred & 8u) red_add_v4(base + o_lb, r.wlb, gq);
        } else {
            const float gq = __ldg(gsrc + u);
            if (r.pred & 1u) red_add(base + o_lt, __fmul_rn(r.wlt, gq));
            if (r.pred & 2u) red_add(base + o_rt, __fmul_rn(r.wrt, gq));
            if (r.pred & 4u) red_add(base + o_rb, __fmul_rn(r.wrb, gq));
            if (r.pred & 8u) red_add(base + o_lb, __fmul_rn(r.wlb, gq));
        }
    }
}

// ------------------------------------------------------------------------------------ NHWC, packed
// Mirror of the packed forward for C in {32, 64, 128, 256}: one thread per bin computes the scatter geometry into
// a shared record (weights of skipped taps are folded to "not scattered"), then every lane owns one float4 of one
// bin per iteration: one coalesced 128-bit load of top_diff, <=4 x (4 FMUL + RED.ADD.v4.f32).  UN iterations' loads
// are issued before the first reduction.
struct __align__(16) ScatRec {
    int pix;                       // (batch*H + t)*W + l
    uint32_t pred;                 // bit0 lt, bit1 rt (pix+1), bit2 rb (pix+W+1), bit3 lb (pix+W); bit4 bin < bins
    float wlt, wrt, wrb, wlb;
    int pad[2];
};

__device__ __forceinline__ float4 ld_v4(const float* ptr, uint32_t pred) {
    float4 r;
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t"
        "mov.b32 %0, 0;\n\tmov.b32 %1, 0;\n\tmov.b32 %2, 0;\n\tmov.b32 %3, 0;\n\t"
        "@q ld.global.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
        : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
        : "l"(ptr), "r"(pred));
    return r;
}

constexpr int kPackWarps = 8;

template <int CT, int TILE, int UN>
__global__ void __launch_bounds__(kPackWarps * 32) rroi_bwd_nhwc_packed_kernel(const BwdParams p) {
    constexpr int LPP = CT / 4;
    constexpr int PPI = LPP >= 32 ? 1 : 32 / LPP;
    constexpr int NCH = LPP > 32 ? LPP / 32 : 1;
    constexpr int PIXW = TILE / kPackWarps;
    constexpr int ITERS = PIXW * NCH / PPI;
    static_assert(PIXW * kPackWarps == TILE && ITERS % UN == 0 && ITERS > 0, "tile shape");
    __shared__ RoiXform sX;
    __shared__ ScatRec rec[TILE];
    const int n = blockIdx.x / p.tiles;
    const int tile = blockIdx.x - n * p.tiles;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bins = p.PH * p.PW;
    const int bin0 = tile * TILE;

    pdl_wait();
    pdl_launch_dependents();
    if (!in_image_window(p, n)) return;
    if (warp == 0) {
        RoiXform X;
        if (p.idx_mode == IDX_NONE) {
            X = roi_xform(p.rois + (size_t)n * 6, p.scale, p.PH);
        } else {
            const float* roi = p.rois + (size_t)n * 6;
            X.batch = __float2int_rz(__ldg(roi));
            X.rpw = __fdiv_rn(__fmul_rn(__ldg(roi + 4), (float)p.PH), __ldg(roi + 3));
        }
        if (lane == 0) sX = X;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < TILE; t += kPackWarps * 32) {
        ScatRec r;
        r.pix = 0; r.pred = 0; r.wlt = r.wrt = r.wrb = r.wlb = 0.0f; r.pad[0] = r.pad[1] = 0;
        const int bin = bin0 + t;
        if (bin < bins) {
            const RoiXform X = sX;
            const int ph = bin / p.PW, pw = bin - ph * p.PW;
            const bool batch_ok = (X.batch >= 0) & (X.batch < p.B);
            float cx, cy;
            if (p.idx_mode == IDX_NONE) bin_center(X, ph, pw, (float)(p.W - 1), (float)(p.H - 1), cx, cy);
            else { cx = __ldg(p.idx_x + (size_t)n * bins + bin); cy = __ldg(p.idx_y + (size_t)n * bins + bin); }
            const bool in = batch_ok & !(X.rpw < (float)pw);
            const ScatterGeom g = scatter_geom<false>(cx, cy, in, p.H, p.W);
            r.pix = (int)(((unsigned)(batch_ok ? X.batch : 0) * (unsigned)p.H + (unsigned)g.t) * (unsigned)p.W + (unsigned)g.l);
            // a tap with non-zero weight is a distinct pixel: rt/rb only when r == l + 1, lb/rb only when b == t + 1
            r.pred = ((g.p_lt & (g.wlt != 0.f)) ? 1u : 0u) | ((g.p_rt & (g.wrt != 0.f) & (g.r == g.l + 1)) ? 2u : 0u) |
                     ((g.p_rb & (g.wrb != 0.f) & (g.r == g.l + 1) & (g.b == g.t + 1)) ? 4u : 0u) |
                     ((g.p_lb & (g.wlb != 0.f) & (g.b == g.t + 1)) ? 8u : 0u) | 16u;
            r.wlt = g.wlt; r.wrt = g.wrt; r.wrb = g.wrb; r.wlb = g.wlb;
        }
        rec[t] = r;
    }
    __syncthreads();

    const int sub = LPP >= 32 ? 0 : lane / LPP;
    const int cvl = LPP >= 32 ? lane : lane % LPP;
    const long long rowC = (long long)p.W * CT;
    float* gbase = p.bottom_diff + cvl * 4;
    const int pw0 = warp * PIXW;
    const float* tbase = p.top_diff + ((size_t)n * bins + bin0 + pw0 + (NCH > 1 ? 0 : sub)) * CT + cvl * 4;
    const ScatRec* rbase = rec + pw0 + (NCH > 1 ? 0 : sub);

#pragma unroll 1
    for (int it0 = 0; it0 < ITERS; it0 += UN) {
        float4 gq[UN];
        ScatRec r[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int it = it0 + u;
            const int dpx = NCH > 1 ? it / NCH : it * PPI;
            const int ch = NCH > 1 ? (it % NCH) * 128 : 0;
            r[u] = rbase[dpx];
            gq[u] = ld_v4(tbase + dpx * CT + ch, r[u].pred & 15u);     // bins that scatter nothing are not even read
        }
        uint32_t any = 0;
#pragma unroll
        for (int u = 0; u < UN; ++u) any |= r[u].pred;
        if (any & 15u) {
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int it = it0 + u;
                const int ch = NCH > 1 ? (it % NCH) * 128 : 0;
                float* d = gbase + (long long)r[u].pix * CT + ch;
                if (r[u].pred & 1u) red_add_v4(d, r[u].wlt, gq[u]);
                if (r[u].pred & 2u) red_add_v4(d + CT, r[u].wrt, gq[u]);
                if (r[u].pred & 4u) red_add_v4(d + rowC + CT, r[u].wrb, gq[u]);
                if (r[u].pred & 8u) red_add_v4(d + rowC, r[u].wlb, gq[u]);
            }
        }
    }
}

// EVALUATED AND REMOVED (round 2, profiles/r02_sweep_bwd2.txt): grouping a CTA's taps by pixel in a shared-memory hash
// table and issuing ONE vector reduction per distinct pixel (739 tap writes land on 386 pixels per RoI) was slower than
// the plain scatter above: 170 vs 161 us on cfg4's per-GPU batch (C = 64), 57 vs 45 us on 8 images.  Halving the
// reductions buys nothing because the scatter is DRAM-bound, not reduction-bound: it moves 171 MB of top_diff plus the
// read-modify-write of 203 MB of zeroed lines in ~85 us (6.8 TB/s, the copy roofline).

template <int CT>
static cudaError_t launch_bwd_nhwc_packed(BwdParams& p, cudaStream_t s, bool pdl) {
    const int bins = p.PH * p.PW;
    constexpr int PPI = CT >= 128 ? 1 : 128 / CT;
    constexpr int NCH = CT > 128 ? CT / 128 : 1;
    constexpr int I64 = 8 * NCH / PPI;
    const long long ctas256 = (long long)p.N * ((bins + 255) / 256);
    if (ctas256 >= 148 * 4) {
        p.tiles = (bins + 255) / 256;
        return launch_1d(rroi_bwd_nhwc_packed_kernel<CT, 256, 4>, (long long)p.N * p.tiles, kPackWarps * 32, p, s, pdl);
    }
    p.tiles = (bins + 63) / 64;
    return launch_1d(rroi_bwd_nhwc_packed_kernel<CT, 64, (I64 >= 4 ? 4 : I64)>, (long long)p.N * p.tiles, kPackWarps * 32, p, s, pdl);
}

cudaError_t launch_bwd_nhwc(const BwdParams& p0, const Opts& o, cudaStream_t s) {
    BwdParams p = p0;
    const int bins = p.PH * p.PW;
    p.tiles = (bins + kTilePix - 1) / kTilePix;
    p.cgroups = 1;
    const long long grid = (long long)p.N * p.tiles;
    const bool pdl = o.pdl;
    const bool vec = (p.C % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(p.top_diff) | reinterpret_cast<uintptr_t>(p.bottom_diff)) % 16 == 0);
    if (vec && o.bwd_mode != 2) {      // opts.bwd_mode = 2 selects the generic kernel (A/B measurements)
        if (p.C == 32)  return launch_bwd_nhwc_packed<32>(p, s, pdl);
        if (p.C == 64)  return launch_bwd_nhwc_packed<64>(p, s, pdl);
        if (p.C == 128) return launch_bwd_nhwc_packed<128>(p, s, pdl);
        if (p.C == 256) return launch_bwd_nhwc_packed<256>(p, s, pdl);
    }
    return vec ? launch_1d(rroi_bwd_nhwc_kernel<true>, grid, kBlock, p, s, pdl)
               : launch_1d(rroi_bwd_nhwc_kernel<false>, grid, kBlock, p, s, pdl);
}

// --------------------------------------------------------------------- zero-fill + scatter, chunked by image
// The gradient map must be defined everywhere, so zero_fill means 4*B*C*H*W bytes of stores in front of the scatter.
// With one memset of a map much larger than L2 (cfg3/cfg4: 472 MB vs 126 MB) every line a RoI touches crosses DRAM
// three times: written as zero, evicted, read back by the first reduction, written again.  Here the map is cleared in
// chunks of a few images on a side stream and each chunk's RoIs are scattered (image window [img_lo, img_hi) of the
// scatter kernels) as soon as ITS chunk is clear: the reductions hit lines that are still dirty in L2, and the
// zero-fill of the next chunks (DRAM-write-bound) overlaps the scatter of this one (L2-reduction-bound).  The side
// stream runs at most two chunks ahead of the scatter so the cleared lines are not evicted before they are used.
// No assumption on the order of the RoI rows: a scatter launch skips (CTA-uniform early exit) every RoI outside its
// window.  Streams and events are cached per host thread and device, so concurrent callers never share them.
namespace {

struct SideLane {
    cudaStream_t side = nullptr;
    cudaEvent_t fork = nullptr;
    cudaEvent_t zeroed[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t scattered[3] = {nullptr, nullptr, nullptr};
    bool ok = false;
};

SideLane* side_lane() {
    thread_local SideLane lanes[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    SideLane& L = lanes[dev];
    if (!L.ok) {
        if (cudaStreamCreateWithFlags(&L.side, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        bool good = cudaEventCreateWithFlags(&L.fork, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; i < 3 && good; ++i)
            good = cudaEventCreateWithFlags(&L.zeroed[i], cudaEventDisableTiming) == cudaSuccess &&
                   cudaEventCreateWithFlags(&L.scattered[i], cudaEventDisableTiming) == cudaSuccess;
        if (!good) return nullptr;
        L.ok = true;
    }
    return &L;
}

}  // namespace

cudaError_t launch_bwd_zero_scatter(const BwdParams& p0, const Opts& o, bool nhwc, cudaStream_t s) {
    BwdParams p = p0;
    const size_t img_bytes = (size_t)p.C * p.H * p.W * sizeof(float);
    const size_t map_bytes = img_bytes * (size_t)p.B;
    auto scatter = [&](int lo, int hi) {
        p.img_lo = lo; p.img_hi = hi;
        return nhwc ? launch_bwd_nhwc(p, o, s) : launch_bwd_nchw(p, o, s);
    };
    // MEASURED on B200 (profiles/r02_sweep_bwd.txt, cfg4's per-GPU batch, C = 64): one memset + one scatter 159 us; chunks
    // of 8 / 4 / 2 / 1 images 184 / 208 / 266 / 408 us -- every chunk costs about 7 us of cross-stream hand-over and the
    // scatter gains nothing from L2-resident zeros because it is bound by the reduction path, not by DRAM.  So the
    // chunked form is opt-in (opts.zero_chunk_images > 0) and the default is the plain memset + scatter.
    const int per = o.zero_chunk_images;
    SideLane* L = nullptr;
    if (per > 0 && per < p.B && map_bytes > ((size_t)96 << 20) && p.N > 0) L = side_lane();
    if (!L) {                                              // small map (stays in L2 anyway), or asked for: memset, scatter
        cudaError_t e = cudaMemsetAsync(p.bottom_diff, 0, map_bytes, s);
        if (e != cudaSuccess || p.N == 0) return e;
        return scatter(0, 0x7fffffff);
    }
    const int chunks = (p.B + per - 1) / per;
    cudaError_t e = cudaEventRecord(L->fork, s);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(L->side, L->fork, 0);
    auto clear = [&](int g) -> cudaError_t {               // on the side stream; at most two chunks ahead of the scatter
        cudaError_t ee = cudaSuccess;
        if (g >= 3) ee = cudaStreamWaitEvent(L->side, L->scattered[(g - 3) % 3], 0);
        const int lo = g * per, hi = lo + per < p.B ? lo + per : p.B;
        if (ee == cudaSuccess) ee = cudaMemsetAsync(reinterpret_cast<char*>(p.bottom_diff) + (size_t)lo * img_bytes, 0, (size_t)(hi - lo) * img_bytes, L->side);
        if (ee == cudaSuccess) ee = cudaEventRecord(L->zeroed[g % 3], L->side);
        return ee;
    };
    // Event slots are reused round-robin (3 each way); a cudaStreamWaitEvent captures the record that precedes it in host
    // order, so each wait below is issued before its slot is recorded again.
    for (int g = 0; g < chunks && g < 2 && e == cudaSuccess; ++g) e = clear(g);
    for (int g = 0; g < chunks && e == cudaSuccess; ++g) {
        e = cudaStreamWaitEvent(s, L->zeroed[g % 3], 0);
        const int lo = g * per, hi = g + 1 == chunks ? 0x7fffffff : lo + per;     // the last window also takes batch >= B (no-ops)
        if (e == cudaSuccess) e = scatter(g == 0 ? -0x7fffffff : lo, hi);
        if (e == cudaSuccess) e = cudaEventRecord(L->scattered[g % 3], s);
        if (e == cudaSuccess && g + 2 < chunks) e = clear(g + 2);
    }
    return e;
}

}  // namespace rroi
