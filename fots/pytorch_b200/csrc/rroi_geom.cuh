// rroi_geom.cuh -- per-RoI affine parameters and per-bin sample geometry of RoIRotate, written so
// that every fp32 result is BIT-IDENTICAL to the reference kernel
// (/root/reference/rroi_align/src/rroi_align_kernel.cu:58-134) compiled by nvcc 12.9 for sm_100a.
//
// The reference leaves fused-multiply-add contraction to the compiler; which a*b+c became one fma
// and which stayed mul+add decides the last bit of the projected corners, and a 1-ulp change flips
// round() at x.5.  The contraction pattern below is the one in that build's PTX, spelled with
// __fmaf_rn/__fmul_rn/__fadd_rn so that hoisting the per-RoI part out of the per-element loop
// (the reference recomputes it C*PH*PW times per RoI) cannot change it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rroi {

// Row of the [N,6] RoI tensor: [batch_idx, cx, cy, h, w, angle_deg] in input-image pixels.
struct RoiXform {
    float M00, M01, M02, M10, M11, M12;  // kernel.cu:78-84
    float rpw;                           // roi_pooled_width, kernel.cu:68
    int   batch;                         // roi_batch_ind,    kernel.cu:60 (cvt.rzi)
};

// kernel.cu:58-84.  ~150 instructions incl. one fp64 divide and libdevice sinf/cosf -- call it once
// per RoI per CTA, not per element.
__device__ __forceinline__ RoiXform roi_xform(const float* __restrict__ roi, float scale, int PH) {
    RoiXform X;
    const float r0 = __ldg(roi + 0), cx = __ldg(roi + 1), cy = __ldg(roi + 2);
    const float h = __ldg(roi + 3), w = __ldg(roi + 4), adeg = __ldg(roi + 5);
    const float PHf = (float)PH;
    const float dy = (float)__dmul_rn((double)(-PH), 0.5);
    X.batch = __float2int_rz(r0);
    const float angle = (float)__dmul_rn(__ddiv_rn((double)adeg, 180.0), 3.1415926535);
    const float rpw = __fdiv_rn(__fmul_rn(w, PHf), h);
    const float ca = cosf(angle);        // libdevice __nv_cosf / __nv_sinf, as in the reference build
    const float sa = sinf(angle);
    const float dx = __fmul_rn(rpw, -0.5f);
    const float Sx = __fdiv_rn(__fmul_rn(scale, w), rpw);
    const float Sy = __fdiv_rn(__fmul_rn(scale, h), PHf);
    X.M00 = __fmul_rn(Sx, ca);
    X.M01 = __fmul_rn(Sy, sa);
    X.M02 = __fmaf_rn(scale, cx, __fmaf_rn(dx, X.M00, __fmul_rn(X.M01, dy)));
    X.M10 = __fmul_rn(Sx, -sa);
    X.M11 = __fmul_rn(Sy, ca);
    X.M12 = __fmaf_rn(scale, cy, __fmaf_rn(X.M11, dy, __fmul_rn(dx, X.M10)));
    X.rpw = rpw;
    return X;
}

// kernel.cu:86-105: project the 4 corners of bin (pw..pw+1, ph..ph+1), take the rounded bbox clamped
// to [0, W-1] x [0, H-1] on one side each, and sample at its centre.  Wm1f/Hm1f = (float)(W-1),(H-1).
// The reference clamps in fp64 (max(round(x),0.0), min(round(x), W-1.0)); with W,H < 2^24 both
// operands are exactly representable in fp32 and min/max select rather than compute, so the fp32
// min.f32/max.f32 below return the same bits (same NaN and signed-zero rules as min.f64/max.f64).
__device__ __forceinline__ void bin_center(const RoiXform& X, int ph, int pw, float Wm1f, float Hm1f,
                                           float& cx, float& cy) {
    const float pwf = (float)pw, phf = (float)ph, pw1 = (float)(pw + 1), ph1 = (float)(ph + 1);
    const float a = __fmul_rn(X.M01, phf), c = __fmul_rn(X.M01, ph1);
    const float b = __fmul_rn(X.M10, pwf), d = __fmul_rn(X.M10, pw1);
    const float P0 = __fadd_rn(__fmaf_rn(X.M00, pwf, a), X.M02);
    const float P1 = __fadd_rn(__fmaf_rn(X.M11, phf, b), X.M12);
    const float P2 = __fadd_rn(__fmaf_rn(X.M00, pwf, c), X.M02);
    const float P3 = __fadd_rn(__fmaf_rn(X.M11, ph1, b), X.M12);
    const float P4 = __fadd_rn(__fmaf_rn(X.M00, pw1, a), X.M02);
    const float P5 = __fadd_rn(__fmaf_rn(X.M11, phf, d), X.M12);
    const float P6 = __fadd_rn(__fmaf_rn(X.M00, pw1, c), X.M02);
    const float P7 = __fadd_rn(__fmaf_rn(X.M11, ph1, d), X.M12);
    const float minx = fminf(fminf(P0, P2), fminf(P4, P6));
    const float maxx = fmaxf(fmaxf(P0, P2), fmaxf(P4, P6));
    const float miny = fminf(fminf(P1, P3), fminf(P5, P7));
    const float maxy = fmaxf(fmaxf(P1, P3), fmaxf(P5, P7));
    const float L = fmaxf(roundf(minx), 0.0f);
    const float R = fminf(roundf(maxx), Wm1f);
    const float T = fmaxf(roundf(miny), 0.0f);
    const float Bm = fminf(roundf(maxy), Hm1f);
    cx = __fmul_rn(__fadd_rn(L, R), 0.5f);
    cy = __fmul_rn(__fadd_rn(T, Bm), 0.5f);
}

// Everything a channel loop needs to sample one bin.  `flags` bits:
//   0..3  tap (lt, rt, lb, rb) passes the reference's border test (kernel.cu:116-126: y>0, x>0, y<H, x<W)
//   4     r != l (two distinct columns)      5     b != t (two distinct rows)
//   6     bin is inside the RoI (pw <= rpw, kernel.cu:107) and its RoI's batch index is in range
enum : uint32_t { TAP_LT = 1u, TAP_RT = 2u, TAP_LB = 4u, TAP_RB = 8u, TWO_COLS = 16u, TWO_ROWS = 32u, BIN_IN = 64u };

struct BinTaps {
    float    cx, cy;              // sample point (what the reference stores in con_idx_x/con_idx_y)
    float    wlt, wrt, wrb, wlb;  // kernel.cu:131-134
    int      l, t;                // floor(cx), floor(cy)
    uint32_t flags;
};

// Bilinear weights, kernel.cu:128-134.  The reference forms the products in fp64 and rounds to fp32.
// The forward's sample point is (L+R)/2 with L,R integers (or +-inf), so rx,ry are 0, 0.5 or NaN and
// every product is exact in fp32 as well: identical bits, no fp64 pipe.
__device__ __forceinline__ void weights_half_grid(float cx, float cy, float& wlt, float& wrt, float& wrb, float& wlb) {
    const float rx = __fsub_rn(cx, floorf(cx));
    const float ry = __fsub_rn(cy, floorf(cy));
    const float ix = __fsub_rn(1.0f, rx), iy = __fsub_rn(1.0f, ry);
    wlt = __fmul_rn(ix, iy);
    wrt = __fmul_rn(iy, rx);
    wrb = __fmul_rn(rx, ry);
    wlb = __fmul_rn(ix, ry);
}

// Same, for an arbitrary saved centre (legacy backward reads con_idx_* from the caller): fp64 products
// exactly as kernel.cu:245-251 compiles.
__device__ __forceinline__ void weights_f64(float cx, float cy, float& wlt, float& wrt, float& wrb, float& wlb) {
    const float rx = __fsub_rn(cx, floorf(cx));
    const float ry = __fsub_rn(cy, floorf(cy));
    const double drx = (double)rx, dry = (double)ry;
    const double ix = __dsub_rn(1.0, drx), iy = __dsub_rn(1.0, dry);
    wlt = (float)__dmul_rn(ix, iy);
    wrt = (float)__dmul_rn(iy, drx);
    wrb = __fmul_rn(rx, ry);
    wlb = (float)__dmul_rn(ix, dry);
}

__device__ __forceinline__ BinTaps bin_taps(const RoiXform& X, int ph, int pw, int H, int W,
                                            float Wm1f, float Hm1f, bool batch_ok) {
    BinTaps g;
    bin_center(X, ph, pw, Wm1f, Hm1f, g.cx, g.cy);
    const int l = __float2int_rz(floorf(g.cx)), r = __float2int_rz(ceilf(g.cx));
    const int t = __float2int_rz(floorf(g.cy)), b = __float2int_rz(ceilf(g.cy));
    g.l = l; g.t = t;
    const bool xl = (l > 0) & (l < W), xr = (r > 0) & (r < W);
    const bool yt = (t > 0) & (t < H), yb = (b > 0) & (b < H);
    uint32_t f = 0;
    f |= (xl & yt) ? TAP_LT : 0u;
    f |= (xr & yt) ? TAP_RT : 0u;
    f |= (xl & yb) ? TAP_LB : 0u;
    f |= (xr & yb) ? TAP_RB : 0u;
    f |= (r != l) ? TWO_COLS : 0u;
    f |= (b != t) ? TWO_ROWS : 0u;
    f |= (((float)pw <= X.rpw) & batch_ok) ? BIN_IN : 0u;
    g.flags = f;
    weights_half_grid(g.cx, g.cy, g.wlt, g.wrt, g.wrb, g.wlb);
    return g;
}

// kernel.cu:136-141: the four products are accumulated lt, rt, rb, lb, each one fma.
__device__ __forceinline__ float blend(float lt, float rt, float rb, float lb, const BinTaps& g) {
    float v = __fmaf_rn(lt, g.wlt, 0.0f);
    v = __fmaf_rn(rt, g.wrt, v);
    v = __fmaf_rn(g.wrb, rb, v);
    v = __fmaf_rn(lb, g.wlb, v);
    return v;
}

}  // namespace rroi
