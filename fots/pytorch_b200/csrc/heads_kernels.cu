// heads_kernels.cu -- the three EAST-style detection heads of the feeder in ONE pass over the shared feature map
// (/root/reference/tools/models.py:440-456: act = Conv2d(256, 1, 1), rbox = Conv2d(256, 4, 1), angle = Conv2d(256, 2, 1);
// seg = sigmoid(act), rbox = sigmoid(.) * 128, angle = sigmoid(.) * 2 - 1 normalised to unit length).
//
// The library path is three cuBLAS GEMMs with 1, 4 and 2 output columns -- each reads the whole 256-channel map -- plus
// about fifteen element-wise launches.  Seven output columns are exactly one N = 8 tensor-core tile: a warp takes 16
// pixels, loads their 256 channels once (16-byte vectors, 64 contiguous bytes per pixel and instruction), multiplies them
// by the 8 x 256 weight block held in registers with mma.sync.m16n8k16 (fp32 accumulation), and finishes sigmoid,
// scaling and the (sin, cos) normalisation in the accumulator registers.  HBM-bound: the map is read once, 28 bytes per
// pixel are written.  mma.sync, not tcgen05: N = 8 is below a UMMA tile and the kernel is bound by the read of x.
#include "../../../include/fots_b200_pipeline.h"
#include "pdl.cuh"
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__device__ __forceinline__ void mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }

// 32-byte load (sm_100 LDG.256): the four lanes of a pixel row then cover one whole 128-byte line per instruction -- with 16-byte
// loads a warp-level load touched 16 lines for 1 KB and the kernels were bound by L1 wavefronts (3.8 TB/s), not by HBM.
struct U8 { uint32_t r[8]; };
__device__ __forceinline__ U8 ldg256(const void* p) {
    U8 v;
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v.r[0]), "=r"(v.r[1]), "=r"(v.r[2]), "=r"(v.r[3]), "=r"(v.r[4]), "=r"(v.r[5]), "=r"(v.r[6]), "=r"(v.r[7]) : "l"(p));
    return v;
}
__device__ __forceinline__ U8 zero256() { U8 v; for (int i = 0; i < 8; ++i) v.r[i] = 0u; return v; }
// acc += A(16 channels of rows g, g + 8) . B(the same 16 channels of head g): four k16 steps' worth of register pairs
__device__ __forceinline__ void mma_u8(float (&acc)[4], const U8& a, const U8& b, const U8& w) {
    mma_16816(acc, a.r[0], b.r[0], a.r[1], b.r[1], w.r[0], w.r[1]);
    mma_16816(acc, a.r[2], b.r[2], a.r[3], b.r[3], w.r[2], w.r[3]);
    mma_16816(acc, a.r[4], b.r[4], a.r[5], b.r[5], w.r[4], w.r[5]);
    mma_16816(acc, a.r[6], b.r[6], a.r[7], b.r[7], w.r[6], w.r[7]);
}

// Squashing + store of one pixel's two columns (2t, 2t + 1) of the 8-column head block {act, 0, rbox0..3, angle0, angle1}.
__device__ __forceinline__ void heads_store(int t, float v0, float v1, long long p, int HW, float* __restrict__ seg, float* __restrict__ rbox,
                                            float* __restrict__ angle) {
    const long long b = p / HW, hw = p - b * HW;
    if (t == 0) {
        seg[p] = sigmoid_f(v0);
    } else if (t < 3) {
        float* r = rbox + (b * 4 + (t - 1) * 2) * HW + hw;
        r[0] = sigmoid_f(v0) * 128.0f;
        r[HW] = sigmoid_f(v1) * 128.0f;
    } else {
        const float s = sigmoid_f(v0) * 2.0f - 1.0f, c = sigmoid_f(v1) * 2.0f - 1.0f;
        const float nrm = sqrtf(s * s + c * c);
        float* a = angle + (b * 2) * HW + hw;
        a[0] = s / nrm;
        a[HW] = c / nrm;
    }
}

// Logical k of the MMA <-> physical channel.  Lane (g = lane / 4, t = lane % 4) loads, for load m, the 16 channels
// 64 m + 16 t .. + 15 of pixel rows g and g + 8 (the four t-lanes of a pixel read 128 contiguous bytes = one line); four k-steps
// consume them four channels at a time (k = 2t, 2t+1 | 2t+8, 2t+9).  The B fragment of lane (g, t) is head g at the same
// channels, so the permutation cancels in the dot product.
// wq: [8 heads][C] bf16, column order {act, 0, rbox0..3, angle0, angle1};  bias: [8] fp32 in the same order.
// RAW: a 1x1 convolution to ONE output channel (the attention gate of the top-down merge, tools/models.py:405-438:
// conv_attenton = Conv2d(256, 1, 1)): column 0 + bias, no squashing, written as bf16 [npix] -- the logits
// fots_b200_fpn_merge_nhwc_bf16 consumes.  Same loads and MMAs; `seg` then points at the bf16 output.
template <int C, int RAW>
__global__ void __launch_bounds__(256) heads_kernel(const uint4* __restrict__ x, const __nv_bfloat16* __restrict__ wq,
                                                    const float* __restrict__ bias, float* __restrict__ seg, float* __restrict__ rbox,
                                                    float* __restrict__ angle, long long npix, int HW) {
    constexpr int NL = C / 64;                        // 32-byte loads per pixel row and lane
    pdl::trigger();                                   // the weights / bias below are constants: loaded before pdl::wait()
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    // this lane's weights: head g, channels 64 m + 16 t .. + 15 -> eight 32-bit registers per m
    U8 wb[NL];
#pragma unroll
    for (int m = 0; m < NL; ++m) wb[m] = ldg256(wq + (size_t)g * C + m * 64 + t * 16);
    const float b0 = __ldg(bias + 2 * t), b1 = __ldg(bias + 2 * t + 1);
    pdl::wait();
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long ntiles = (npix + 15) / 16;
    const __nv_bfloat16* xe = reinterpret_cast<const __nv_bfloat16*>(x);
    for (long long tile = warp0; tile < ntiles; tile += nwarps) {
        const long long p0 = tile * 16 + g, p1 = p0 + 8;
        const bool ok0 = p0 < npix, ok1 = p1 < npix;
        U8 xa[NL], xb[NL];
#pragma unroll
        for (int m = 0; m < NL; ++m) {
            xa[m] = ok0 ? ldg256(xe + p0 * C + m * 64 + t * 16) : zero256();
            xb[m] = ok1 ? ldg256(xe + p1 * C + m * 64 + t * 16) : zero256();
        }
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int m = 0; m < NL; ++m) mma_u8(acc, xa[m], xb[m], wb[m]);
        // acc[0], acc[1] = (pixel p0, columns 2t, 2t+1); acc[2], acc[3] = (pixel p1, same columns)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const long long p = half ? p1 : p0;
            if (!(half ? ok1 : ok0)) continue;
            const float v0 = acc[2 * half] + b0, v1 = acc[2 * half + 1] + b1;
            if (RAW) {                                       // 1: logits, 2: sigmoid(logit)
                if (t == 0) reinterpret_cast<__nv_bfloat16*>(seg)[p] = __float2bfloat16_rn(RAW == 2 ? sigmoid_f(v0) : v0);
                continue;
            }
            heads_store(t, v0, v1, p, HW, seg, rbox, angle);
        }
    }
}


// ---- the last level of the top-down merge COLLAPSED into the heads (inference) --------------------------------------------
// tools/models.py:430-456 computes, at 1/4 scale,  x = upconv2_pw(d) + feature1(s) * gate  and then the three 1x1 heads on x,
// where d = upconv2_dw(upsample(f2)), s = the stage-1 output (64 channels) and gate = upsample(sigmoid(conv_attenton(f2))).
// x feeds nothing else at inference time (the recogniser reads the stage-0 map), and everything between d / s and the head
// logits is linear:   logits = (Wh Wpw) d + gate * (Wh Wf1) s + bh.   The two products are 8 x 256 and 8 x 64 matrices folded
// once at placement time (FOTSNet.to_b200), so the 256 -> 256 pointwise convolution, the 64 -> 256 lateral convolution, the
// merge kernel and the 236 MB map x itself (written once, read twice per 8 images) disappear: this kernel reads d and s once.
// Same fragment scheme as heads_kernel; the gate is interpolated per pixel with the merge kernel's arithmetic.
template <int C1, int C2>
__global__ void __launch_bounds__(256, 2) heads_dual_kernel(const uint4* __restrict__ x1, const __nv_bfloat16* __restrict__ w1,
                                                         const uint4* __restrict__ x2, const __nv_bfloat16* __restrict__ w2,
                                                         const __nv_bfloat16* __restrict__ gate, const float* __restrict__ bias,
                                                         float* __restrict__ seg, float* __restrict__ rbox, float* __restrict__ angle,
                                                         long long npix, int H, int W, int gh, int gw) {
    constexpr int N1 = C1 / 64, N2 = C2 / 64;
    pdl::trigger();
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    // weight fragments of every lane in shared memory ([m][lane] x 32 bytes, read back as two LDS.128 right before the MMAs):
    // kept in registers they push the kernel to 156 registers = one CTA per SM, too few warps for a streaming kernel
    __shared__ __align__(32) U8 wsm[(N1 + N2) * 32];
    if (threadIdx.x < 32) {
#pragma unroll
        for (int m = 0; m < N1; ++m) wsm[m * 32 + lane] = ldg256(w1 + (size_t)g * C1 + m * 64 + t * 16);
#pragma unroll
        for (int m = 0; m < N2; ++m) wsm[(N1 + m) * 32 + lane] = ldg256(w2 + (size_t)g * C2 + m * 64 + t * 16);
    }
    __syncthreads();
    const float b0 = __ldg(bias + 2 * t), b1 = __ldg(bias + 2 * t + 1);
    pdl::wait();
    const int HW = H * W;
    const float sy = H > 1 ? (float)(gh - 1) / (float)(H - 1) : 0.0f, sx = W > 1 ? (float)(gw - 1) / (float)(W - 1) : 0.0f;
    auto gate_at = [&](long long p) {                                    // bilinear, align_corners (fpn_merge's arithmetic)
        const long long b = p / HW;
        const int hw = (int)(p - b * HW), Y = hw / W, X = hw - Y * W;
        const float fy = sy * (float)Y, fx = sx * (float)X;
        const int y0 = (int)fy, x0 = (int)fx;
        const int y1 = y0 + (y0 < gh - 1 ? 1 : 0), xx1 = x0 + (x0 < gw - 1 ? 1 : 0);
        const float ly1 = fy - (float)y0, lx1 = fx - (float)x0, ly0 = 1.0f - ly1, lx0 = 1.0f - lx1;
        const __nv_bfloat16* gp = gate + b * (long long)gh * gw;
        const float s00 = __bfloat162float(gp[y0 * gw + x0]), s01 = __bfloat162float(gp[y0 * gw + xx1]);
        const float s10 = __bfloat162float(gp[y1 * gw + x0]), s11 = __bfloat162float(gp[y1 * gw + xx1]);
        return ly0 * (lx0 * s00 + lx1 * s01) + ly1 * (lx0 * s10 + lx1 * s11);
    };
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long ntiles = (npix + 15) / 16;
    const __nv_bfloat16* x1e = reinterpret_cast<const __nv_bfloat16*>(x1);
    const __nv_bfloat16* x2e = reinterpret_cast<const __nv_bfloat16*>(x2);
    for (long long tile = warp0; tile < ntiles; tile += nwarps) {
        const long long p0 = tile * 16 + g, p1 = p0 + 8;
        const bool ok0 = p0 < npix, ok1 = p1 < npix;
        U8 xa[N1], xb[N1], ya[N2], yb[N2];
#pragma unroll
        for (int m = 0; m < N1; ++m) {
            xa[m] = ok0 ? ldg256(x1e + p0 * C1 + m * 64 + t * 16) : zero256();
            xb[m] = ok1 ? ldg256(x1e + p1 * C1 + m * 64 + t * 16) : zero256();
        }
#pragma unroll
        for (int m = 0; m < N2; ++m) {
            ya[m] = ok0 ? ldg256(x2e + p0 * C2 + m * 64 + t * 16) : zero256();
            yb[m] = ok1 ? ldg256(x2e + p1 * C2 + m * 64 + t * 16) : zero256();
        }
        const float g0 = ok0 ? gate_at(p0) : 0.0f, g1 = ok1 ? gate_at(p1) : 0.0f;
        float acc[4] = {0.f, 0.f, 0.f, 0.f}, acs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int m = 0; m < N1; ++m) mma_u8(acc, xa[m], xb[m], wsm[m * 32 + lane]);
#pragma unroll
        for (int m = 0; m < N2; ++m) mma_u8(acs, ya[m], yb[m], wsm[(N1 + m) * 32 + lane]);
        if (ok0) heads_store(t, fmaf(g0, acs[0], acc[0]) + b0, fmaf(g0, acs[1], acc[1]) + b1, p0, HW, seg, rbox, angle);
        if (ok1) heads_store(t, fmaf(g1, acs[2], acc[2]) + b0, fmaf(g1, acs[3], acc[3]) + b1, p1, HW, seg, rbox, angle);
    }
}


// ---- ... and the depthwise half of upconv2 and its 2x upsampling folded in as well -------------------------------------------
// d = dw3x3(up(f2)) is linear too, and the bilinear upsampling (spatial, per channel) commutes with any per-pixel channel
// mix.  With M = Wh Wpw [8, 256] and the depthwise taps w_dw[c, tap]:
//     (M d)(p) = sum_tap  up(T_tap)(p + tap),     T_tap = (M . diag(w_dw[:, tap])) f2          (zero outside the map: the padding)
// T = [9 taps][8 heads] = 72 channels at the LOW resolution comes out of one 1x1 convolution on f2 (the tcgen05 kernel, 72
// padded to TC output channels, bf16); this kernel then gathers, per full-resolution pixel, 9 taps x 4 bilinear samples of
// 16 bytes (its eight head columns of one tap at one low-resolution pixel; served by L1 / L2) instead of reading 512 bytes
// of d.  The 256-channel map d at 1/4 scale (944 MB per 32 images, written once and read once) and the upsample-on-load
// depthwise kernel disappear.
// Phase 1: lane = pixel (32 consecutive pixels of a row per warp: neighbouring lanes share their low-resolution samples), the
// eight sums and the gate go to the warp's slice of shared memory.  Phase 2: the s-part on tensor cores in heads_dual_kernel's
// fragment scheme (two 16-pixel MMA tiles), gate, bias, squashing, stores.
template <int C2, int MINB>
__global__ void __launch_bounds__(256, MINB) heads_gather_kernel(const uint4* __restrict__ T, int TC, const uint4* __restrict__ x2,
                                                        const __nv_bfloat16* __restrict__ w2, const __nv_bfloat16* __restrict__ gate,
                                                        const float* __restrict__ bias, float* __restrict__ seg, float* __restrict__ rbox,
                                                        float* __restrict__ angle, long long npix, int H, int W, int h, int w) {
    constexpr int N2 = C2 / 64;
    __shared__ float dsm[8][32][9];                          // [warp][pixel][8 sums + gate]
    pdl::trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    U8 wb[N2];
#pragma unroll
    for (int m = 0; m < N2; ++m) wb[m] = ldg256(w2 + (size_t)g * C2 + m * 64 + t * 16);
    const float b0 = __ldg(bias + 2 * t), b1 = __ldg(bias + 2 * t + 1);
    pdl::wait();
    const int HW = H * W;
    const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.0f, sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.0f;
    const int rec = TC / 8;                                  // uint4 per low-resolution pixel
    // CTA tile = 8 consecutive rows (one per warp) x 32 pixels: the rows share their low-resolution samples, so T is read
    // from L2 once per tile (~17 KB) and from L1 otherwise -- with one row segment per warp anywhere in the batch the kernel
    // was bound by L2 -> L1 sector traffic (1.1 GB per 32 images)
    const int xb = (W + 31) / 32, yb = (H + 7) / 8;
    const long long ntiles = (long long)(npix / HW) * yb * xb;
    const __nv_bfloat16* x2e = reinterpret_cast<const __nv_bfloat16*>(x2);
    float (*my)[9] = dsm[warp];
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long b = tile / ((long long)yb * xb);
        const int rem = (int)(tile - b * yb * xb), Y = (rem / xb) * 8 + warp, X0 = (rem % xb) * 32, X = X0 + lane;
        const bool row_ok = Y < H;
        const long long prow = (b * H + Y) * (long long)W;   // pixel index of (b, Y, 0)
        // ---- phase 1: this lane's pixel
        float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, gt = 0.0f;
        if (row_ok && X < W) {
            int ro[3][2], co[3][2];                          // low-resolution row offsets (* w) / columns of the three taps
            float rl[3], cl[3];                              // interpolation fractions; < 0: the tap lies in the padding
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int qy = Y + k - 1, qx = X + k - 1;
                const float fy = sy * (float)qy, fx = sx * (float)qx;
                const int y0 = (int)fy, x0 = (int)fx;
                ro[k][0] = y0 * w; ro[k][1] = (y0 + (y0 < h - 1 ? 1 : 0)) * w;
                co[k][0] = x0;     co[k][1] = x0 + (x0 < w - 1 ? 1 : 0);
                rl[k] = (qy >= 0 && qy < H) ? fy - (float)y0 : -1.0f;
                cl[k] = (qx >= 0 && qx < W) ? fx - (float)x0 : -1.0f;
            }
            {                                                // gate at (Y, X) = tap (1, 1)'s sample position
                const __nv_bfloat16* gp = gate + b * (long long)h * w;
                const float ly1 = rl[1], lx1 = cl[1], ly0 = 1.0f - ly1, lx0 = 1.0f - lx1;
                const float s00 = __bfloat162float(gp[ro[1][0] + co[1][0]]), s01 = __bfloat162float(gp[ro[1][0] + co[1][1]]);
                const float s10 = __bfloat162float(gp[ro[1][1] + co[1][0]]), s11 = __bfloat162float(gp[ro[1][1] + co[1][1]]);
                gt = ly0 * (lx0 * s00 + lx1 * s01) + ly1 * (lx0 * s10 + lx1 * s11);
            }
            const uint4* Tb = T + (size_t)b * h * w * rec;
            auto acc8 = [&](const uint4 v, float wgt) {
                a[0] = fmaf(wgt, __uint_as_float(v.x << 16), a[0]); a[1] = fmaf(wgt, __uint_as_float(v.x & 0xffff0000u), a[1]);
                a[2] = fmaf(wgt, __uint_as_float(v.y << 16), a[2]); a[3] = fmaf(wgt, __uint_as_float(v.y & 0xffff0000u), a[3]);
                a[4] = fmaf(wgt, __uint_as_float(v.z << 16), a[4]); a[5] = fmaf(wgt, __uint_as_float(v.z & 0xffff0000u), a[5]);
                a[6] = fmaf(wgt, __uint_as_float(v.w << 16), a[6]); a[7] = fmaf(wgt, __uint_as_float(v.w & 0xffff0000u), a[7]);
            };
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                if (rl[r] < 0.0f) continue;
                const float ly1 = rl[r], ly0 = 1.0f - ly1;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if (cl[c] < 0.0f) continue;
                    const float lx1 = cl[c], lx0 = 1.0f - lx1;
                    const uint4* q = Tb + (r * 3 + c);
                    const uint4 v00 = __ldg(q + (size_t)(ro[r][0] + co[c][0]) * rec), v01 = __ldg(q + (size_t)(ro[r][0] + co[c][1]) * rec);
                    const uint4 v10 = __ldg(q + (size_t)(ro[r][1] + co[c][0]) * rec), v11 = __ldg(q + (size_t)(ro[r][1] + co[c][1]) * rec);
                    acc8(v00, ly0 * lx0); acc8(v01, ly0 * lx1); acc8(v10, ly1 * lx0); acc8(v11, ly1 * lx1);
                }
            }
        }
        __syncwarp();                                        // the previous tile's phase 2 has read the slice
#pragma unroll
        for (int k = 0; k < 8; ++k) my[lane][k] = a[k];
        my[lane][8] = gt;
        __syncwarp();
        // ---- phase 2: two 16-pixel MMA tiles of this warp's row segment
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int l0 = half * 16 + g, l1 = l0 + 8;
            const bool ok0 = row_ok && X0 + l0 < W, ok1 = row_ok && X0 + l1 < W;
            const long long p0 = prow + X0 + l0, p1 = prow + X0 + l1;
            U8 ya[N2], yb2[N2];
#pragma unroll
            for (int m = 0; m < N2; ++m) {
                ya[m] = ok0 ? ldg256(x2e + p0 * C2 + m * 64 + t * 16) : zero256();
                yb2[m] = ok1 ? ldg256(x2e + p1 * C2 + m * 64 + t * 16) : zero256();
            }
            float acs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int m = 0; m < N2; ++m) mma_u8(acs, ya[m], yb2[m], wb[m]);
            if (ok0) heads_store(t, fmaf(my[l0][8], acs[0], my[l0][2 * t]) + b0, fmaf(my[l0][8], acs[1], my[l0][2 * t + 1]) + b1, p0, HW, seg, rbox, angle);
            if (ok1) heads_store(t, fmaf(my[l1][8], acs[2], my[l1][2 * t]) + b0, fmaf(my[l1][8], acs[3], my[l1][2 * t + 1]) + b1, p1, HW, seg, rbox, angle);
        }
    }
}

}  // namespace

extern "C" int fots_b200_heads_nhwc_bf16(const void* x, const void* wq, const float* bias, float* seg, float* rbox, float* angle,
                                         int B, int H, int W, int C, cudaStream_t stream) {
    if (!x || !wq || !bias || !seg || !rbox || !angle || B <= 0 || H <= 0 || W <= 0) return RROI_B200_ERR_INVALID_ARG;
    if (C != 128 && C != 256 && C != 512) return RROI_B200_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(wq)) & 31) return RROI_B200_ERR_INVALID_ARG;      // 32-byte loads
    const long long npix = (long long)B * H * W;
    const long long tiles = (npix + 15) / 16;
    long long ctas = (tiles + 7) / 8;
    if (ctas > 148 * 8) ctas = 148 * 8;
    const uint4* xp = static_cast<const uint4*>(x);
    const __nv_bfloat16* wp = static_cast<const __nv_bfloat16*>(wq);
    if (C == 128) (void)pdl::launch(heads_kernel<128, 0>, dim3((unsigned)ctas), dim3(256), 0, stream, xp, wp, bias, seg, rbox, angle, npix, H * W);
    else if (C == 256) (void)pdl::launch(heads_kernel<256, 0>, dim3((unsigned)ctas), dim3(256), 0, stream, xp, wp, bias, seg, rbox, angle, npix, H * W);
    else (void)pdl::launch(heads_kernel<512, 0>, dim3((unsigned)ctas), dim3(256), 0, stream, xp, wp, bias, seg, rbox, angle, npix, H * W);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
    return RROI_B200_OK;
}

// One-output-channel 1x1 convolution (+ bias) -> bf16 logits [B, 1, H, W]; wq / bias in the 8-row layout above with the
// filter in row 0 and zeros elsewhere.
extern "C" int fots_b200_conv1x1_to1_nhwc_bf16(const void* x, const void* wq, const float* bias, void* out, int B, int H, int W, int C,
                                               int sigmoid, cudaStream_t stream) {
    if (!x || !wq || !bias || !out || B <= 0 || H <= 0 || W <= 0) return RROI_B200_ERR_INVALID_ARG;
    if (C != 128 && C != 256 && C != 512) return RROI_B200_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(wq)) & 31) return RROI_B200_ERR_INVALID_ARG;      // 32-byte loads
    const long long npix = (long long)B * H * W;
    const long long tiles = (npix + 15) / 16;
    long long ctas = (tiles + 7) / 8;
    if (ctas > 148 * 8) ctas = 148 * 8;
    const uint4* xp = static_cast<const uint4*>(x);
    const __nv_bfloat16* wp = static_cast<const __nv_bfloat16*>(wq);
    float* o = static_cast<float*>(out);
    const unsigned g = (unsigned)ctas;
    if (sigmoid) {
        if (C == 128) (void)pdl::launch(heads_kernel<128, 2>, dim3(g), dim3(256), 0, stream, xp, wp, bias, o, (float*)nullptr, (float*)nullptr, npix, H * W);
        else if (C == 256) (void)pdl::launch(heads_kernel<256, 2>, dim3(g), dim3(256), 0, stream, xp, wp, bias, o, (float*)nullptr, (float*)nullptr, npix, H * W);
        else (void)pdl::launch(heads_kernel<512, 2>, dim3(g), dim3(256), 0, stream, xp, wp, bias, o, (float*)nullptr, (float*)nullptr, npix, H * W);
    } else {
        if (C == 128) (void)pdl::launch(heads_kernel<128, 1>, dim3(g), dim3(256), 0, stream, xp, wp, bias, o, (float*)nullptr, (float*)nullptr, npix, H * W);
        else if (C == 256) (void)pdl::launch(heads_kernel<256, 1>, dim3(g), dim3(256), 0, stream, xp, wp, bias, o, (float*)nullptr, (float*)nullptr, npix, H * W);
        else (void)pdl::launch(heads_kernel<512, 1>, dim3(g), dim3(256), 0, stream, xp, wp, bias, o, (float*)nullptr, (float*)nullptr, npix, H * W);
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
    return RROI_B200_OK;
}

// The three heads on  (w1 . x1) + upsample(gate_prob) * (w2 . x2) + bias  (see heads_dual_kernel): x1 bf16 [B, H, W, 256],
// x2 bf16 [B, H, W, 64], w1 bf16 [8, 256], w2 bf16 [8, 64] (8-row head layout), gate_prob bf16 [B, gh, gw], bias fp32 [8].
extern "C" int fots_b200_heads_merged_nhwc_bf16(const void* x1, const void* w1, const void* x2, const void* w2, const void* gate_prob,
                                                const float* bias, float* seg, float* rbox, float* angle, int B, int H, int W, int C1,
                                                int C2, int gh, int gw, cudaStream_t stream) {
    if (!x1 || !w1 || !x2 || !w2 || !gate_prob || !bias || !seg || !rbox || !angle || B <= 0 || H <= 0 || W <= 0 || gh <= 0 || gw <= 0)
        return RROI_B200_ERR_INVALID_ARG;
    if (C1 != 256 || C2 != 64) return RROI_B200_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(x1) | reinterpret_cast<uintptr_t>(w1) | reinterpret_cast<uintptr_t>(x2) | reinterpret_cast<uintptr_t>(w2)) & 31)
        return RROI_B200_ERR_INVALID_ARG;
    const long long npix = (long long)B * H * W;
    const long long tiles = (npix + 15) / 16;
    long long ctas = (tiles + 7) / 8;
    if (ctas > 148 * 8) ctas = 148 * 8;
    const cudaError_t e = pdl::launch(heads_dual_kernel<256, 64>, dim3((unsigned)ctas), dim3(256), 0, stream, static_cast<const uint4*>(x1),
                                      static_cast<const __nv_bfloat16*>(w1), static_cast<const uint4*>(x2), static_cast<const __nv_bfloat16*>(w2),
                                      static_cast<const __nv_bfloat16*>(gate_prob), bias, seg, rbox, angle, npix, H, W, gh, gw);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
    return RROI_B200_OK;
}

// The three heads on  sum_tap up(T_tap)(p + tap) + upsample(gate_prob) * (w2 . x2) + bias  (see heads_gather_kernel):
// T bf16 [B, h, w, TC] at the low resolution (channel tap * 8 + head column; TC >= 72, TC % 8 == 0: the 1x1 convolution that
// produces it pads its output channels), x2 bf16 [B, H, W, 64], w2 bf16 [8, 64], gate_prob bf16 [B, h, w], bias fp32 [8].
// Upsampling = bilinear, align_corners (what fots_b200_dwconv3x3_up_nhwc_bf16 computes).
extern "C" int fots_b200_heads_gather_nhwc_bf16(const void* T, int TC, const void* x2, const void* w2, const void* gate_prob,
                                                const float* bias, float* seg, float* rbox, float* angle, int B, int H, int W, int h,
                                                int w, int C2, cudaStream_t stream) {
    if (!T || !x2 || !w2 || !gate_prob || !bias || !seg || !rbox || !angle || B <= 0 || H <= 0 || W <= 0 || h <= 0 || w <= 0)
        return RROI_B200_ERR_INVALID_ARG;
    if (C2 != 64 || TC < 72 || TC % 8 != 0) return RROI_B200_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(x2) | reinterpret_cast<uintptr_t>(w2)) & 31) return RROI_B200_ERR_INVALID_ARG;       // 32-byte loads
    if (reinterpret_cast<uintptr_t>(T) & 15) return RROI_B200_ERR_INVALID_ARG;
    const long long npix = (long long)B * H * W;
    long long ctas = (long long)B * ((H + 7) / 8) * ((W + 31) / 32);      // CTA tiles of 8 rows x 32 pixels, grid-strided
    if (ctas > 148 * 16) ctas = 148 * 16;
    // 4 CTAs per SM (64 registers, a few spilled): the kernel waits on L1 / L2 loads, measured 209 us per 32 images against
    // 262 us at 2 CTAs per SM (113 registers)
    const cudaError_t e = pdl::launch(heads_gather_kernel<64, 4>, dim3((unsigned)ctas), dim3(256), 0, stream, static_cast<const uint4*>(T), TC,
                                      static_cast<const uint4*>(x2), static_cast<const __nv_bfloat16*>(w2),
                                      static_cast<const __nv_bfloat16*>(gate_prob), bias, seg, rbox, angle, npix, H, W, h, w);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
    return RROI_B200_OK;
}
