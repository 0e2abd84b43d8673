// instnorm_kernels.cu -- fused channels-last InstanceNorm + affine + residual + leaky-ReLU (bf16 in/out, fp32/fp64
// statistics) for the feeder/consumer networks.  See include/fots_b200_pipeline.h.  HBM-bound: the tensor is read
// twice and written once; every access is a 16-byte vector, consecutive threads on consecutive channel groups.
#include "../../../include/fots_b200_pipeline.h"
#include "pdl.cuh"
#include <cuda_bf16.h>
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int kThreads = 256;

struct __align__(16) Bf16x8 { __nv_bfloat162 v[4]; };

__device__ __forceinline__ Bf16x8 ld8(const Bf16x8* ptr) {
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(ptr));
    return *reinterpret_cast<const Bf16x8*>(&raw);
}
__device__ __forceinline__ void unpack8(const Bf16x8& p, float (&f)[8]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(p.v[i]);
        f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ Bf16x8 pack8(const float (&f)[8]) {
    Bf16x8 p;
#pragma unroll
    for (int i = 0; i < 4; ++i) p.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return p;
}

// grid (chunks, B).  Thread -> channel group g = tid % G (8 channels), row phase tid / G.
__global__ void __launch_bounds__(kThreads, 4) in_stats_kernel(const Bf16x8* __restrict__ x, double* __restrict__ ws,
                                                             int HW, int C, int rows_per_cta) {
    pdl::trigger();
    pdl::wait();
    const int G = C / 8;
    const int b = blockIdx.y;
    const int g = threadIdx.x % G, phase = threadIdx.x / G, nphase = kThreads / G;
    const int r0 = blockIdx.x * rows_per_cta;
    const int r1 = min(HW, r0 + rows_per_cta);
    float s[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.0f;
    if (phase < nphase) {
        const Bf16x8* base = x + ((size_t)b * HW) * G + g;
        constexpr int U = 4;                                  // rows in flight per thread (memory-level parallelism)
        for (int r = r0 + phase; r < r1; r += U * nphase) {
            Bf16x8 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (r + u * nphase < r1) v[u] = ld8(base + (size_t)(r + u * nphase) * G);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (r + u * nphase >= r1) break;
                float f[8];
                unpack8(v[u], f);
#pragma unroll
                for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] = fmaf(f[i], f[i], q[i]); }
            }
        }
    }
    // reduce over the row phases that share a channel group
    __shared__ float sh[kThreads][17];
#pragma unroll
    for (int i = 0; i < 8; ++i) { sh[threadIdx.x][i] = s[i]; sh[threadIdx.x][8 + i] = q[i]; }
    __syncthreads();
    for (int idx = threadIdx.x; idx < G * 16; idx += kThreads) {
        const int gg = idx / 16, k = idx % 16;
        double acc = 0.0;
        for (int ph = 0; ph < nphase; ++ph) acc += (double)sh[ph * G + gg][k];
        const int c = gg * 8 + (k & 7);
        atomicAdd(ws + ((size_t)b * C + c) * 2 + (k >> 3), acc);
    }
}

// grid (chunks, B).  y = act(((+-)(x - mean) * rstd) * gamma + beta [+ residual])
template <bool kCRelu, bool kResidual>
__global__ void __launch_bounds__(kThreads, kResidual ? 3 : 4) in_apply_kernel(const Bf16x8* __restrict__ x, Bf16x8* __restrict__ y,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             const Bf16x8* __restrict__ res, const double* __restrict__ ws,
                                                             int HW, int C, int rows_per_cta, float eps, float slope) {
    extern __shared__ float coef[];          // scale[Cout], shift[Cout]
    pdl::trigger();
    pdl::wait();
    const int G = C / 8;
    const int Cout = kCRelu ? 2 * C : C;
    const int b = blockIdx.y;
    float* scale = coef;
    float* shift = coef + Cout;
    for (int c = threadIdx.x; c < C; c += kThreads) {
        const double m = ws[((size_t)b * C + c) * 2] / HW;
        double var = ws[((size_t)b * C + c) * 2 + 1] / HW - m * m;
        var = var < 0.0 ? 0.0 : var;
        const float rstd = rsqrtf((float)var + eps);
        const float mean = (float)m;
        const float g0 = gamma ? gamma[c] : 1.0f, b0 = beta ? beta[c] : 0.0f;
        scale[c] = rstd * g0;
        shift[c] = b0 - mean * rstd * g0;
        if (kCRelu) {                        // channel C + c is IN(-x): mean -> -mean, same variance
            const float g1 = gamma ? gamma[C + c] : 1.0f, b1 = beta ? beta[C + c] : 0.0f;
            scale[C + c] = -rstd * g1;
            shift[C + c] = b1 + mean * rstd * g1;
        }
    }
    __syncthreads();
    // Thread -> one OUTPUT channel group g (8 channels) and a row phase.  CReLU: output groups [0, G) are IN(x), [G, 2G) are
    // IN(-x) of input group g - G (the sign lives in the coefficients), so consecutive lanes still store consecutive 16-byte
    // vectors (whole lines per warp) and the two lanes that share an input vector load the same address (one transaction).
    const int Gout = Cout / 8;
    const int g = threadIdx.x % Gout, phase = threadIdx.x / Gout, nphase = kThreads / Gout;
    if (phase >= nphase) return;
    const int gin = kCRelu ? (g >= G ? g - G : g) : g;
    const int r0 = blockIdx.x * rows_per_cta;
    const int r1 = min(HW, r0 + rows_per_cta);
    const size_t plane = (size_t)b * HW;
    // this thread's 8 channels: coefficients in registers, U rows in flight (memory-level parallelism)
    float sc[8], sh[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { sc[i] = scale[g * 8 + i]; sh[i] = shift[g * 8 + i]; }
    constexpr int U = 4;
    for (int r = r0 + phase; r < r1; r += U * nphase) {
        Bf16x8 xv[U], rv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int rr_ = r + u * nphase;
            if (rr_ < r1) {
                xv[u] = ld8(x + (plane + rr_) * G + gin);
                if (kResidual) rv[u] = ld8(res + (plane + rr_) * G + g);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int rr_ = r + u * nphase;
            if (rr_ >= r1) break;
            float f[8], o[8];
            unpack8(xv[u], f);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = fmaf(f[i], sc[i], sh[i]);
            if (kResidual) {
                float rr[8];
                unpack8(rv[u], rr);
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] += rr[i];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = o[i] > 0.0f ? o[i] : o[i] * slope;
            y[(plane + rr_) * Gout + g] = pack8(o);
        }
    }
}

// ---- backward of y = act(IN(x) * gamma + beta [+ residual]) --------------------------------------------------------------
// With xh = (x - mean) * rstd, g = dy * act'(.) (act' from the sign of the stored output y: 1 where y > 0, else slope):
//     dx   = gamma * rstd * (g - mean_hw(g) - xh * mean_hw(g * xh))          (per image and channel)
//     dres = g,   dgamma[c] = sum_n S2[n, c],   dbeta[c] = sum_n S1[n, c]     with S1 = sum_hw g, S2 = sum_hw g * xh
// Two passes like the forward: in_bwd_stats accumulates S1 / S2 per (image, channel) into an fp64 workspace, in_bwd_apply
// writes dx (and dres).  Same thread layout as in_stats_kernel / in_apply_kernel: thread = 8 channels x a row phase.
__global__ void __launch_bounds__(kThreads, 3) in_bwd_stats_kernel(const Bf16x8* __restrict__ x, const Bf16x8* __restrict__ y,
                                                                    const Bf16x8* __restrict__ dy, const double* __restrict__ ws,
                                                                    double* __restrict__ wsb, int HW, int C, int rows_per_cta, float eps,
                                                                    float slope) {
    pdl::trigger();
    pdl::wait();
    const int G = C / 8;
    const int b = blockIdx.y;
    const int g = threadIdx.x % G, phase = threadIdx.x / G, nphase = kThreads / G;
    const int r0 = blockIdx.x * rows_per_cta;
    const int r1 = min(HW, r0 + rows_per_cta);
    float s1[8], s2[8], mean[8], rstd[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        s1[i] = s2[i] = 0.0f;
        const int c = g * 8 + i;
        const double m = ws[((size_t)b * C + c) * 2] / HW;
        double var = ws[((size_t)b * C + c) * 2 + 1] / HW - m * m;
        var = var < 0.0 ? 0.0 : var;
        mean[i] = (float)m;
        rstd[i] = rsqrtf((float)var + eps);
    }
    if (phase < nphase) {
        const size_t base = ((size_t)b * HW) * G + g;
        constexpr int U = 2;
        for (int r = r0 + phase; r < r1; r += U * nphase) {
            Bf16x8 xv[U], yv[U], gv[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (r + u * nphase < r1) {
                    const size_t o = base + (size_t)(r + u * nphase) * G;
                    xv[u] = ld8(x + o); yv[u] = ld8(y + o); gv[u] = ld8(dy + o);
                }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (r + u * nphase >= r1) break;
                float fx[8], fy[8], fg[8];
                unpack8(xv[u], fx); unpack8(yv[u], fy); unpack8(gv[u], fg);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float gg = fy[i] > 0.0f ? fg[i] : fg[i] * slope;
                    s1[i] += gg;
                    s2[i] = fmaf(gg, (fx[i] - mean[i]) * rstd[i], s2[i]);
                }
            }
        }
    }
    __shared__ float sh[kThreads][17];
#pragma unroll
    for (int i = 0; i < 8; ++i) { sh[threadIdx.x][i] = s1[i]; sh[threadIdx.x][8 + i] = s2[i]; }
    __syncthreads();
    for (int idx = threadIdx.x; idx < G * 16; idx += kThreads) {
        const int gg = idx / 16, k = idx % 16;
        double acc = 0.0;
        for (int ph = 0; ph < nphase; ++ph) acc += (double)sh[ph * G + gg][k];
        const int c = gg * 8 + (k & 7);
        atomicAdd(wsb + ((size_t)b * C + c) * 2 + (k >> 3), acc);
    }
}

template <bool kResidual>
__global__ void __launch_bounds__(kThreads, 2) in_bwd_apply_kernel(const Bf16x8* __restrict__ x, const Bf16x8* __restrict__ y,
                                                                    const Bf16x8* __restrict__ dy, const float* __restrict__ gamma,
                                                                    const double* __restrict__ ws, const double* __restrict__ wsb,
                                                                    Bf16x8* __restrict__ dx, Bf16x8* __restrict__ dres, int HW, int C,
                                                                    int rows_per_cta, float eps, float slope) {
    extern __shared__ float coef[];          // a[C], m1[C], m2[C], mean[C], rstd[C]
    pdl::trigger();
    pdl::wait();
    const int G = C / 8;
    const int b = blockIdx.y;
    float *ca = coef, *cm1 = coef + C, *cm2 = coef + 2 * C, *cmean = coef + 3 * C, *crstd = coef + 4 * C;
    for (int c = threadIdx.x; c < C; c += kThreads) {
        const double m = ws[((size_t)b * C + c) * 2] / HW;
        double var = ws[((size_t)b * C + c) * 2 + 1] / HW - m * m;
        var = var < 0.0 ? 0.0 : var;
        const float rs = rsqrtf((float)var + eps);
        ca[c] = (gamma ? gamma[c] : 1.0f) * rs;
        cm1[c] = (float)(wsb[((size_t)b * C + c) * 2] / HW);
        cm2[c] = (float)(wsb[((size_t)b * C + c) * 2 + 1] / HW);
        cmean[c] = (float)m;
        crstd[c] = rs;
    }
    __syncthreads();
    const int g = threadIdx.x % G, phase = threadIdx.x / G, nphase = kThreads / G;
    if (phase >= nphase) return;
    const int r0 = blockIdx.x * rows_per_cta;
    const int r1 = min(HW, r0 + rows_per_cta);
    const size_t base = ((size_t)b * HW) * G + g;
    // this thread's 8 channels: dx = A * g - (B0 + x * B1) with A = gamma * rstd, B1 = A * rstd * m2, B0 = A * (m1 - mean * rstd * m2)
    float cA[8], cB0[8], cB1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = g * 8 + i;
        cA[i] = ca[c];
        cB1[i] = ca[c] * crstd[c] * cm2[c];
        cB0[i] = ca[c] * (cm1[c] - cmean[c] * crstd[c] * cm2[c]);
    }
    constexpr int U = 2;
    for (int r = r0 + phase; r < r1; r += U * nphase) {
        Bf16x8 xv[U], yv[U], gv[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (r + u * nphase < r1) {
                const size_t o = base + (size_t)(r + u * nphase) * G;
                xv[u] = ld8(x + o); yv[u] = ld8(y + o); gv[u] = ld8(dy + o);
            }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (r + u * nphase >= r1) break;
            const size_t o = base + (size_t)(r + u * nphase) * G;
            float fx[8], fy[8], fg[8], od[8], orr[8];
            unpack8(xv[u], fx); unpack8(yv[u], fy); unpack8(gv[u], fg);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float gg = fy[i] > 0.0f ? fg[i] : fg[i] * slope;
                od[i] = fmaf(cA[i], gg, -fmaf(fx[i], cB1[i], cB0[i]));
                orr[i] = gg;
            }
            dx[o] = pack8(od);
            if (kResidual) dres[o] = pack8(orr);
        }
    }
}

// ---- backward of the CReLU form y = act(IN(concat(x, -x)) * gamma + beta)  (y has 2C channels) -----------------------------
// IN(-x) = -xh with the same rstd, so with g1 / g2 the activation-masked gradients of the two halves and a1 = gamma[c],
// a2 = gamma[C + c]:  G = a1 * g1 - a2 * g2,  dx = rstd * (G - mean(G) - xh * mean(G * xh)).
// Workspace [B, 2C, 2]: (sum g1, sum g1 * xh) for channel c, (sum g2, -sum g2 * xh) for channel C + c, so that
// dbeta[j] = sum_b ws[b, j, 0] and dgamma[j] = sum_b ws[b, j, 1] for all 2C output channels alike.
__global__ void __launch_bounds__(kThreads, 2) in_bwd_crelu_stats_kernel(const Bf16x8* __restrict__ x, const Bf16x8* __restrict__ y,
                                                                          const Bf16x8* __restrict__ dy, const double* __restrict__ ws,
                                                                          double* __restrict__ wsb, int HW, int C, int rows_per_cta, float eps,
                                                                          float slope) {
    pdl::trigger();
    pdl::wait();
    const int G = C / 8;
    const int b = blockIdx.y;
    const int g = threadIdx.x % G, phase = threadIdx.x / G, nphase = kThreads / G;
    const int r0 = blockIdx.x * rows_per_cta;
    const int r1 = min(HW, r0 + rows_per_cta);
    float s[4][8], mean[8], rstd[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        s[0][i] = s[1][i] = s[2][i] = s[3][i] = 0.0f;
        const int c = g * 8 + i;
        const double m = ws[((size_t)b * C + c) * 2] / HW;
        double var = ws[((size_t)b * C + c) * 2 + 1] / HW - m * m;
        var = var < 0.0 ? 0.0 : var;
        mean[i] = (float)m;
        rstd[i] = rsqrtf((float)var + eps);
    }
    if (phase < nphase) {
        for (int r = r0 + phase; r < r1; r += nphase) {
            const size_t oi = ((size_t)b * HW + r) * G + g, oo = ((size_t)b * HW + r) * 2 * G + g;
            float fx[8], y1[8], y2[8], d1[8], d2[8];
            unpack8(ld8(x + oi), fx);
            unpack8(ld8(y + oo), y1); unpack8(ld8(y + oo + G), y2);
            unpack8(ld8(dy + oo), d1); unpack8(ld8(dy + oo + G), d2);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float g1 = y1[i] > 0.0f ? d1[i] : d1[i] * slope, g2 = y2[i] > 0.0f ? d2[i] : d2[i] * slope;
                const float xh = (fx[i] - mean[i]) * rstd[i];
                s[0][i] += g1; s[1][i] = fmaf(g1, xh, s[1][i]);
                s[2][i] += g2; s[3][i] = fmaf(g2, -xh, s[3][i]);
            }
        }
    }
    __shared__ float sh[kThreads][33];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) sh[threadIdx.x][k * 8 + i] = s[k][i];
    __syncthreads();
    for (int idx = threadIdx.x; idx < G * 32; idx += kThreads) {
        const int gg = idx / 32, k = idx % 32;
        double acc = 0.0;
        for (int ph = 0; ph < nphase; ++ph) acc += (double)sh[ph * G + gg][k];
        const int c = gg * 8 + (k & 7), which = k >> 3;                              // 0: S1 of c, 1: S2 of c, 2: S1 of C + c, 3: S2 of C + c
        atomicAdd(wsb + ((size_t)b * 2 * C + (which >= 2 ? C : 0) + c) * 2 + (which & 1), acc);
    }
}

__global__ void __launch_bounds__(kThreads, 2) in_bwd_crelu_apply_kernel(const Bf16x8* __restrict__ x, const Bf16x8* __restrict__ y,
                                                                          const Bf16x8* __restrict__ dy, const float* __restrict__ gamma,
                                                                          const double* __restrict__ ws, const double* __restrict__ wsb,
                                                                          Bf16x8* __restrict__ dx, int HW, int C, int rows_per_cta, float eps,
                                                                          float slope) {
    pdl::trigger();
    pdl::wait();
    const int G = C / 8;
    const int b = blockIdx.y;
    const int g = threadIdx.x % G, phase = threadIdx.x / G, nphase = kThreads / G;
    if (phase >= nphase) return;
    // dx = rstd * (a1 g1 - a2 g2 - mG - xh * mGx) = A1 g1 - A2 g2 - (B0 + x * B1)
    float A1[8], A2[8], B0[8], B1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = g * 8 + i;
        const double m = ws[((size_t)b * C + c) * 2] / HW;
        double var = ws[((size_t)b * C + c) * 2 + 1] / HW - m * m;
        var = var < 0.0 ? 0.0 : var;
        const float rs = rsqrtf((float)var + eps), mean = (float)m;
        const float a1 = gamma ? gamma[c] : 1.0f, a2 = gamma ? gamma[C + c] : 1.0f;
        const double* w1 = wsb + ((size_t)b * 2 * C + c) * 2;
        const double* w2 = wsb + ((size_t)b * 2 * C + C + c) * 2;
        const float mG = (float)((a1 * w1[0] - a2 * w2[0]) / HW);
        const float mGx = (float)((a1 * w1[1] + a2 * w2[1]) / HW);                  // w2[1] already carries the minus sign
        A1[i] = rs * a1; A2[i] = rs * a2;
        B1[i] = rs * rs * mGx;
        B0[i] = rs * (mG - mean * rs * mGx);
    }
    const int r0 = blockIdx.x * rows_per_cta;
    const int r1 = min(HW, r0 + rows_per_cta);
    for (int r = r0 + phase; r < r1; r += nphase) {
        const size_t oi = ((size_t)b * HW + r) * G + g, oo = ((size_t)b * HW + r) * 2 * G + g;
        float fx[8], y1[8], y2[8], d1[8], d2[8], od[8];
        unpack8(ld8(x + oi), fx);
        unpack8(ld8(y + oo), y1); unpack8(ld8(y + oo + G), y2);
        unpack8(ld8(dy + oo), d1); unpack8(ld8(dy + oo + G), d2);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float g1 = y1[i] > 0.0f ? d1[i] : d1[i] * slope, g2 = y2[i] > 0.0f ? d2[i] : d2[i] * slope;
            od[i] = fmaf(A1[i], g1, fmaf(-A2[i], g2, -fmaf(fx[i], B1[i], B0[i])));
        }
        dx[oi] = pack8(od);
    }
}

// ---- single-pass InstanceNorm for instances that fit in a cluster's shared memory --------------------------------
// The two-pass form above costs three launches (memset, statistics, apply) and reads x twice.  Most InstanceNorms of the
// step act on small instances (stages 2-4 of the feeder: <= 3.7 MB per image; the recogniser: 131 KB per RoI), where the
// launches' ramp and tail cost as much as the data.  Here ONE launch does it: a cluster of CS CTAs owns one (image, slice
// of the channels); every CTA loads its share of the pixels ONCE into shared memory while accumulating per-channel
// sums, the per-CTA partial sums are exchanged through distributed shared memory (each CTA adds the CS partials in rank
// order: deterministic, no atomics, no workspace), and the normalised / residual-added / activated tensor is written
// from the shared-memory copy.  x is read from HBM exactly once.
__device__ __forceinline__ uint32_t in_cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void in_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem_f32(const float* local, uint32_t rank) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(local), ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra));
    return v;
}

// grid (CS * slices, B), cluster (CS, 1, 1).  Cs = channels per slice (multiple of 8, Cs / 8 divides 256).
template <bool kResidual>
__global__ void __launch_bounds__(kThreads) in_fused_cluster_kernel(const Bf16x8* __restrict__ x, Bf16x8* __restrict__ y,
                                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                     const Bf16x8* __restrict__ res, int HW, int C, int Cs, int CS,
                                                                     int rows_per_cta, float eps, float slope) {
    extern __shared__ __align__(16) uint8_t in_smem[];
    pdl::trigger();
    pdl::wait();
    const int G = Cs / 8, Gall = C / 8;
    Bf16x8* const slab = reinterpret_cast<Bf16x8*>(in_smem);                                   // [rows_per_cta * G]
    float* const red = reinterpret_cast<float*>(in_smem + (size_t)rows_per_cta * G * 16);      // [256][17] block reduction
    float* const part = red + kThreads * 17;                                                   // [2 * Cs] this CTA's partial sums
    float* const coef = part + 2 * Cs;                                                         // [2 * Cs] scale, shift
    const uint32_t rank = in_cluster_rank();
    const int slice = blockIdx.x / CS, b = blockIdx.y;
    const int r0 = (int)rank * rows_per_cta, r1 = min(HW, r0 + rows_per_cta);
    const int g = threadIdx.x % G;                                  // this thread's channel group (G divides 256)
    const size_t base = ((size_t)b * HW + r0) * Gall + (size_t)slice * G;

    // ---- phase 1: load the slab once, keep it in shared memory, accumulate the sums
    float s[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.0f;
    constexpr int U = 4;
    const int nrows = max(r1 - r0, 0), nph = kThreads / G, phase = threadIdx.x / G;   // thread -> rows phase, phase + nph, ...
    // every 16-byte vector of the slab goes straight to shared memory with cp.async: no registers are held, so the WHOLE
    // slab (up to 160 KB) is in flight at once -- with one resident CTA per SM that is what hides the HBM latency
    // (register-staged loads, 4 per thread, kept only 16 KB in flight per SM and ran at a third of the bandwidth)
    {
        const uint32_t slab_s = (uint32_t)__cvta_generic_to_shared(slab);
        for (int row = phase; row < nrows; row += nph)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slab_s + (uint32_t)((row * G + g) * 16)),
                         "l"(x + base + (size_t)row * Gall + g) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    for (int row = phase; row < nrows; row += nph) {            // this thread sums exactly the vectors it copied
        float f[8];
        unpack8(slab[row * G + g], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) { s[k] += f[k]; q[k] = fmaf(f[k], f[k], q[k]); }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) { red[threadIdx.x * 17 + k] = s[k]; red[threadIdx.x * 17 + 8 + k] = q[k]; }
    __syncthreads();
    for (int idx = threadIdx.x; idx < G * 16; idx += kThreads) {
        const int gg = idx / 16, k = idx % 16;
        float acc = 0.0f;
        for (int ph = 0; ph < nph; ++ph) acc += red[(ph * G + gg) * 17 + k];
        part[(k >> 3) * Cs + gg * 8 + (k & 7)] = acc;               // [sum | sumsq][channel of the slice]
    }
    in_cluster_sync();                                              // every CTA's partials are visible cluster-wide

    // ---- statistics of the whole instance: add the CS partials in rank order (every CTA gets the same bits)
    for (int c = threadIdx.x; c < Cs; c += kThreads) {
        double sum = 0.0, sq = 0.0;
        for (int rk = 0; rk < CS; ++rk) {
            sum += (double)ld_dsmem_f32(part + c, (uint32_t)rk);
            sq += (double)ld_dsmem_f32(part + Cs + c, (uint32_t)rk);
        }
        const double m = sum / HW;
        double var = sq / HW - m * m;
        var = var < 0.0 ? 0.0 : var;
        const float rstd = rsqrtf((float)var + eps), mean = (float)m;
        const int cg = slice * Cs + c;
        const float g0 = gamma ? gamma[cg] : 1.0f, b0 = beta ? beta[cg] : 0.0f;
        coef[c] = rstd * g0;
        coef[Cs + c] = b0 - mean * rstd * g0;
    }
    in_cluster_sync();                                              // peers are done reading this CTA's partials; coef visible to the CTA

    // ---- phase 2: normalise from shared memory
    float sc[8], sh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { sc[k] = coef[g * 8 + k]; sh[k] = coef[Cs + g * 8 + k]; }
    for (int row = phase; row < nrows; row += U * nph) {
        Bf16x8 rv[U];
        if (kResidual) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int rr = row + u * nph;
                if (rr < nrows) rv[u] = ld8(res + base + (size_t)rr * Gall + g);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int rr = row + u * nph;
            if (rr >= nrows) break;
            float f[8], o[8];
            unpack8(slab[rr * G + g], f);
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = fmaf(f[k], sc[k], sh[k]);
            if (kResidual) {
                float rr8[8];
                unpack8(rv[u], rr8);
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] += rr8[k];
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = o[k] > 0.0f ? o[k] : o[k] * slope;
            y[base + (size_t)rr * Gall + g] = pack8(o);
        }
    }
}

// Shape of the single-pass launch, or false when the instance does not fit: cluster size CS (<= 8, portable), channel
// slices, rows per CTA, dynamic shared memory.
struct FusedPlan { int CS, slices, Cs, rows; size_t smem; };
// MEASURED on B200 (tools/in_bench.py, profiles/r02_instnorm_single_pass.txt): with slabs of 115-160 KB only one CTA is
// resident per SM and the load -> sums -> exchange -> apply phases of a cluster do not overlap with anything: 42 vs 23 us
// (stage 2), 22 vs 17 (stage 3), 13 vs 12 (stage 4), 45 vs 42 (recogniser, 131 KB per RoI) against the two-pass kernels,
// which run at full occupancy.  It wins where several instances share an SM: 12.2 vs 16.0 us on batch10_s (32 KB per
// RoI).  Hence the automatic limit of 40 KB per CTA; `force` (tests, sweeps) allows the full 160 KB.
static bool plan_fused(int HW, int C, FusedPlan* out, bool force) {
    const size_t kSlabMax = (force ? 160 : 40) * 1024;
    const int max_split = force ? 8 : 1;              // automatic: whole instance in ONE CTA (no channel slices, no cluster)
    for (int slices = 1; slices <= max_split; slices <<= 1) {
        if (C % slices) break;
        const int Cs = C / slices;
        if (Cs < 32 || Cs % 8 != 0 || (kThreads % (Cs / 8)) != 0) break;        // >= 64-byte pixel segments; fixed channel group per thread
        for (int CS = 1; CS <= max_split; CS <<= 1) {
            const int rows = (HW + CS - 1) / CS;
            const size_t slab = (size_t)rows * Cs * 2;
            if (slab <= kSlabMax) {
                out->CS = CS; out->slices = slices; out->Cs = Cs; out->rows = rows;
                out->smem = slab + (size_t)kThreads * 17 * 4 + (size_t)4 * Cs * 4;
                return true;
            }
        }
    }
    return false;
}

template <bool kResidual>
static cudaError_t launch_fused(const FusedPlan& pl, const void* x, void* y, const float* gamma, const float* beta, const void* residual,
                                int B, int HW, int C, float eps, float slope, cudaStream_t stream) {
    static bool done[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !done[dev]) {
        e = cudaFuncSetAttribute(in_fused_cluster_kernel<kResidual>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) done[dev] = true;
    }
    return pdl::launch_cluster(in_fused_cluster_kernel<kResidual>, dim3((unsigned)(pl.CS * pl.slices), (unsigned)B), dim3(kThreads), pl.smem,
                               stream, (unsigned)pl.CS, static_cast<const Bf16x8*>(x), static_cast<Bf16x8*>(y), gamma, beta,
                               static_cast<const Bf16x8*>(residual), HW, C, pl.Cs, pl.CS, pl.rows, eps, slope);
}

// ---- fused top-down merge: y = (up(a_lo) | c_hi) + b_hi * (up(sigmoid(g_lo)) | 1) --------------------------------
// thread = one output pixel x 8 channels; consecutive threads on consecutive channel groups (16-byte accesses).
struct Lerp { int i0, i1; float l0, l1; };
__device__ __forceinline__ Lerp lerp_coord(int dst, int in, float scale) {   // torch: area_pixel_compute_source_index, align_corners
    Lerp r;
    const float s = scale * (float)dst;
    r.i0 = (int)s;
    r.i1 = r.i0 + ((r.i0 < in - 1) ? 1 : 0);
    r.l1 = s - (float)r.i0;
    r.l0 = 1.0f - r.l1;
    return r;
}

// Work item = kMergeU * kThreads consecutive 16-byte vectors of ONE output row (b, Y): the row's vertical coordinates are
// CTA-uniform, everything per thread is 32-bit, and a thread has kMergeU independent vectors (all their loads) in flight.
// These kernels are ISSUE-bound, not HBM-bound (ncu: 390 warp-level instructions per output vector in the first version,
// 70 % issue utilisation at 0.36 of the copy roofline), hence: row-relative 32-bit offsets, a shift instead of a division
// when C / 8 is a power of two, and gate_is_prob -- the gate map may already hold sigmoid(logit) (what the one-channel
// convolution kernel can emit), so the four expf + divisions per vector, repeated by all C / 8 lanes of a pixel, disappear.
// kUp: the first operand is the low-resolution a_lo (four taps per vector, kMergeU = 2); else the full-resolution c_hi (4).
template <bool kUp, int kMergeU>
__global__ void __launch_bounds__(kThreads, 3) fpn_merge_kernel(const Bf16x8* __restrict__ a_lo, const Bf16x8* __restrict__ c_hi,
                                                                 const Bf16x8* __restrict__ b_hi, const __nv_bfloat16* __restrict__ g_lo,
                                                                 Bf16x8* __restrict__ y, int B, int h, int w, int H, int W, int C,
                                                                 int segs, long long items, int gshift, int gate_is_prob) {
    pdl::trigger();
    pdl::wait();
    const int G = C / 8;
    const int rowv = W * G;                                   // vectors per output row
    const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.0f;
    const float sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.0f;
    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
        const int seg = (int)(item % segs);
        const long long row = item / segs;                    // b * H + Y
        const int Y = (int)(row % H);
        const int b = (int)(row / H);
        const Lerp ly = lerp_coord(Y, h, sy);
        // row bases: everything below is a 32-bit offset from one of these
        const Bf16x8* a0 = kUp ? a_lo + ((size_t)b * h + ly.i0) * w * G : nullptr;
        const Bf16x8* a1 = kUp ? a_lo + ((size_t)b * h + ly.i1) * w * G : nullptr;
        const __nv_bfloat16* g0 = g_lo ? g_lo + ((size_t)b * h + ly.i0) * w : nullptr;
        const __nv_bfloat16* g1 = g_lo ? g_lo + ((size_t)b * h + ly.i1) * w : nullptr;
        const Bf16x8* crow = kUp ? nullptr : c_hi + (size_t)row * rowv;
        const Bf16x8* brow = b_hi ? b_hi + (size_t)row * rowv : nullptr;
        Bf16x8* yrow = y + (size_t)row * rowv;
        const int i0 = seg * (kMergeU * kThreads) + threadIdx.x;
        Bf16x8 va[kMergeU][kUp ? 4 : 1], vb[kMergeU];
        float gl[kMergeU][4];
#pragma unroll
        for (int u = 0; u < kMergeU; ++u) {
            const int i = i0 + u * kThreads;
            if (i >= rowv) break;
            const int X = gshift >= 0 ? (i >> gshift) : i / G;
            const int g = i - X * G;
            const Lerp lxu = lerp_coord(X, w, sx);
            if (kUp) {
                va[u][0] = ld8(a0 + lxu.i0 * G + g); va[u][1] = ld8(a0 + lxu.i1 * G + g);
                va[u][2] = ld8(a1 + lxu.i0 * G + g); va[u][3] = ld8(a1 + lxu.i1 * G + g);
            } else {
                va[u][0] = ld8(crow + i);
            }
            if (b_hi) {
                vb[u] = ld8(brow + i);
                if (g_lo) {
                    gl[u][0] = __bfloat162float(g0[lxu.i0]); gl[u][1] = __bfloat162float(g0[lxu.i1]);
                    gl[u][2] = __bfloat162float(g1[lxu.i0]); gl[u][3] = __bfloat162float(g1[lxu.i1]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kMergeU; ++u) {
            const int i = i0 + u * kThreads;
            if (i >= rowv) break;
            float o[8];
            const Lerp lxu = lerp_coord(gshift >= 0 ? (i >> gshift) : i / G, w, sx);   // recomputed: cheaper than registers held across the loads
            if (kUp) {
                float v00[8], v01[8], v10[8], v11[8];
                unpack8(va[u][0], v00); unpack8(va[u][1], v01); unpack8(va[u][2], v10); unpack8(va[u][3], v11);
#pragma unroll
                for (int k = 0; k < 8; ++k) {                 // explicit contraction: the depthwise kernel's upsample-on-load repeats it bit for bit
                    const float t0 = fmaf(lxu.l1, v01[k], __fmul_rn(lxu.l0, v00[k])), t1 = fmaf(lxu.l1, v11[k], __fmul_rn(lxu.l0, v10[k]));
                    o[k] = fmaf(ly.l1, t1, __fmul_rn(ly.l0, t0));
                }
            } else {
                unpack8(va[u][0], o);
            }
            if (b_hi) {
                float gate = 1.0f;
                if (g_lo) {
                    float s00 = gl[u][0], s01 = gl[u][1], s10 = gl[u][2], s11 = gl[u][3];
                    if (!gate_is_prob) {
                        s00 = 1.0f / (1.0f + __expf(-s00)); s01 = 1.0f / (1.0f + __expf(-s01));
                        s10 = 1.0f / (1.0f + __expf(-s10)); s11 = 1.0f / (1.0f + __expf(-s11));
                    }
                    gate = ly.l0 * (lxu.l0 * s00 + lxu.l1 * s01) + ly.l1 * (lxu.l0 * s10 + lxu.l1 * s11);
                }
                float bb[8];
                unpack8(vb[u], bb);
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = fmaf(bb[k], gate, o[k]);
            }
            yrow[i] = pack8(o);
        }
    }
}

// ---- backward of the bilinear (align_corners) upsampling of the top-down merge: dlo = U^T dhi --------------------------------
// Gather form: thread = one LOW-resolution pixel x 8 channels; it visits the high-resolution rows / columns whose source
// index pair (lerp_coord, the forward's own arithmetic) contains its row / column and accumulates weight * gradient in fp32.
// No atomics, deterministic; every high-resolution vector is read by the <= 4 low-resolution pixels it was interpolated from.
__global__ void __launch_bounds__(kThreads) upsample_bwd_kernel(const Bf16x8* __restrict__ dhi, Bf16x8* __restrict__ dlo, int B, int h, int w,
                                                                int H, int W, int C) {
    pdl::trigger();
    pdl::wait();
    const int G = C / 8;
    const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.0f;
    const float sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.0f;
    const long long total = (long long)B * h * w * G;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
        const int g = (int)(i % G);
        long long p = i / G;
        const int x = (int)(p % w); p /= w;
        const int y = (int)(p % h);
        const int b = (int)(p / h);
        // candidate ranges: source index s * Y lies in (y - 1, y + 1); two extra on each side absorb the float rounding
        int Ya = 0, Yb = H - 1, Xa = 0, Xb = W - 1;
        if (sy > 0.0f) { Ya = max(0, (int)floorf((float)(y - 1) / sy) - 1); Yb = min(H - 1, (int)ceilf((float)(y + 1) / sy) + 1); }
        if (sx > 0.0f) { Xa = max(0, (int)floorf((float)(x - 1) / sx) - 1); Xb = min(W - 1, (int)ceilf((float)(x + 1) / sx) + 1); }
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.0f;
        for (int Y = Ya; Y <= Yb; ++Y) {
            const Lerp ly = lerp_coord(Y, h, sy);
            const float wy = (ly.i0 == y ? ly.l0 : 0.0f) + (ly.i1 == y ? ly.l1 : 0.0f);
            if (wy == 0.0f) continue;
            const Bf16x8* rowp = dhi + (((size_t)b * H + Y) * W) * G + g;
            for (int X = Xa; X <= Xb; ++X) {
                const Lerp lx = lerp_coord(X, w, sx);
                const float wx = (lx.i0 == x ? lx.l0 : 0.0f) + (lx.i1 == x ? lx.l1 : 0.0f);
                if (wx == 0.0f) continue;
                float f[8];
                unpack8(ld8(rowp + (size_t)X * G), f);
                const float wgt = wy * wx;
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] = fmaf(wgt, f[k], acc[k]);
            }
        }
        dlo[i] = pack8(acc);
    }
}

// ---- max-pool (2,1)/(2,1) over H of a channels-last tensor (tools/models.py:344, :360 `max2`) --------------------
// thread = 8 channels of one output pixel; the two input rows are W*C elements apart.  NaNs propagate like torch's.
__global__ void __launch_bounds__(kThreads) maxpool_h2_kernel(const uint4* __restrict__ x, uint4* __restrict__ y,
                                                               long long total, int Ho, long long rowv, int H) {
    pdl::trigger();
    pdl::wait();
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
        const long long r = i / rowv, e = i - r * rowv;        // r = n * Ho + ho
        const long long n = r / Ho, ho = r - n * Ho;
        const uint4* src = x + ((n * H + 2 * ho) * rowv + e);
        const uint4 a = __ldg(src), b = __ldg(src + rowv);
        uint4 o;
        const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
        const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
        __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int k = 0; k < 4; ++k) po[k] = __hmax2_nan(pa[k], pb[k]);
        y[i] = o;
    }
}


// CTAs of `kernel` the device holds at once (kThreads threads, `smem` dynamic bytes).  The grids below are sized in whole
// waves of this number: a grid of 1 192 CTAs on 592 slots runs as THREE waves, the last one 1 % full.
template <typename K>
static int resident_ctas(K kernel, size_t smem) {
    int dev = 0, sms = 0, occ = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kThreads, smem) != cudaSuccess || sms <= 0 || occ <= 0) {
        (void)cudaGetLastError();
        return 148 * 4;
    }
    return sms * occ;
}

// Chunks per image so that B * chunks CTAs fill whole waves of `slots`: the best wave efficiency among the splits that keep
// at least `min_rows` rows per CTA and at most two waves (more CTAs only repeat the per-CTA prologue); ties -> finer split.
static int wave_chunks(int B, int HW, int slots, int min_rows, int nphase, int* rows_out) {
    int max_chunks = HW / (min_rows > 0 ? min_rows : 1);
    if (max_chunks < 1) max_chunks = 1;
    int best = 1;
    double best_eff = -1.0;
    for (int c = 1; c <= max_chunks; ++c) {
        const long long total = (long long)B * c;
        if (c > 1 && total > 2LL * slots) break;
        const long long waves = (total + slots - 1) / slots;
        const double eff = (double)total / (double)(waves * slots);
        if (eff >= best_eff) { best_eff = eff; best = c; }
    }
    int rows = (HW + best - 1) / best;
    rows = ((rows + nphase - 1) / nphase) * nphase;
    *rows_out = rows;
    return (HW + rows - 1) / rows;
}

}  // namespace

extern "C" int fots_b200_maxpool_h2_nhwc_bf16(const void* x, void* y, int N, int H, int W, int C, cudaStream_t stream) {
    if (!x || !y || N <= 0 || H < 2 || W <= 0 || C <= 0 || C % 8 != 0) return RROI_B200_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) return RROI_B200_ERR_INVALID_ARG;
    const int Ho = H / 2;
    const long long rowv = (long long)W * (C / 8);
    const long long total = (long long)N * Ho * rowv;
    long long grid = (total + kThreads - 1) / kThreads;
    if (grid > 148LL * 16) grid = 148LL * 16;
    (void)pdl::launch(maxpool_h2_kernel, dim3((unsigned)grid), dim3(kThreads), 0, stream, static_cast<const uint4*>(x), static_cast<uint4*>(y), total, Ho, rowv, H);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
    return RROI_B200_OK;
}

static int fpn_merge_impl(const void* a_lo, const void* c_hi, const void* b_hi, const void* g_lo, void* y, int B, int h, int w, int H,
                          int W, int C, int gate_is_prob, cudaStream_t stream) {
    if (!y || ((a_lo == nullptr) == (c_hi == nullptr)) || (g_lo && !b_hi) || B <= 0 || h <= 0 || w <= 0 || H <= 0 ||
        W <= 0 || C <= 0 || C % 8 != 0)
        return RROI_B200_ERR_INVALID_ARG;
    if ((long long)W * (C / 8) > (1LL << 30) || (long long)w * (C / 8) > (1LL << 30)) return RROI_B200_ERR_INVALID_ARG;
    const int G = C / 8;
    const int rowv = W * G;
    int gshift = -1;
    if ((G & (G - 1)) == 0) { gshift = 0; while ((1 << gshift) < G) ++gshift; }
    const int U = a_lo ? 2 : 4;
    const int segs = (rowv + U * kThreads - 1) / (U * kThreads);
    const long long items = (long long)B * H * segs;
    const long long slots = a_lo ? resident_ctas(fpn_merge_kernel<true, 2>, 0) : resident_ctas(fpn_merge_kernel<false, 4>, 0);
    long long grid = items;
    if (grid > slots) {                                       // equal shares of the items, at most two waves of CTAs
        const long long per = (items + 2 * slots - 1) / (2 * slots);
        grid = (items + per - 1) / per;
    }
    const Bf16x8 *ap = static_cast<const Bf16x8*>(a_lo), *cp = static_cast<const Bf16x8*>(c_hi), *bp = static_cast<const Bf16x8*>(b_hi);
    const __nv_bfloat16* gp = static_cast<const __nv_bfloat16*>(g_lo);
    Bf16x8* yp = static_cast<Bf16x8*>(y);
    if (a_lo) (void)pdl::launch(fpn_merge_kernel<true, 2>, dim3((unsigned)grid), dim3(kThreads), 0, stream, ap, cp, bp, gp, yp, B, h, w, H, W, C, segs, items, gshift, gate_is_prob);
    else (void)pdl::launch(fpn_merge_kernel<false, 4>, dim3((unsigned)grid), dim3(kThreads), 0, stream, ap, cp, bp, gp, yp, B, h, w, H, W, C, segs, items, gshift, gate_is_prob);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
    return RROI_B200_OK;
}

extern "C" int fots_b200_fpn_merge_nhwc_bf16(const void* a_lo, const void* c_hi, const void* b_hi, const void* g_lo,
                                             void* y, int B, int h, int w, int H, int W, int C, cudaStream_t stream) {
    return fpn_merge_impl(a_lo, c_hi, b_hi, g_lo, y, B, h, w, H, W, C, 0, stream);
}

// The same with the gate map already holding sigmoid(logit) (fots_b200_conv1x1_to1_nhwc_bf16 with sigmoid = 1).
extern "C" int fots_b200_fpn_merge_prob_nhwc_bf16(const void* a_lo, const void* c_hi, const void* b_hi, const void* g_prob_lo,
                                                  void* y, int B, int h, int w, int H, int W, int C, cudaStream_t stream) {
    return fpn_merge_impl(a_lo, c_hi, b_hi, g_prob_lo, y, B, h, w, H, W, C, 1, stream);
}

// A/B switch for sweeps and tests (fots_b200_instnorm_set_single_pass): the single-pass kernel is the default
static std::atomic<int> g_in_single_pass_sw{1};   // 0 = never, 1 = automatic (small instances), 2 = whenever the instance fits
extern "C" int fots_b200_instnorm_set_single_pass(int mode) {
    if (mode < 0 || mode > 2) return RROI_B200_ERR_INVALID_ARG;
    g_in_single_pass_sw.store(mode, std::memory_order_relaxed);
    return RROI_B200_OK;
}

static int instnorm_impl(const void* x, void* y, const float* gamma, const float* beta, const void* residual,
                         double* workspace, int B, int HW, int C, float eps, float slope, int crelu, bool have_stats,
                         cudaStream_t stream) {
    if (!x || !y || !workspace || B <= 0 || HW <= 0 || C <= 0 || C % 8 != 0 || C > 1024 || (C / 8) > kThreads ||
        ((gamma == nullptr) != (beta == nullptr)) || (crelu && residual))
        return RROI_B200_ERR_INVALID_ARG;
    // small instances: one launch, x read once (single-pass cluster kernel); B >= 65536 exceeds gridDim.y
    FusedPlan pl;
    const int g_in_single_pass = g_in_single_pass_sw.load(std::memory_order_relaxed);
    if (!have_stats && !crelu && B < 65536 && g_in_single_pass && plan_fused(HW, C, &pl, g_in_single_pass == 2)) {
        const cudaError_t ef = residual ? launch_fused<true>(pl, x, y, gamma, beta, residual, B, HW, C, eps, slope, stream)
                                        : launch_fused<false>(pl, x, y, gamma, beta, nullptr, B, HW, C, eps, slope, stream);
        if (ef != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
        return RROI_B200_OK;
    }
    const int G = C / 8;
    const int nphase = kThreads / G;
    // whole waves of resident CTAs, at least a handful of rows per thread; the apply kernel is the longer of the two launches
    const size_t smem = (size_t)(crelu ? 4 : 2) * C * sizeof(float);
    const int slots = crelu ? resident_ctas(in_apply_kernel<true, false>, smem)
                            : residual ? resident_ctas(in_apply_kernel<false, true>, smem) : resident_ctas(in_apply_kernel<false, false>, smem);
    int rows = 0;
    const int chunks = wave_chunks(B, HW, slots, nphase * 4, nphase, &rows);
    cudaError_t e = cudaSuccess;
    const dim3 grid(chunks, B);
    if (!have_stats) {
        e = pdl::zero_f64(workspace, (size_t)B * C * 2, stream);
        if (e != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
        int srows = 0;
        const int schunks = wave_chunks(B, HW, resident_ctas(in_stats_kernel, 0), nphase * 4, nphase, &srows);
        (void)pdl::launch(in_stats_kernel, dim3(schunks, B), dim3(kThreads), 0, stream, static_cast<const Bf16x8*>(x), workspace, HW, C, srows);
    }
    const Bf16x8* xr = static_cast<const Bf16x8*>(x);
    Bf16x8* yr = static_cast<Bf16x8*>(y);
    const Bf16x8* rr = static_cast<const Bf16x8*>(residual);
    if (crelu)
        (void)pdl::launch(in_apply_kernel<true, false>, dim3(grid), dim3(kThreads), smem, stream, xr, yr, gamma, beta, (const Bf16x8*)nullptr, workspace, HW, C, rows, eps, slope);
    else if (residual)
        (void)pdl::launch(in_apply_kernel<false, true>, dim3(grid), dim3(kThreads), smem, stream, xr, yr, gamma, beta, rr, workspace, HW, C, rows, eps, slope);
    else
        (void)pdl::launch(in_apply_kernel<false, false>, dim3(grid), dim3(kThreads), smem, stream, xr, yr, gamma, beta, (const Bf16x8*)nullptr, workspace, HW, C, rows, eps, slope);
    e = cudaGetLastError();
    if (e != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
    return RROI_B200_OK;
}

// statistics pass only: ws [B, C, 2] (cleared here) <- per-image per-channel sum / sum of squares of x
extern "C" int fots_b200_instnorm_stats_nhwc_bf16(const void* x, double* workspace, int B, int HW, int C, cudaStream_t stream) {
    if (!x || !workspace || B <= 0 || HW <= 0 || C <= 0 || C % 8 != 0 || C > 1024 || (C / 8) > kThreads || B > 65535)
        return RROI_B200_ERR_INVALID_ARG;
    const int G = C / 8, nphase = kThreads / G;
    int rows = 0;
    const int chunks = wave_chunks(B, HW, resident_ctas(in_stats_kernel, 0), nphase * 4, nphase, &rows);
    cudaError_t e = pdl::zero_f64(workspace, (size_t)B * C * 2, stream);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
    (void)pdl::launch(in_stats_kernel, dim3(chunks, B), dim3(kThreads), 0, stream, static_cast<const Bf16x8*>(x), workspace, HW, C, rows);
    e = cudaGetLastError();
    if (e != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
    return RROI_B200_OK;
}

extern "C" int fots_b200_instnorm_nhwc_bf16(const void* x, void* y, const float* gamma, const float* beta,
                                            const void* residual, double* workspace, int B, int HW, int C,
                                            float eps, float slope, int crelu, cudaStream_t stream) {
    return instnorm_impl(x, y, gamma, beta, residual, workspace, B, HW, C, eps, slope, crelu, false, stream);
}

extern "C" int fots_b200_instnorm_apply_nhwc_bf16(const void* x, void* y, const float* gamma, const float* beta,
                                                  const void* residual, const double* stats, int B, int HW, int C,
                                                  float eps, float slope, int crelu, cudaStream_t stream) {
    return instnorm_impl(x, y, gamma, beta, residual, const_cast<double*>(stats), B, HW, C, eps, slope, crelu, true, stream);
}

// Backward of fots_b200_instnorm_nhwc_bf16 (crelu == 0): see the kernels.  stats = the forward's fp64 [B, C, 2] sums of x;
// stats_bwd fp64 [B, C, 2] is cleared and filled here (S1 = sum g, S2 = sum g * xhat per image and channel: the caller reduces
// them over the images for dgamma = sum_n S2, dbeta = sum_n S1).  dres may be NULL.
extern "C" int fots_b200_instnorm_bwd_nhwc_bf16(const void* x, const void* y, const void* dy, const float* gamma, const double* stats,
                                                double* stats_bwd, void* dx, void* dres, int B, int HW, int C, float eps, float slope,
                                                cudaStream_t stream) {
    if (!x || !y || !dy || !stats || !stats_bwd || !dx || B <= 0 || HW <= 0 || C <= 0 || C % 8 != 0 || C > 1024 || (C / 8) > kThreads ||
        B > 65535)
        return RROI_B200_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx) |
         reinterpret_cast<uintptr_t>(dres)) & 15)
        return RROI_B200_ERR_INVALID_ARG;
    const int G = C / 8, nphase = kThreads / G;
    cudaError_t e = pdl::zero_f64(stats_bwd, (size_t)B * C * 2, stream);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
    const Bf16x8 *xp = static_cast<const Bf16x8*>(x), *yp = static_cast<const Bf16x8*>(y), *gp = static_cast<const Bf16x8*>(dy);
    int rows = 0;
    const int chunks = wave_chunks(B, HW, resident_ctas(in_bwd_stats_kernel, 0), nphase * 4, nphase, &rows);
    e = pdl::launch(in_bwd_stats_kernel, dim3(chunks, B), dim3(kThreads), 0, stream, xp, yp, gp, stats, stats_bwd, HW, C, rows, eps, slope);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
    const size_t smem = (size_t)5 * C * sizeof(float);
    int arows = 0;
    if (dres) {
        const int ach = wave_chunks(B, HW, resident_ctas(in_bwd_apply_kernel<true>, smem), nphase * 4, nphase, &arows);
        e = pdl::launch(in_bwd_apply_kernel<true>, dim3(ach, B), dim3(kThreads), smem, stream, xp, yp, gp, gamma, stats, (const double*)stats_bwd,
                        static_cast<Bf16x8*>(dx), static_cast<Bf16x8*>(dres), HW, C, arows, eps, slope);
    } else {
        const int ach = wave_chunks(B, HW, resident_ctas(in_bwd_apply_kernel<false>, smem), nphase * 4, nphase, &arows);
        e = pdl::launch(in_bwd_apply_kernel<false>, dim3(ach, B), dim3(kThreads), smem, stream, xp, yp, gp, gamma, stats, (const double*)stats_bwd,
                        static_cast<Bf16x8*>(dx), (Bf16x8*)nullptr, HW, C, arows, eps, slope);
    }
    if (e != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
    return RROI_B200_OK;
}

// Backward of the CReLU form (crelu != 0): y / dy bf16 [B, HW, 2C], gamma fp32 [2C] or NULL, stats = the forward's sums of x
// [B, C, 2], stats_bwd fp64 [B, 2C, 2] cleared and filled (dbeta[j] = sum_b [b, j, 0], dgamma[j] = sum_b [b, j, 1]).
extern "C" int fots_b200_instnorm_crelu_bwd_nhwc_bf16(const void* x, const void* y, const void* dy, const float* gamma, const double* stats,
                                                      double* stats_bwd, void* dx, int B, int HW, int C, float eps, float slope,
                                                      cudaStream_t stream) {
    if (!x || !y || !dy || !stats || !stats_bwd || !dx || B <= 0 || HW <= 0 || C <= 0 || C % 8 != 0 || C > 512 || B > 65535)
        return RROI_B200_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15)
        return RROI_B200_ERR_INVALID_ARG;
    const int G = C / 8, nphase = kThreads / G;
    cudaError_t e = pdl::zero_f64(stats_bwd, (size_t)B * 2 * C * 2, stream);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
    const Bf16x8 *xp = static_cast<const Bf16x8*>(x), *yp = static_cast<const Bf16x8*>(y), *gp = static_cast<const Bf16x8*>(dy);
    int rows = 0;
    const int chunks = wave_chunks(B, HW, resident_ctas(in_bwd_crelu_stats_kernel, 0), nphase * 4, nphase, &rows);
    e = pdl::launch(in_bwd_crelu_stats_kernel, dim3(chunks, B), dim3(kThreads), 0, stream, xp, yp, gp, stats, stats_bwd, HW, C, rows, eps, slope);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
    int arows = 0;
    const int ach = wave_chunks(B, HW, resident_ctas(in_bwd_crelu_apply_kernel, 0), nphase * 4, nphase, &arows);
    e = pdl::launch(in_bwd_crelu_apply_kernel, dim3(ach, B), dim3(kThreads), 0, stream, xp, yp, gp, gamma, stats, (const double*)stats_bwd,
                    static_cast<Bf16x8*>(dx), HW, C, arows, eps, slope);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
    return RROI_B200_OK;
}

// dlo [B, h, w, C] = U^T dhi [B, H, W, C]: the backward of the align_corners bilinear upsampling fots_b200_fpn_merge_nhwc_bf16
// computes for a_lo (training step; torch's channels-last upsample backward runs at ~0.55 TB/s).
extern "C" int fots_b200_upsample_bilinear_bwd_nhwc_bf16(const void* dhi, void* dlo, int B, int h, int w, int H, int W, int C,
                                                         cudaStream_t stream) {
    if (!dhi || !dlo || B <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 8 != 0) return RROI_B200_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(dhi) | reinterpret_cast<uintptr_t>(dlo)) & 15) return RROI_B200_ERR_INVALID_ARG;
    const long long total = (long long)B * h * w * (C / 8);
    long long grid = (total + kThreads - 1) / kThreads;
    const long long cap = 4LL * resident_ctas(upsample_bwd_kernel, 0);
    if (grid > cap) grid = cap;
    const cudaError_t e = pdl::launch(upsample_bwd_kernel, dim3((unsigned)grid), dim3(kThreads), 0, stream, static_cast<const Bf16x8*>(dhi),
                                      static_cast<Bf16x8*>(dlo), B, h, w, H, W, C);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
    return RROI_B200_OK;
}
