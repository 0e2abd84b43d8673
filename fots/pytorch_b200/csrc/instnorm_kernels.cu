// instnorm_kernels.cu -- fused channels-last InstanceNorm + affine + residual + leaky-ReLU (bf16 in/out, fp32/fp64
// statistics) for the feeder/consumer networks.  See include/fots_b200_pipeline.h.  HBM-bound: the tensor is read
// twice and written once; every access is a 16-byte vector, consecutive threads on consecutive channel groups.
#include "../../../include/fots_b200_pipeline.h"
#include <cuda_bf16.h>
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
__global__ void __launch_bounds__(kThreads) in_stats_kernel(const Bf16x8* __restrict__ x, double* __restrict__ ws,
                                                             int HW, int C, int rows_per_cta) {
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
__global__ void __launch_bounds__(kThreads) in_apply_kernel(const Bf16x8* __restrict__ x, Bf16x8* __restrict__ y,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             const Bf16x8* __restrict__ res, const double* __restrict__ ws,
                                                             int HW, int C, int rows_per_cta, float eps, float slope) {
    extern __shared__ float coef[];          // scale[Cout], shift[Cout]
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
    const int g = threadIdx.x % G, phase = threadIdx.x / G, nphase = kThreads / G;
    if (phase >= nphase) return;
    const int r0 = blockIdx.x * rows_per_cta;
    const int r1 = min(HW, r0 + rows_per_cta);
    const size_t plane = (size_t)b * HW;
    const int Gout = Cout / 8;
    // this thread's 8 channels: coefficients in registers, U rows in flight (memory-level parallelism)
    float sc[8], sh[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { sc[i] = scale[g * 8 + i]; sh[i] = shift[g * 8 + i]; }
    const float* sc2 = scale + C + g * 8;                 // CReLU's second half: read from shared memory (saves 16 registers)
    const float* sh2 = shift + C + g * 8;
    constexpr int U = 4;
    for (int r = r0 + phase; r < r1; r += U * nphase) {
        Bf16x8 xv[U], rv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int rr_ = r + u * nphase;
            if (rr_ < r1) {
                xv[u] = ld8(x + (plane + rr_) * G + g);
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
            if (kCRelu) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float v = fmaf(f[i], sc2[i], sh2[i]);
                    o[i] = v > 0.0f ? v : v * slope;
                }
                y[(plane + rr_) * Gout + G + g] = pack8(o);
            }
        }
    }
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

__global__ void __launch_bounds__(kThreads) fpn_merge_kernel(const Bf16x8* __restrict__ a_lo, const Bf16x8* __restrict__ c_hi,
                                                              const Bf16x8* __restrict__ b_hi, const __nv_bfloat16* __restrict__ g_lo,
                                                              Bf16x8* __restrict__ y, int B, int h, int w, int H, int W, int C) {
    const int G = C / 8;
    const long long total = (long long)B * H * W * G;
    const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.0f;
    const float sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.0f;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
        const int g = (int)(i % G);
        long long p = i / G;
        const int X = (int)(p % W); p /= W;
        const int Y = (int)(p % H);
        const int b = (int)(p / H);
        const Lerp ly = lerp_coord(Y, h, sy), lx = lerp_coord(X, w, sx);
        const long long lo00 = ((long long)b * h + ly.i0) * w + lx.i0, lo01 = ((long long)b * h + ly.i0) * w + lx.i1;
        const long long lo10 = ((long long)b * h + ly.i1) * w + lx.i0, lo11 = ((long long)b * h + ly.i1) * w + lx.i1;
        float o[8];
        if (a_lo) {
            float v00[8], v01[8], v10[8], v11[8];
            unpack8(ld8(a_lo + lo00 * G + g), v00); unpack8(ld8(a_lo + lo01 * G + g), v01);
            unpack8(ld8(a_lo + lo10 * G + g), v10); unpack8(ld8(a_lo + lo11 * G + g), v11);
#pragma unroll
            for (int k = 0; k < 8; ++k)
                o[k] = ly.l0 * (lx.l0 * v00[k] + lx.l1 * v01[k]) + ly.l1 * (lx.l0 * v10[k] + lx.l1 * v11[k]);
        } else {
            unpack8(ld8(c_hi + i), o);
        }
        if (b_hi) {
            float gate = 1.0f;
            if (g_lo) {
                const float s00 = 1.0f / (1.0f + __expf(-__bfloat162float(g_lo[lo00])));
                const float s01 = 1.0f / (1.0f + __expf(-__bfloat162float(g_lo[lo01])));
                const float s10 = 1.0f / (1.0f + __expf(-__bfloat162float(g_lo[lo10])));
                const float s11 = 1.0f / (1.0f + __expf(-__bfloat162float(g_lo[lo11])));
                gate = ly.l0 * (lx.l0 * s00 + lx.l1 * s01) + ly.l1 * (lx.l0 * s10 + lx.l1 * s11);
            }
            float bb[8];
            unpack8(ld8(b_hi + i), bb);
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = fmaf(bb[k], gate, o[k]);
        }
        y[i] = pack8(o);
    }
}

// ---- max-pool (2,1)/(2,1) over H of a channels-last tensor (tools/models.py:344, :360 `max2`) --------------------
// thread = 8 channels of one output pixel; the two input rows are W*C elements apart.  NaNs propagate like torch's.
__global__ void __launch_bounds__(kThreads) maxpool_h2_kernel(const uint4* __restrict__ x, uint4* __restrict__ y,
                                                               long long total, int Ho, long long rowv, int H) {
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

}  // namespace

extern "C" int fots_b200_maxpool_h2_nhwc_bf16(const void* x, void* y, int N, int H, int W, int C, cudaStream_t stream) {
    if (!x || !y || N <= 0 || H < 2 || W <= 0 || C <= 0 || C % 8 != 0) return RROI_B200_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) return RROI_B200_ERR_INVALID_ARG;
    const int Ho = H / 2;
    const long long rowv = (long long)W * (C / 8);
    const long long total = (long long)N * Ho * rowv;
    long long grid = (total + kThreads - 1) / kThreads;
    if (grid > 148LL * 16) grid = 148LL * 16;
    maxpool_h2_kernel<<<(unsigned)grid, kThreads, 0, stream>>>(static_cast<const uint4*>(x), static_cast<uint4*>(y), total, Ho, rowv, H);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
    return RROI_B200_OK;
}

extern "C" int fots_b200_fpn_merge_nhwc_bf16(const void* a_lo, const void* c_hi, const void* b_hi, const void* g_lo,
                                             void* y, int B, int h, int w, int H, int W, int C, cudaStream_t stream) {
    if (!y || ((a_lo == nullptr) == (c_hi == nullptr)) || (g_lo && !b_hi) || B <= 0 || h <= 0 || w <= 0 || H <= 0 ||
        W <= 0 || C <= 0 || C % 8 != 0)
        return RROI_B200_ERR_INVALID_ARG;
    const long long total = (long long)B * H * W * (C / 8);
    long long grid = (total + kThreads - 1) / kThreads;
    if (grid > 148LL * 32) grid = 148LL * 32;
    fpn_merge_kernel<<<(unsigned)grid, kThreads, 0, stream>>>(static_cast<const Bf16x8*>(a_lo), static_cast<const Bf16x8*>(c_hi),
                                                              static_cast<const Bf16x8*>(b_hi), static_cast<const __nv_bfloat16*>(g_lo),
                                                              static_cast<Bf16x8*>(y), B, h, w, H, W, C);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
    return RROI_B200_OK;
}

static int instnorm_impl(const void* x, void* y, const float* gamma, const float* beta, const void* residual,
                         double* workspace, int B, int HW, int C, float eps, float slope, int crelu, bool have_stats,
                         cudaStream_t stream) {
    if (!x || !y || !workspace || B <= 0 || HW <= 0 || C <= 0 || C % 8 != 0 || C > 1024 || (C / 8) > kThreads ||
        ((gamma == nullptr) != (beta == nullptr)) || (crelu && residual))
        return RROI_B200_ERR_INVALID_ARG;
    const int G = C / 8;
    const int nphase = kThreads / G;
    // enough CTAs for a few waves, at least a handful of rows per thread
    long long want_ctas = 148LL * 8 / B + 1;
    int rows = (int)((HW + want_ctas - 1) / want_ctas);
    rows = ((rows + nphase - 1) / nphase) * nphase;
    if (rows < nphase * 4) rows = nphase * 4;
    const int chunks = (HW + rows - 1) / rows;
    cudaError_t e = cudaSuccess;
    const dim3 grid(chunks, B);
    if (!have_stats) {
        e = cudaMemsetAsync(workspace, 0, (size_t)B * C * 2 * sizeof(double), stream);
        if (e != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
        in_stats_kernel<<<grid, kThreads, 0, stream>>>(static_cast<const Bf16x8*>(x), workspace, HW, C, rows);
    }
    const size_t smem = (size_t)(crelu ? 4 : 2) * C * sizeof(float);
    const Bf16x8* xr = static_cast<const Bf16x8*>(x);
    Bf16x8* yr = static_cast<Bf16x8*>(y);
    const Bf16x8* rr = static_cast<const Bf16x8*>(residual);
    if (crelu)
        in_apply_kernel<true, false><<<grid, kThreads, smem, stream>>>(xr, yr, gamma, beta, nullptr, workspace, HW, C, rows, eps, slope);
    else if (residual)
        in_apply_kernel<false, true><<<grid, kThreads, smem, stream>>>(xr, yr, gamma, beta, rr, workspace, HW, C, rows, eps, slope);
    else
        in_apply_kernel<false, false><<<grid, kThreads, smem, stream>>>(xr, yr, gamma, beta, nullptr, workspace, HW, C, rows, eps, slope);
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
