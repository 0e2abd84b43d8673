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
        for (int r = r0 + phase; r < r1; r += nphase) {
            float f[8];
            unpack8(ld8(base + (size_t)r * G), f);
#pragma unroll
            for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] = fmaf(f[i], f[i], q[i]); }
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
    for (int r = r0 + phase; r < r1; r += nphase) {
        float f[8], o[8];
        unpack8(ld8(x + (plane + r) * G + g), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = fmaf(f[i], scale[g * 8 + i], shift[g * 8 + i]);
        if (kResidual) {
            float rr[8];
            unpack8(ld8(res + (plane + r) * G + g), rr);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] += rr[i];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = o[i] > 0.0f ? o[i] : o[i] * slope;
        y[(plane + r) * Gout + g] = pack8(o);
        if (kCRelu) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float v = fmaf(f[i], scale[C + g * 8 + i], shift[C + g * 8 + i]);
                o[i] = v > 0.0f ? v : v * slope;
            }
            y[(plane + r) * Gout + G + g] = pack8(o);
        }
    }
}

}  // namespace

extern "C" int fots_b200_instnorm_nhwc_bf16(const void* x, void* y, const float* gamma, const float* beta,
                                            const void* residual, double* workspace, int B, int HW, int C,
                                            float eps, float slope, int crelu, cudaStream_t stream) {
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
    cudaError_t e = cudaMemsetAsync(workspace, 0, (size_t)B * C * 2 * sizeof(double), stream);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
    const dim3 grid(chunks, B);
    in_stats_kernel<<<grid, kThreads, 0, stream>>>(static_cast<const Bf16x8*>(x), workspace, HW, C, rows);
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
