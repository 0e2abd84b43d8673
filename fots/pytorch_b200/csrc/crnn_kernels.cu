// crnn_kernels.cu -- the two pieces of consumer B's CNN (/root/reference/tools/models.py:853-897, CRNN.cnn) that cannot go
// through the tcgen05 convolution: the FIRST layer (3 input channels: K = 27 does not fill a 64-wide k-block) and the
// max-poolings with rectangular windows / strides / padding ((2,2)/(2,2) and (2,2)/(2,1) pad (0,1)).  Both are tiny next to
// the other six convolutions (1.8 of 150 GFLOP for 64 crops of 32 x 256) -- they exist so that the whole consumer runs on this
// repository's kernels, not for speed.
//
//   fots_b200_conv3x3_c3_pool_nhwc_bf16: x fp32 NCHW [N, 3, H, W] (what RoIRotate of the raw image returns) ->
//       relu(conv3x3(x, w) + bias) -> optional 2x2 / stride 2 max-pool -> y bf16 NHWC [N, H', W', Cout].
//       Thread = one OUTPUT pixel x 8 output channels; the (<= 4 x 4 x 3) input patch is read once into registers, the
//       weights ([27][Cout] fp32) and the bias sit in shared memory.  fp32 accumulation in the order (cin, r, s).
//   fots_b200_maxpool_nhwc_bf16: general MaxPool2d of a channels-last bf16 tensor, thread = one output pixel x 8 channels,
//       padding behaves as -inf (torch), NaNs propagate.
#include "../../../include/fots_b200_pipeline.h"
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// grid-stride over (output pixel, channel group); consecutive threads = consecutive channel groups of one pixel.
template <bool POOL>
__global__ void __launch_bounds__(kThreads) conv3x3_c3_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                                             const float* __restrict__ bias, uint4* __restrict__ y, int N, int H, int W,
                                                             int Cout) {
    extern __shared__ float wsm[];                              // [27][Cout] then bias [Cout]
    float* bsm = wsm + 27 * Cout;
    for (int i = threadIdx.x; i < 27 * Cout; i += kThreads) {   // w is [Cout][3][3][3] (cout, cin, r, s) contiguous
        const int co = i / 27, k = i - co * 27;
        wsm[k * Cout + co] = __bfloat162float(w[i]);
    }
    for (int i = threadIdx.x; i < Cout; i += kThreads) bsm[i] = bias ? bias[i] : 0.0f;
    __syncthreads();
    constexpr int P = POOL ? 2 : 1;                             // conv positions per output pixel and axis
    const int Ho = H / P, Wo = W / P, G = Cout / 8;
    const long long total = (long long)N * Ho * Wo * G;
    for (long long idx = (long long)blockIdx.x * kThreads + threadIdx.x; idx < total; idx += (long long)gridDim.x * kThreads) {
        const int g = (int)(idx % G);
        long long p = idx / G;
        const int ox = (int)(p % Wo); p /= Wo;
        const int oy = (int)(p % Ho);
        const int n = (int)(p / Ho);
        // input patch rows [oy*P - 1, oy*P + P], cols [ox*P - 1, ox*P + P], zero outside the image (the padding)
        float patch[3][P + 2][P + 2];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int r = 0; r < P + 2; ++r)
#pragma unroll
                for (int s = 0; s < P + 2; ++s) {
                    const int iy = oy * P - 1 + r, ix = ox * P - 1 + s;
                    patch[c][r][s] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(x + (((size_t)n * 3 + c) * H + iy) * W + ix) : 0.0f;
                }
        float best[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) best[k] = 0.0f;             // relu(.) >= 0, so 0 is the identity of max over the window
#pragma unroll
        for (int py = 0; py < P; ++py)
#pragma unroll
            for (int px = 0; px < P; ++px) {
                float acc[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] = bsm[g * 8 + k];
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int s = 0; s < 3; ++s) {
                            const float v = patch[c][py + r][px + s];
                            const float4 w0 = *reinterpret_cast<const float4*>(wsm + ((c * 3 + r) * 3 + s) * Cout + g * 8);
                            const float4 w1 = *reinterpret_cast<const float4*>(wsm + ((c * 3 + r) * 3 + s) * Cout + g * 8 + 4);
                            acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]); acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
                            acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]); acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
                        }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    // torch: relu then max-pool; a NaN anywhere in the window must survive both
                    const float a = acc[k];
                    best[k] = (a != a) ? a : ((best[k] != best[k]) ? best[k] : fmaxf(best[k], a));
                }
            }
        uint4 pk;
        pk.x = pack2(best[0], best[1]); pk.y = pack2(best[2], best[3]); pk.z = pack2(best[4], best[5]); pk.w = pack2(best[6], best[7]);
        y[idx] = pk;
    }
}

// thread = one output pixel x 8 channels.
__global__ void __launch_bounds__(kThreads) maxpool_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int N, int H, int W, int G,
                                                          int Ho, int Wo, int kh, int kw, int sh, int sw, int ph, int pw) {
    const long long total = (long long)N * Ho * Wo * G;
    for (long long idx = (long long)blockIdx.x * kThreads + threadIdx.x; idx < total; idx += (long long)gridDim.x * kThreads) {
        const int g = (int)(idx % G);
        long long p = idx / G;
        const int ox = (int)(p % Wo); p /= Wo;
        const int oy = (int)(p % Ho);
        const int n = (int)(p / Ho);
        __nv_bfloat162 best[4];
        bool have = false;
        for (int r = 0; r < kh; ++r) {
            const int iy = oy * sh - ph + r;
            if (iy < 0 || iy >= H) continue;
            for (int s = 0; s < kw; ++s) {
                const int ix = ox * sw - pw + s;
                if (ix < 0 || ix >= W) continue;
                const uint4 raw = __ldg(x + (((size_t)n * H + iy) * W + ix) * G + g);
                const __nv_bfloat162* v = reinterpret_cast<const __nv_bfloat162*>(&raw);
                if (!have) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) best[k] = v[k];
                    have = true;
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) best[k] = __hmax2_nan(best[k], v[k]);
                }
            }
        }
        uint4 out;
        __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&out);
        const __nv_bfloat162 ninf = __floats2bfloat162_rn(-INFINITY, -INFINITY);        // window entirely in the padding
#pragma unroll
        for (int k = 0; k < 4; ++k) po[k] = have ? best[k] : ninf;
        y[idx] = out;
    }
}

}  // namespace

extern "C" int fots_b200_conv3x3_c3_pool_nhwc_bf16(const float* x, const void* w, const float* bias, void* y, int N, int H, int W, int Cout,
                                                   int pool2x2, cudaStream_t stream) {
    if (!x || !w || !y || N <= 0 || H <= 0 || W <= 0 || Cout <= 0 || Cout % 8 != 0 || Cout > 256) return RROI_B200_ERR_INVALID_ARG;
    if (pool2x2 && (H < 2 || W < 2)) return RROI_B200_ERR_INVALID_ARG;
    if (reinterpret_cast<uintptr_t>(y) & 15) return RROI_B200_ERR_INVALID_ARG;
    const int P = pool2x2 ? 2 : 1;
    const long long total = (long long)N * (H / P) * (W / P) * (Cout / 8);
    long long grid = (total + kThreads - 1) / kThreads;
    if (grid > 148LL * 8) grid = 148LL * 8;
    const size_t smem = (size_t)28 * Cout * sizeof(float);
    const __nv_bfloat16* wp = static_cast<const __nv_bfloat16*>(w);
    if (pool2x2) conv3x3_c3_kernel<true><<<(unsigned)grid, kThreads, smem, stream>>>(x, wp, bias, static_cast<uint4*>(y), N, H, W, Cout);
    else conv3x3_c3_kernel<false><<<(unsigned)grid, kThreads, smem, stream>>>(x, wp, bias, static_cast<uint4*>(y), N, H, W, Cout);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
    return RROI_B200_OK;
}

extern "C" int fots_b200_maxpool_nhwc_bf16(const void* x, void* y, int N, int H, int W, int C, int kh, int kw, int sh, int sw, int ph,
                                           int pw, cudaStream_t stream) {
    if (!x || !y || N <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 8 != 0 || kh <= 0 || kw <= 0 || sh <= 0 || sw <= 0 || ph < 0 || pw < 0 ||
        2 * ph > kh || 2 * pw > kw)                                  // torch: pad <= kernel / 2
        return RROI_B200_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) return RROI_B200_ERR_INVALID_ARG;
    const int Ho = (H + 2 * ph - kh) / sh + 1, Wo = (W + 2 * pw - kw) / sw + 1;     // floor mode
    if (Ho <= 0 || Wo <= 0) return RROI_B200_ERR_INVALID_ARG;
    const long long total = (long long)N * Ho * Wo * (C / 8);
    long long grid = (total + kThreads - 1) / kThreads;
    if (grid > 148LL * 16) grid = 148LL * 16;
    maxpool_kernel<<<(unsigned)grid, kThreads, 0, stream>>>(static_cast<const uint4*>(x), static_cast<uint4*>(y), N, H, W, C / 8, Ho, Wo, kh,
                                                            kw, sh, sw, ph, pw);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
    return RROI_B200_OK;
}
