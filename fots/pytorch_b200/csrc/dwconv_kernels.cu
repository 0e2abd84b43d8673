// dwconv_kernels.cu -- depthwise 3x3 convolution (groups == channels, pad 1, stride 1 or 2, no bias) of channels-last
// bf16 activations: the first halves of the depthwise-separable residual blocks of stages 3-4 and of the two
// up-convolutions of the top-down merge (/root/reference/tools/models.py:59-111 BasicBlockSepIn, :300-301 upconv1/2).
//
// 9 MACs per output value: pure bandwidth (read x once, write y once).  cuDNN's channels-last depthwise kernel
// (conv2d_c1_k1_nhwc_specialized) reaches about a quarter of the copy roofline on these shapes (18 us for a 14.7 MB map,
// 140 us for the 236 MB one); a first hand-written attempt in round 1 read its nine taps straight from global memory and
// was L1-bound.  Here a CTA stages its input tile (+ halo) for 64 channels in shared memory with cp.async (every pixel
// is 128 contiguous bytes; the whole tile is in flight at once, out-of-image vectors are zero-filled by the copy), and
// every thread produces a strip of 4 horizontally adjacent outputs for 8 channels, so each staged vector is read from
// shared memory once per (row, strip) instead of once per tap: 18 LDS.128 per 4 outputs at stride 1, 27 at stride 2.
// fp32 accumulation in the order (r, s) = (0,0) .. (2,2), one rounding to bf16.
#include "../../../include/fots_b200_pipeline.h"
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int kTW = 16;            // output pixels per tile row: 4 strips of 4
constexpr int kCB = 64;            // channels per CTA: 8 vectors of 8 bf16

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// NORM: the input is the RAW output z of the preceding convolution and the kernel computes dw(act(IN(z))) -- the
// InstanceNorm + leaky-ReLU that sits between the pointwise and the depthwise half of a separable block
// (tools/models.py:59-111) is applied to the staged tile in shared memory (once per staged vector, padding stays zero),
// from the per-image statistics of z (fots_b200_instnorm_stats_nhwc_bf16).  The normalised tensor never exists in HBM:
// one read + one write of the activation map less per block.
struct DwNorm {
    const double* stats;      // [N, C, 2] sum / sum of squares of z per image and channel
    const float* gamma;       // [C] or nullptr (affine = False)
    const float* beta;
    float eps, slope;
    double* stats_out;        // optional [N, C, 2]: sums of the bf16 outputs of THIS kernel (cleared by the host wrapper) --
                              // the statistics pass of the InstanceNorm that follows, for free in the epilogue
    int lh, lw;               // UP: x is a LOW-resolution map [N, lh, lw, C]; the tile is its bilinear (align_corners) upsampling
    float sy, sx;             //     to H x W, computed while staging (tools/models.py:418-436: upconv(F.interpolate(...)))
};

// torch's area_pixel_compute_source_index for align_corners = True (the same helper fots_b200_fpn_merge_nhwc_bf16 uses)
struct Lerp { int i0, i1; float l0, l1; };
__device__ __forceinline__ Lerp lerp_coord(int dst, int in, float scale) {
    Lerp r;
    const float s = scale * (float)dst;
    r.i0 = (int)s;
    r.i1 = r.i0 + ((r.i0 < in - 1) ? 1 : 0);
    r.l1 = s - (float)r.i0;
    r.l0 = 1.0f - r.l1;
    return r;
}

// grid (tiles_w * tiles_h, C / 64, N); block TH * 4 strips * 8 vectors.
template <int STRIDE, int TH, bool NORM, bool UP = false>
__global__ void __launch_bounds__(TH * 32) dwconv3x3_kernel(const uint4* __restrict__ x, const __nv_bfloat16* __restrict__ wgt,
                                                            uint4* __restrict__ y, int H, int W, int C, int Ho, int Wo, int tiles_w,
                                                            const DwNorm nrm) {
    constexpr int IH = (TH - 1) * STRIDE + 3, IW = (kTW - 1) * STRIDE + 3;      // input tile incl. halo
    constexpr int NT = TH * 32;
    __shared__ __align__(16) uint4 tile[IH * IW * 8];
    const int tile_id = blockIdx.x, ty = tile_id / tiles_w, tx = tile_id - ty * tiles_w;
    const int c0 = blockIdx.y * kCB, n = blockIdx.z;
    const int oy0 = ty * TH, ox0 = tx * kTW;
    const int iy0 = oy0 * STRIDE - 1, ix0 = ox0 * STRIDE - 1;
    const int CV = C / 8;                                                          // 16-byte vectors per pixel
    const uint4* img = x + (size_t)n * H * W * CV + c0 / 8;

    // ---- stage the input tile: thread -> (pixel, vector); consecutive threads = the 8 vectors (128 B) of one pixel
    const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);
    if (UP) {
        // the tile is the bilinear 2x upsampling of the low-resolution map, computed here instead of being written to and
        // read back from HBM by a separate kernel; same arithmetic and the same single bf16 rounding as that kernel
        const uint4* lo = x + (size_t)n * nrm.lh * nrm.lw * CV + c0 / 8;
        for (int i = threadIdx.x; i < IH * IW * 8; i += NT) {
            const int v = i & 7, p = i >> 3, r = p / IW, c = p - r * IW;
            const int iy = iy0 + r, ix = ix0 + c;
            uint4 pk = make_uint4(0, 0, 0, 0);
            if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
                const Lerp ly = lerp_coord(iy, nrm.lh, nrm.sy), lx = lerp_coord(ix, nrm.lw, nrm.sx);
                float v00[8], v01[8], v10[8], v11[8], o[8];
                unpack8(__ldg(lo + ((size_t)ly.i0 * nrm.lw + lx.i0) * CV + v), v00); unpack8(__ldg(lo + ((size_t)ly.i0 * nrm.lw + lx.i1) * CV + v), v01);
                unpack8(__ldg(lo + ((size_t)ly.i1 * nrm.lw + lx.i0) * CV + v), v10); unpack8(__ldg(lo + ((size_t)ly.i1 * nrm.lw + lx.i1) * CV + v), v11);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    o[k] = ly.l0 * (lx.l0 * v00[k] + lx.l1 * v01[k]) + ly.l1 * (lx.l0 * v10[k] + lx.l1 * v11[k]);
                pk.x = pack2(o[0], o[1]); pk.y = pack2(o[2], o[3]); pk.z = pack2(o[4], o[5]); pk.w = pack2(o[6], o[7]);
            }
            tile[i] = pk;
        }
    }
    for (int i = threadIdx.x; i < (UP ? 0 : IH * IW * 8); i += NT) {
        const int v = i & 7, p = i >> 3, r = p / IW, c = p - r * IW;
        const int iy = iy0 + r, ix = ix0 + c;
        const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
        const uint4* src = ok ? img + ((size_t)iy * W + ix) * CV + v : img;       // a valid address even when nothing is read
        const uint32_t nbytes = ok ? 16u : 0u;                                    // 0: the 16 bytes are zero-filled (the padding)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tile_s + (uint32_t)i * 16u), "l"(src), "r"(nbytes) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");

    // ---- this thread's 8 channels x 9 taps of weights (fp32 registers); w is [C][3][3]
    const int v = threadIdx.x & 7, strip = threadIdx.x >> 3;                       // strip = row * 4 + quarter
    const int row = strip >> 2, q4 = strip & 3;
    float wr[9][8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int t = 0; t < 9; ++t) wr[t][k] = __bfloat162float(wgt[(size_t)(c0 + v * 8 + k) * 9 + t]);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (NORM) {
        // scale / shift of this CTA's 64 channels for image n, then normalise + activate the staged tile in place
        __shared__ float coef[2 * kCB];
        if (threadIdx.x < kCB) {
            const int c = c0 + threadIdx.x;
            const double hw = (double)H * (double)W;
            const double m = nrm.stats[((size_t)n * C + c) * 2] / hw;
            double var = nrm.stats[((size_t)n * C + c) * 2 + 1] / hw - m * m;
            var = var < 0.0 ? 0.0 : var;
            const float rstd = rsqrtf((float)var + nrm.eps), mean = (float)m;
            const float g0 = nrm.gamma ? nrm.gamma[c] : 1.0f, b0 = nrm.beta ? nrm.beta[c] : 0.0f;
            coef[threadIdx.x] = rstd * g0;
            coef[kCB + threadIdx.x] = b0 - mean * rstd * g0;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < IH * IW * 8; i += NT) {
            const int vv = i & 7, p = i >> 3, r = p / IW, c = p - r * IW;
            const int iy = iy0 + r, ix = ix0 + c;
            if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;                 // padding: stays exactly zero
            float f[8];
            unpack8(tile[i], f);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float t = fmaf(f[k], coef[vv * 8 + k], coef[kCB + vv * 8 + k]);
                f[k] = t > 0.0f ? t : t * nrm.slope;
            }
            uint4 pk;                                                              // one bf16 rounding, as the separate apply pass stores it
            pk.x = pack2(f[0], f[1]); pk.y = pack2(f[2], f[3]); pk.z = pack2(f[4], f[5]); pk.w = pack2(f[6], f[7]);
            tile[i] = pk;
        }
        __syncthreads();
    }

    float acc[4][8];
#pragma unroll
    for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[o][k] = 0.0f;
    constexpr int NCOL = 3 * STRIDE + 3;                                           // input columns a strip of 4 outputs touches
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const uint4* trow = tile + ((row * STRIDE + r) * IW + q4 * 4 * STRIDE) * 8 + v;
#pragma unroll
        for (int j = 0; j < NCOL; ++j) {
            float f[8];
            unpack8(trow[j * 8], f);
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const int s = j - o * STRIDE;                                       // tap column of output o fed by input column j
                if (s >= 0 && s < 3) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[o][k] = fmaf(f[k], wr[r * 3 + s][k], acc[o][k]);
                }
            }
        }
    }
    const int oy = oy0 + row;
    float ssum[8], ssq[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) ssum[k] = ssq[k] = 0.0f;
    if (oy < Ho) {
        uint4* out = y + (((size_t)n * Ho + oy) * Wo) * CV + c0 / 8 + v;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            const int ox = ox0 + q4 * 4 + o;
            if (ox < Wo) {
                uint4 pk;
                pk.x = pack2(acc[o][0], acc[o][1]); pk.y = pack2(acc[o][2], acc[o][3]);
                pk.z = pack2(acc[o][4], acc[o][5]); pk.w = pack2(acc[o][6], acc[o][7]);
                out[(size_t)ox * CV] = pk;
                if (nrm.stats_out) {                                     // of the ROUNDED values, as a separate pass would see them
                    float f[8];
                    unpack8(pk, f);
#pragma unroll
                    for (int k = 0; k < 8; ++k) { ssum[k] += f[k]; ssq[k] = fmaf(f[k], f[k], ssq[k]); }
                }
            }
        }
    }
    if (nrm.stats_out) {                                                 // CTA-uniform
        __syncthreads();                                                 // everybody is done reading the tile: reuse it
        float* red = reinterpret_cast<float*>(tile);                    // [NT][17]
#pragma unroll
        for (int k = 0; k < 8; ++k) { red[threadIdx.x * 17 + k] = ssum[k]; red[threadIdx.x * 17 + 8 + k] = ssq[k]; }
        __syncthreads();
        if (threadIdx.x < 128) {                                         // 8 vectors x 16 values
            const int vv = threadIdx.x >> 4, k = threadIdx.x & 15;
            float a = 0.0f;
            for (int st = 0; st < NT / 8; ++st) a += red[(st * 8 + vv) * 17 + k];
            atomicAdd(nrm.stats_out + ((size_t)n * C + c0 + vv * 8 + (k & 7)) * 2 + (k >> 3), (double)a);
        }
    }
}

}  // namespace

static int dw_launch(const void* x, const void* w, void* y, int N, int H, int W, int C, int stride, const DwNorm* nrm, cudaStream_t stream) {
    if (!x || !w || !y || N <= 0 || H <= 0 || W <= 0 || C <= 0 || C % kCB != 0 || (stride != 1 && stride != 2) || N > 65535 ||
        C / kCB > 65535)
        return RROI_B200_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) return RROI_B200_ERR_INVALID_ARG;
    const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
    const int tiles_w = (Wo + kTW - 1) / kTW;
    const uint4* xp = static_cast<const uint4*>(x);
    const __nv_bfloat16* wp = static_cast<const __nv_bfloat16*>(w);
    uint4* yp = static_cast<uint4*>(y);
    DwNorm none = {nullptr, nullptr, nullptr, 0.f, 1.f, nullptr, 0, 0, 0.f, 0.f};
    if (nrm) none = *nrm;
    if (none.stats_out) {
        const cudaError_t em = cudaMemsetAsync(none.stats_out, 0, (size_t)N * C * 2 * sizeof(double), stream);
        if (em != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
    }
    const bool norm_on = nrm != nullptr && nrm->stats != nullptr;
    if (none.lh > 0) {                                        // upsample-on-load: stride 1, no normalisation
        if (stride != 1 || norm_on) return RROI_B200_ERR_INVALID_ARG;
        constexpr int TH = 8;
        const dim3 grid((unsigned)(tiles_w * ((Ho + TH - 1) / TH)), (unsigned)(C / kCB), (unsigned)N);
        dwconv3x3_kernel<1, TH, false, true><<<grid, TH * 32, 0, stream>>>(xp, wp, yp, H, W, C, Ho, Wo, tiles_w, none);
        const cudaError_t eu = cudaGetLastError();
        if (eu != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
        return RROI_B200_OK;
    }
    if (stride == 1) {
        constexpr int TH = 8;
        const dim3 grid((unsigned)(tiles_w * ((Ho + TH - 1) / TH)), (unsigned)(C / kCB), (unsigned)N);
        if (norm_on) dwconv3x3_kernel<1, TH, true><<<grid, TH * 32, 0, stream>>>(xp, wp, yp, H, W, C, Ho, Wo, tiles_w, none);
        else dwconv3x3_kernel<1, TH, false><<<grid, TH * 32, 0, stream>>>(xp, wp, yp, H, W, C, Ho, Wo, tiles_w, none);
    } else {
        constexpr int TH = 4;
        const dim3 grid((unsigned)(tiles_w * ((Ho + TH - 1) / TH)), (unsigned)(C / kCB), (unsigned)N);
        if (norm_on) dwconv3x3_kernel<2, TH, true><<<grid, TH * 32, 0, stream>>>(xp, wp, yp, H, W, C, Ho, Wo, tiles_w, none);
        else dwconv3x3_kernel<2, TH, false><<<grid, TH * 32, 0, stream>>>(xp, wp, yp, H, W, C, Ho, Wo, tiles_w, none);
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
    return RROI_B200_OK;
}

extern "C" int fots_b200_dwconv3x3_nhwc_bf16(const void* x, const void* w, void* y, int N, int H, int W, int C, int stride,
                                             cudaStream_t stream) {
    return dw_launch(x, w, y, N, H, W, C, stride, nullptr, stream);
}

extern "C" int fots_b200_dwconv3x3_norm_nhwc_bf16(const void* x, const void* w, void* y, const double* stats, const float* gamma,
                                                  const float* beta, float eps, float slope, double* stats_out, int N, int H, int W,
                                                  int C, int stride, cudaStream_t stream) {
    if ((gamma == nullptr) != (beta == nullptr) || (!stats && (gamma || beta))) return RROI_B200_ERR_INVALID_ARG;
    const DwNorm nrm = {stats, gamma, beta, eps, slope, stats_out, 0, 0, 0.f, 0.f};
    return dw_launch(x, w, y, N, H, W, C, stride, &nrm, stream);
}

// dw(upsample(x_lo)): x_lo bf16 [N, h, w, C] -> bilinear (align_corners = True) upsampling to H x W computed while the tile
// is staged -> depthwise 3x3 stride 1 -> y bf16 [N, H, W, C].  The upsampled map (the largest tensor of the top-down merge:
// 236 MB per 8 images at 1/4 scale) is never written.
extern "C" int fots_b200_dwconv3x3_up_nhwc_bf16(const void* x_lo, const void* w, void* y, int N, int h, int wlo, int H, int W, int C,
                                                cudaStream_t stream) {
    if (h <= 0 || wlo <= 0) return RROI_B200_ERR_INVALID_ARG;
    DwNorm nrm = {nullptr, nullptr, nullptr, 0.f, 1.f, nullptr, h, wlo, 0.f, 0.f};
    nrm.sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.0f;
    nrm.sx = W > 1 ? (float)(wlo - 1) / (float)(W - 1) : 0.0f;
    return dw_launch(x_lo, w, y, N, H, W, C, 1, &nrm, stream);
}
