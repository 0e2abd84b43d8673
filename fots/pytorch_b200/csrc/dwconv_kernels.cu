// dwconv_kernels.cu -- depthwise 3x3 convolution (groups == channels, pad 1, stride 1 or 2, no bias) of channels-last
// bf16 activations: the first halves of the depthwise-separable residual blocks of stages 3-4 and of the two
// up-convolutions of the top-down merge (/root/reference/tools/models.py:59-111 BasicBlockSepIn, :300-301 upconv1/2).
//
// 9 MACs per output value: pure bandwidth (read x once, write y once).  cuDNN's channels-last depthwise kernel
// (conv2d_c1_k1_nhwc_specialized) reaches about a quarter of the copy roofline on these shapes (18 us for a 14.7 MB map,
// 140 us for the 236 MB one); a first hand-written attempt in round 1 read its nine taps straight from global memory and
// was L1-bound.  A CTA stages its input tile (+ halo) for 64 channels in shared memory (every pixel is 128 contiguous
// bytes, out-of-image pixels arrive as zeros = the padding) and every thread produces a strip of 4 horizontally adjacent
// outputs for 8 channels, so each staged vector is read from shared memory once per (row, strip) instead of once per tap:
// 18 LDS.128 per 4 outputs at stride 1, 27 at stride 2.  fp32 accumulation in the order (r, s) = (0,0) .. (2,2), one
// rounding to bf16.
//   stride 2: one tile per CTA, staged with cp.async (dwconv3x3_kernel; two small launches per step);
//   stride 1: persistent CTAs, tiles staged by TMA box loads into an mbarrier ring, mixed-precision FMAs on the packed bf16
//             operands (dwconv3x3_s1_pipe_kernel: 0.91 of the copy roofline on the 236 MB map), with the InstanceNorm in
//             front of it or the bilinear upsampling of the top-down merge applied to the staged tile (NORM / UP modes).
#include "../../../include/fots_b200_pipeline.h"
#include "pdl.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <mutex>

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

// One tile per CTA (stride 2; the stride-1 shapes take the persistent pipelined kernel below).
// grid (tiles_w * tiles_h, C / 64, N); block TH * 4 strips * 8 vectors.
template <int STRIDE, int TH, bool NORM>
__global__ void __launch_bounds__(TH * 32) dwconv3x3_kernel(const uint4* __restrict__ x, const __nv_bfloat16* __restrict__ wgt,
                                                            uint4* __restrict__ y, int H, int W, int C, int Ho, int Wo, int tiles_w,
                                                            const DwNorm nrm) {
    constexpr int IH = (TH - 1) * STRIDE + 3, IW = (kTW - 1) * STRIDE + 3;      // input tile incl. halo
    constexpr int NT = TH * 32;
    __shared__ __align__(16) uint4 tile[IH * IW * 8];
    pdl::trigger();
    pdl::wait();
    const int tile_id = blockIdx.x, ty = tile_id / tiles_w, tx = tile_id - ty * tiles_w;
    const int c0 = blockIdx.y * kCB, n = blockIdx.z;
    const int oy0 = ty * TH, ox0 = tx * kTW;
    const int iy0 = oy0 * STRIDE - 1, ix0 = ox0 * STRIDE - 1;
    const int CV = C / 8;                                                          // 16-byte vectors per pixel
    const uint4* img = x + (size_t)n * H * W * CV + c0 / 8;

    // ---- stage the input tile: thread -> (pixel, vector); consecutive threads = the 8 vectors (128 B) of one pixel
    const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);
    for (int i = threadIdx.x; i < IH * IW * 8; i += NT) {
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

// ---- TMA staging (cp.async.bulk.tensor + mbarrier), used by the persistent stride-1 kernel -------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Spin on the phase parity; trap after ~2 s instead of hanging the device if a copy never completes.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    long long t0 = 0;
    for (uint32_t spin = 0;; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return;
        if ((spin & 0xfff) == 0xfff) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000LL) __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// ---- stride 1: persistent, software-pipelined ---------------------------------------------------------------------
// The one-tile-per-CTA form above loads, waits, computes, stores: with three CTAs per SM about a third of the SM's tiles
// are in flight at any time and the kernel sits at 0.4 of the copy roofline.  Here a CTA owns ONE 64-channel block (its
// 72 weights per thread stay in registers) and walks over that block's tiles.  A tile (18 x 10 pixels x 64 channels incl. halo;
// UP: the 12 x 8 low-resolution footprint) is ONE TMA box load: thread 0 issues cp.async.bulk.tensor for the tile kStages - 1
// ahead into a ring of shared-memory buffers (mbarrier completion; out-of-image rows / columns are zero-filled by the copy
// engine = the padding), so loads are always in flight and nobody spends issue slots on per-vector address arithmetic (the
// cp.async form: ~90 of ~400 instructions per thread and tile).  The 3x3 accumulation uses packed fp32
// FMAs (fma.rn.f32x2: each half an ordinary IEEE fp32 FMA in the same (r, s) order -> bit-identical to the scalar chain).
constexpr int kTH = 8, kIH = kTH + 2, kIW = kTW + 2;                 // 16 x 8 outputs from an 18 x 10 input tile
constexpr int kTileVec = kIH * kIW * 8;                              // 1 440 uint4 = 23 040 B
constexpr int kLoH = 8, kLoW = 12, kLoVec = kLoH * kLoW * 8;          // UP: low-resolution footprint of a tile (2x: 7 x 11), 12 288 B
constexpr int kNT = 256;
constexpr int kStages = 3;                                           // ring depth of the TMA staging
enum { kPlain = 0, kNorm = 1, kUp = 2 };

__device__ __forceinline__ uint64_t pair_f32(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ uint64_t widen2(uint32_t w) { return pair_f32(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// {acc_lo, acc_hi} += {x.lo * w.lo, x.hi * w.hi}: two mixed-precision FMAs on the halves of packed bf16 pairs
__device__ __forceinline__ void fma_bf16x2(float& acc_lo, float& acc_hi, uint32_t x, uint32_t w) {
    asm("{\n\t.reg .b16 xl, xh, wl, wh;\n\t"
        "mov.b32 {xl, xh}, %2;\n\tmov.b32 {wl, wh}, %3;\n\t"
        "fma.rn.f32.bf16 %0, xl, wl, %0;\n\tfma.rn.f32.bf16 %1, xh, wh, %1;\n\t}"
        : "+f"(acc_lo), "+f"(acc_hi) : "r"(x), "r"(w));
}
__device__ __forceinline__ void split2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }

struct __align__(16) Tap { int o0, o1; float l0, l1; };
struct TilePos { int n, oy0, ox0; };
__device__ __forceinline__ TilePos tile_pos(int t, int tiles_w, int tiles_h) {
    TilePos p;
    const int tx = t % tiles_w;
    t /= tiles_w;
    p.n = t / tiles_h;
    p.oy0 = (t - p.n * tiles_h) * kTH;
    p.ox0 = tx * kTW;
    return p;
}

// Low-resolution footprint of a tile for the UP mode: rows [y0, y0 + nh), columns [x0, x0 + nw); fits = it can be staged.
struct LoBox { int y0, x0, nh, nw; bool fits; };
__device__ __forceinline__ LoBox lo_box(const TilePos& tp, int H, int W, const DwNorm& nrm) {
    LoBox b;
    const int ya = max(tp.oy0 - 1, 0), yb = min(tp.oy0 + kTH, H - 1), xa = max(tp.ox0 - 1, 0), xb = min(tp.ox0 + kTW, W - 1);
    b.y0 = lerp_coord(ya, nrm.lh, nrm.sy).i0;
    b.x0 = lerp_coord(xa, nrm.lw, nrm.sx).i0;
    b.nh = lerp_coord(yb, nrm.lh, nrm.sy).i1 - b.y0 + 1;
    b.nw = lerp_coord(xb, nrm.lw, nrm.sx).i1 - b.x0 + 1;
    b.fits = b.nh <= kLoH && b.nw <= kLoW;
    return b;
}

// grid (CTAs per channel block, C / 64); block 256 = 8 rows x 4 strips x 8 vectors.  Dynamic shared memory:
//   plain / NORM: two input tiles (+ NORM: 128 coefficients, + stats_out: the [256][17] reduction scratch)
//   UP:           two low-resolution footprints + one (upsampled) input tile
template <int MODE>
__global__ void __launch_bounds__(kNT, 2) dwconv3x3_s1_pipe_kernel(const __grid_constant__ CUtensorMap map_x, const uint4* __restrict__ x,
                                                                   const __nv_bfloat16* __restrict__ wgt, uint4* __restrict__ y, int H, int W,
                                                                   int C, int tiles_w, int tiles_h, int ntiles, const DwNorm nrm) {
    extern __shared__ __align__(128) uint4 dsm[];
    __shared__ __align__(8) unsigned long long bars[kStages];
    constexpr int kStageVec = MODE == kUp ? kLoVec : kTileVec;
    uint4* const stage0 = dsm;                                                    // [kStages][kStageVec]
    uint4* const hi_tile = dsm + kStages * kLoVec;                                // UP only
    float* const coef = reinterpret_cast<float*>(dsm + kStages * kTileVec);      // NORM only: scale[64], shift[64]
    float* const red = coef + 2 * kCB;                                            // NORM + stats_out only: [256][17]
    const int c0 = blockIdx.y * kCB;
    const int CV = C / 8;
    const int tid = threadIdx.x;
    const int v = tid & 7, strip = tid >> 3, row = strip >> 2, q4 = strip & 3;
    const uint32_t stage_s = smem_u32(stage0);
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < kStages; ++k) mbar_init(smem_u32(&bars[k]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // ---- issue of one tile's box load (thread 0; the slot was released by the __syncthreads() that ended its last use).
    // UP: the footprint box always starts at the tile's first low-resolution row / column; when the footprint does not fit
    // the box the expand phase reads the map itself, the (unused) copy still keeps the ring's phases in step.
    // A CTA walks a CONTIGUOUS range of tiles (raster order inside an image, images in order): consecutive tiles share halo
    // rows in L2 and, above all, the image -- the NORM coefficients and the statistics flush are then per image change,
    // not per tile (with tiles dealt gridDim.x apart every tile of a small map was a new image: a dependent chain of fp64
    // loads from L2, divisions, two barriers and 128 atomics per tile)
    // The plain and UP modes keep the strided deal (CTAs that run together work on neighbouring tiles: measured 8 % faster on
    // maps beyond L2, where NORM still gains 14 % from the contiguous walk).
    constexpr bool kContig = MODE == kNorm;
    const int per_cta = (ntiles + (int)gridDim.x - 1) / (int)gridDim.x;
    const int t_first = kContig ? (int)blockIdx.x * per_cta : (int)blockIdx.x;
    const int t_end = kContig ? min(t_first + per_cta, ntiles) : ntiles;
    const int t_step = kContig ? 1 : (int)gridDim.x;
    auto issue = [&](int itq) {
        const long long tq = (long long)t_first + (long long)itq * t_step;
        if (tid != 0 || tq >= t_end) return;
        const TilePos tp = tile_pos((int)tq, tiles_w, tiles_h);
        const int slot = itq % kStages;
        const uint32_t bar = smem_u32(&bars[slot]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                // generic-proxy accesses of the slot -> before the async copy
        mbar_expect_tx(bar, (uint32_t)kStageVec * 16u);
        if (MODE == kUp) {
            const LoBox lb = lo_box(tp, H, W, nrm);
            tma_load_4d(stage_s + (uint32_t)(slot * kStageVec) * 16u, &map_x, bar, c0, lb.x0, lb.y0, tp.n);
        } else {
            tma_load_4d(stage_s + (uint32_t)(slot * kStageVec) * 16u, &map_x, bar, c0, tp.ox0 - 1, tp.oy0 - 1, tp.n);
        }
    };

    pdl::trigger();
    // ---- this thread's 8 channels x 9 taps of weights, kept as the bf16 PAIRS they are stored as ({ch 2k, ch 2k+1} per
    // register: 36 registers instead of 72 fp32 values); w is [C][3][3]
    uint32_t wr[9][4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int tp = 0; tp < 9; ++tp) {
            const uint32_t lo = *reinterpret_cast<const unsigned short*>(wgt + (size_t)(c0 + v * 8 + 2 * k) * 9 + tp);
            const uint32_t hi = *reinterpret_cast<const unsigned short*>(wgt + (size_t)(c0 + v * 8 + 2 * k + 1) * 9 + tp);
            wr[tp][k] = lo | (hi << 16);
        }
    pdl::wait();                                              // the weights above are constants; everything below depends on the stream
    int t = t_first;
#pragma unroll
    for (int k = 0; k < kStages - 1; ++k) issue(k);

    // stats_out: every thread keeps running sums of its outputs in ITS row of `red` (shared memory, not registers: the
    // weights and accumulators already fill the register file), flushed to the fp64 workspace when the image changes
    const bool want_sums = MODE == kNorm && nrm.stats_out != nullptr;
    if (want_sums) {
#pragma unroll
        for (int k = 0; k < 16; ++k) red[tid * 17 + k] = 0.0f;
    }
    int sums_n = -1;
    auto flush_sums = [&](int n) {                            // CTA-uniform call
        __syncthreads();
        float a = 0.0f;
        const int vv = tid >> 4, k = tid & 15;                // tid < 128: 8 vectors x 16 values
        if (tid < 128) {
            for (int st = 0; st < kNT / 8; ++st) a += red[(st * 8 + vv) * 17 + k];
            atomicAdd(nrm.stats_out + ((size_t)n * C + c0 + vv * 8 + (k & 7)) * 2 + (k >> 3), (double)a);
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) red[tid * 17 + kk] = 0.0f;
    };

    int coef_n = -1;                                          // image whose scale / shift `coef` holds
    for (int it = 0; t < t_end; t += t_step, ++it) {
        const int cur = it % kStages;
        issue(it + kStages - 1);                              // into the slot the previous iteration finished with
        mbar_wait(smem_u32(&bars[cur]), (uint32_t)(it / kStages) & 1u);
        const TilePos tp = tile_pos(t, tiles_w, tiles_h);
        const int iy0 = tp.oy0 - 1, ix0 = tp.ox0 - 1;
        uint4* tile = MODE == kUp ? hi_tile : stage0 + cur * kTileVec;

        if (MODE == kUp) {
            // ---- expand: the staged tile is the bilinear upsampling of the low-resolution map, four taps per staged vector read
            // from the staged footprint.  o = fma(ly.l1, t1, ly.l0 * t0), t = fma(lx.l1, v1, lx.l0 * v0) in fp32 pairs: the
            // arithmetic of fots_b200_fpn_merge_nhwc_bf16 (instnorm_kernels.cu), one rounding to bf16.
            {
                // per-tile tables: for every tile row / column the two source offsets (uint4 units from `src`) and weights
                __shared__ Tap rowtab[kIH], coltab[kIW];
                const LoBox lb = lo_box(tp, H, W, nrm);
                const uint4* src;                            // generic pointer: the staged footprint or (it did not fit) the map itself
                int pitch_r, pitch_c, yorg, xorg;
                if (lb.fits) { src = stage0 + cur * kLoVec; pitch_r = kLoW * 8; pitch_c = 8; yorg = lb.y0; xorg = lb.x0; }   // the TMA box [kLoH][kLoW][64 ch]
                else { src = x + (size_t)tp.n * nrm.lh * nrm.lw * CV + c0 / 8; pitch_r = nrm.lw * CV; pitch_c = CV; yorg = 0; xorg = 0; }
                if (tid < kIH) {
                    const int iy = iy0 + tid;
                    Tap tq = {-1, -1, 0.f, 0.f};
                    if (iy >= 0 && iy < H) {
                        const Lerp l = lerp_coord(iy, nrm.lh, nrm.sy);
                        tq.o0 = (l.i0 - yorg) * pitch_r; tq.o1 = (l.i1 - yorg) * pitch_r; tq.l0 = l.l0; tq.l1 = l.l1;
                    }
                    rowtab[tid] = tq;
                } else if (tid >= 32 && tid < 32 + kIW) {
                    const int c = tid - 32, ix = ix0 + c;
                    Tap tq = {-1, -1, 0.f, 0.f};
                    if (ix >= 0 && ix < W) {
                        const Lerp l = lerp_coord(ix, nrm.lw, nrm.sx);
                        tq.o0 = (l.i0 - xorg) * pitch_c; tq.o1 = (l.i1 - xorg) * pitch_c; tq.l0 = l.l0; tq.l1 = l.l1;
                    }
                    coltab[c] = tq;
                }
                __syncthreads();
#pragma unroll
                for (int kk = 0; kk < (kTileVec + kNT - 1) / kNT; ++kk) {
                    const int i = tid + kk * kNT;
                    if (i >= kTileVec) break;
                    const int vv = i & 7, p = i >> 3, r = p / kIW, c = p - r * kIW;
                    const Tap tr = rowtab[r], tc = coltab[c];
                    uint4 pk = make_uint4(0, 0, 0, 0);
                    if (tr.o0 >= 0 && tc.o0 >= 0) {
                        const uint4* r0 = src + (tr.o0 + vv);
                        const uint4* r1 = src + (tr.o1 + vv);
                        const uint4 q00 = r0[tc.o0], q01 = r0[tc.o1], q10 = r1[tc.o0], q11 = r1[tc.o1];
                        const uint32_t w00[4] = {q00.x, q00.y, q00.z, q00.w}, w01[4] = {q01.x, q01.y, q01.z, q01.w};
                        const uint32_t w10[4] = {q10.x, q10.y, q10.z, q10.w}, w11[4] = {q11.x, q11.y, q11.z, q11.w};
                        const uint64_t lx0 = pair_f32(tc.l0, tc.l0), lx1 = pair_f32(tc.l1, tc.l1);
                        const uint64_t ly0 = pair_f32(tr.l0, tr.l0), ly1 = pair_f32(tr.l1, tr.l1);
                        uint32_t ow[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t t0 = fma2(lx1, widen2(w01[k]), mul2(lx0, widen2(w00[k])));
                            const uint64_t t1 = fma2(lx1, widen2(w11[k]), mul2(lx0, widen2(w10[k])));
                            float lo_, hi_;
                            split2(fma2(ly1, t1, mul2(ly0, t0)), lo_, hi_);
                            ow[k] = pack2(lo_, hi_);
                        }
                        pk = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                    }
                    tile[i] = pk;
                }
            }
            __syncthreads();
        }
        if (MODE == kNorm) {
            // scale / shift of this CTA's 64 channels for image n, then normalise + activate the staged tile in place
            if (want_sums && sums_n != tp.n) {
                if (sums_n >= 0) flush_sums(sums_n);
                sums_n = tp.n;
            }
            if (nrm.stats == nullptr) goto conv;              // CTA-uniform: output statistics only, nothing to normalise
            if (coef_n != tp.n) {                             // CTA-uniform
                coef_n = tp.n;
            if (tid < kCB) {
                const int c = c0 + tid;
                const double hw = (double)H * (double)W;
                const double m = nrm.stats[((size_t)tp.n * C + c) * 2] / hw;
                double var = nrm.stats[((size_t)tp.n * C + c) * 2 + 1] / hw - m * m;
                var = var < 0.0 ? 0.0 : var;
                const float rstd = rsqrtf((float)var + nrm.eps), mean = (float)m;
                const float g0 = nrm.gamma ? nrm.gamma[c] : 1.0f, b0 = nrm.beta ? nrm.beta[c] : 0.0f;
                coef[tid] = rstd * g0;
                coef[kCB + tid] = b0 - mean * rstd * g0;
            }
            __syncthreads();
            }
            for (int i = tid; i < kTileVec; i += kNT) {
                const int vv = i & 7, p = i >> 3, r = p / kIW, c = p - r * kIW;
                const int iy = iy0 + r, ix = ix0 + c;
                if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;                 // padding: stays exactly zero
                float f[8];
                unpack8(tile[i], f);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float tt = fmaf(f[k], coef[vv * 8 + k], coef[kCB + vv * 8 + k]);
                    f[k] = tt > 0.0f ? tt : tt * nrm.slope;
                }
                uint4 pk;                                                              // one bf16 rounding, as the separate apply pass stores it
                pk.x = pack2(f[0], f[1]); pk.y = pack2(f[2], f[3]); pk.z = pack2(f[4], f[5]); pk.w = pack2(f[6], f[7]);
                tile[i] = pk;
            }
            __syncthreads();
        }

    conv:
        // ---- 3x3 accumulation: 4 outputs x 8 channels per thread, order (r, s) = (0,0) .. (2,2).  fma.rn.f32.bf16 (sm_100
        // FHFMA.BF16) multiplies two bf16 operands -- selected as halves of the packed registers, nothing is widened -- into
        // an fp32 accumulator: exact product, one rounding, i.e. bit-identical to fmaf(float(x), float(w), acc), without the
        // shift / mask per element that kept the ALU pipe at 58 % (profiles/r02_dwconv_up_first_version.txt).
        float acc[4][8];
#pragma unroll
        for (int o = 0; o < 4; ++o)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[o][k] = 0.0f;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const uint4* trow = tile + ((row + r) * kIW + q4 * 4) * 8 + v;
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const uint4 raw = trow[j * 8];
                const uint32_t xw[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    const int sx = j - o;                                              // tap column of output o fed by input column j
                    if (sx >= 0 && sx < 3) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) fma_bf16x2(acc[o][2 * k], acc[o][2 * k + 1], xw[k], wr[r * 3 + sx][k]);
                    }
                }
            }
        }
        const int oy = tp.oy0 + row;
        float ssum[8], ssq[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) ssum[k] = ssq[k] = 0.0f;
        if (oy < H) {
            uint4* out = y + (((size_t)tp.n * H + oy) * W) * CV + c0 / 8 + v;
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const int ox = tp.ox0 + q4 * 4 + o;
                if (ox < W) {
                    const float* a = acc[o];
                    uint4 pk;
                    pk.x = pack2(a[0], a[1]); pk.y = pack2(a[2], a[3]); pk.z = pack2(a[4], a[5]); pk.w = pack2(a[6], a[7]);
                    out[(size_t)ox * CV] = pk;
                    if (want_sums) {                                     // of the ROUNDED values, as a separate pass would see them
                        float f[8];
                        unpack8(pk, f);
#pragma unroll
                        for (int k = 0; k < 8; ++k) { ssum[k] += f[k]; ssq[k] = fmaf(f[k], f[k], ssq[k]); }
                    }
                }
            }
        }
        if (want_sums) {
#pragma unroll
            for (int k = 0; k < 8; ++k) { red[tid * 17 + k] += ssum[k]; red[tid * 17 + 8 + k] += ssq[k]; }
        }
        __syncthreads();                                      // the tile buffer is free: the next iteration prefetches into it
    }
    if (want_sums && sums_n >= 0) flush_sums(sums_n);
}

}  // namespace

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn dw_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}
// bf16 [N, H, W, C] with a box of 64 channels x bw x bh pixels of one image, dense in shared memory ([bh][bw][64]); rows /
// columns outside the image arrive as zeros
static bool dw_make_map(CUtensorMap* m, const void* ptr, int N, int H, int W, int C, int bw, int bh) {
    if (!dw_encode_fn()) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    const cuuint32_t box[4] = {(cuuint32_t)kCB, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    return dw_encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Persistent launch: as many CTAs per channel block as the device holds at once (whole waves; see resident sizing in
// instnorm_kernels.cu), never more than there are tiles.  The dynamic shared-memory opt-in is set on every launch (it is
// per device, and cheap).
template <int MODE>
static cudaError_t dw_pipe_launch(size_t smem, const CUtensorMap& map, const uint4* xp, const __nv_bfloat16* wp, uint4* yp, int H, int W, int C,
                                  int tiles_w, int tiles_h, int ntiles, int cblocks, const DwNorm& nrm, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(dwconv3x3_s1_pipe_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 0, occ = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dwconv3x3_s1_pipe_kernel<MODE>, kNT, smem)) != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    int gx = sms * occ / cblocks;
    if (gx < 1) gx = 1;
    if (gx > ntiles) gx = ntiles;
    return pdl::launch(dwconv3x3_s1_pipe_kernel<MODE>, dim3((unsigned)gx, (unsigned)cblocks), dim3(kNT), smem, stream, map, xp, wp, yp, H, W, C,
                       tiles_w, tiles_h, ntiles, nrm);
}

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
        const cudaError_t em = pdl::zero_f64(none.stats_out, (size_t)N * C * 2, stream);
        if (em != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
    }
    const bool norm_on = nrm != nullptr && nrm->stats != nullptr;
    const bool up = none.lh > 0;
    if (up && (stride != 1 || norm_on)) return RROI_B200_ERR_INVALID_ARG;     // upsample-on-load: stride 1, no normalisation
    if (stride == 1) {
        const int tiles_h = (Ho + kTH - 1) / kTH;
        const long long ntiles = (long long)tiles_w * tiles_h * N;
        if (ntiles > (1LL << 30) || (long long)H * W * (C / 8) >= (1LL << 31)) return RROI_B200_ERR_INVALID_ARG;   // 32-bit offsets inside a plane
        const int cblocks = C / kCB;
        cudaError_t e1 = cudaSuccess;
        CUtensorMap map;
        if (up ? !dw_make_map(&map, x, N, none.lh, none.lw, C, kLoW, kLoH) : !dw_make_map(&map, x, N, H, W, C, kIW, kIH)) return RROI_B200_ERR_CUDA;
        if (up) e1 = dw_pipe_launch<kUp>((size_t)(kStages * kLoVec + kTileVec) * 16, map, xp, wp, yp, H, W, C, tiles_w, tiles_h, (int)ntiles, cblocks, none, stream);
        else if (norm_on || none.stats_out) e1 = dw_pipe_launch<kNorm>((size_t)kStages * kTileVec * 16 + 2 * kCB * 4 + (none.stats_out ? kNT * 17 * 4 : 0), map, xp, wp, yp, H, W, C, tiles_w, tiles_h, (int)ntiles, cblocks, none, stream);
        else e1 = dw_pipe_launch<kPlain>((size_t)kStages * kTileVec * 16, map, xp, wp, yp, H, W, C, tiles_w, tiles_h, (int)ntiles, cblocks, none, stream);
        if (e1 != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
        return RROI_B200_OK;
    }
    {
        constexpr int TH = 4;
        const dim3 grid((unsigned)(tiles_w * ((Ho + TH - 1) / TH)), (unsigned)(C / kCB), (unsigned)N);
        if (norm_on) (void)pdl::launch(dwconv3x3_kernel<2, TH, true>, dim3(grid), dim3(TH * 32), 0, stream, xp, wp, yp, H, W, C, Ho, Wo, tiles_w, none);
        else (void)pdl::launch(dwconv3x3_kernel<2, TH, false>, dim3(grid), dim3(TH * 32), 0, stream, xp, wp, yp, H, W, C, Ho, Wo, tiles_w, none);
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
