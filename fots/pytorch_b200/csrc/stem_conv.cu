// stem_conv.cu -- the first convolution of the shared feature extractor (tools/models.py:250-251: Conv2d(3, 16, 3,
// stride 1, pad 1, bias=False) on the full-resolution image) fused with the statistics pass of the CReLU_IN that
// consumes it (tools/models.py:41-48).
//
// Why a dedicated kernel: this layer is pure bandwidth -- 27 MACs per output value, 88 MB of fp32 image in and 236 MB of
// bf16 activations out per 8 images, ~50 us of HBM time -- but with 3 input channels no library implicit-GEMM tile
// fits it (cuDNN: 326 us, plus a 45 us statistics pass over its output).  Here the image tile (+1 pixel halo) is read
// once with coalesced loads, rounded to bf16 (what autocast feeds the library convolution) and staged in shared
// memory as [row][pixel][4 channels] (channel 3 = 0); each warp then computes 16 pixels x 16 output channels with six
// mma.sync.m16n8k16 (one k-step per filter row: k = s*4 + ch, 12 of 16 used), so the arithmetic is off the critical
// path and every A fragment is an aligned, conflict-free 32-bit shared load.  fp32 accumulation, one rounding to bf16,
// and the per-image per-channel sum / sum of squares of the ROUNDED values (what a statistics pass over y would see)
// accumulated in registers across the CTA's tiles and flushed once per image.
// tcgen05 is not used on purpose: K = 27 cannot fill a UMMA tile and the tensor pipe is idle either way.
#include "../../../include/fots_b200_pipeline.h"
#include "pdl.cuh"
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int TH = 8, TW = 128;                 // output pixels per CTA tile
constexpr int kThreads = 256;                   // 8 warps: warp w owns tile row w
constexpr int HALO_W = TW + 2, HALO_H = TH + 2;
constexpr int PITCH = HALO_W * 4;               // bf16 elements per staged row

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// Input element: fp32 (the reference's preprocessed image, test.py:80-83) or the raw uint8 pixel, normalised here as
// the reference does on the host (x / 128 - 1: exact in fp32, so both inputs give the same bf16 operand).
__device__ __forceinline__ float4 load4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 load4(const uint8_t* p) {
    const uchar4 u = __ldg(reinterpret_cast<const uchar4*>(p));
    return make_float4(fmaf((float)u.x, 0.0078125f, -1.0f), fmaf((float)u.y, 0.0078125f, -1.0f),
                       fmaf((float)u.z, 0.0078125f, -1.0f), fmaf((float)u.w, 0.0078125f, -1.0f));
}
__device__ __forceinline__ float load1(const float* p) { return __ldg(p); }
__device__ __forceinline__ float load1(const uint8_t* p) { return fmaf((float)__ldg(p), 0.0078125f, -1.0f); }

template <typename In>
__global__ void __launch_bounds__(kThreads, 3)
stem_conv_kernel(const In* __restrict__ x, const __nv_bfloat16* __restrict__ wgt, uint32_t* __restrict__ y,
                 double* __restrict__ stats, int B, int H, int W, int tiles_x, int tiles_y, int tiles_per_cta) {
    __shared__ __align__(16) __nv_bfloat16 tile[HALO_H * PITCH];
    __shared__ float red[32];                   // [16 channels][sum, sumsq] of one flush
    pdl::trigger();                             // weights are constants: fragments are built before pdl::wait()
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;

    // B fragments (weights), one k-step per filter row r: B[k][n] = w[n][r][s][ch], k = s*4 + ch (ch < 3, s < 3)
    uint32_t bw[3][2][2];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int nh = 0; nh < 2; ++nh)
#pragma unroll
            for (int kh = 0; kh < 2; ++kh) {
                const int n = nh * 8 + g;
                uint32_t v = 0;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int k = kh * 8 + 2 * t + e, s = k >> 2, ch = k & 3;
                    uint16_t bits = 0;
                    if (s < 3 && ch < 3) bits = *reinterpret_cast<const uint16_t*>(wgt + ((n * 3 + r) * 3 + s) * 3 + ch);
                    v |= (uint32_t)bits << (16 * e);
                }
                bw[r][nh][kh] = v;
            }
    for (int i = threadIdx.x; i < HALO_H * PITCH; i += kThreads) tile[i] = __float2bfloat16(0.0f);   // channel 3 stays 0
    if (threadIdx.x < 32) red[threadIdx.x] = 0.0f;
    pdl::wait();

    float ssum[4] = {0.f, 0.f, 0.f, 0.f}, ssq[4] = {0.f, 0.f, 0.f, 0.f};   // this thread's 4 output channels
    int cur_b = -1;
    const int tiles_img = tiles_x * tiles_y, total = tiles_img * B;
    const int first = blockIdx.x * tiles_per_cta;
    const int last = min(total, first + tiles_per_cta);

    auto flush = [&]() {            // called by all threads of the CTA at the same points
        if (cur_b >= 0 && stats != nullptr) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float a = ssum[i], q = ssq[i];
#pragma unroll
                for (int m = 4; m < 32; m <<= 1) {
                    a += __shfl_xor_sync(0xffffffffu, a, m);
                    q += __shfl_xor_sync(0xffffffffu, q, m);
                }
                if (g == 0) {
                    const int c = (i >> 1) * 8 + 2 * t + (i & 1);
                    atomicAdd(&red[2 * c], a);
                    atomicAdd(&red[2 * c + 1], q);
                }
                ssum[i] = ssq[i] = 0.f;
            }
            __syncthreads();
            if (threadIdx.x < 32) {
                atomicAdd(stats + (size_t)cur_b * 32 + threadIdx.x, (double)red[threadIdx.x]);
                red[threadIdx.x] = 0.0f;
            }
            __syncthreads();
        }
    };

    // Staging, fast path (W % 4 == 0 and a 16-byte aligned image): a halo row is covered by 99 ALIGNED float4 vectors
    // starting one vector before pixel x0 (vector v holds halo elements 4v-1 .. 4v+2); because rows are a multiple of
    // four floats long, a vector is either entirely inside its image row or entirely outside.  The vectors of the
    // NEXT tile are fetched into registers before the current tile is computed, so the global latency hides behind
    // the MMAs and the epilogue.
    constexpr int VEC_ROW = (HALO_W * 3 + 1 + 3) / 4 + 1;          // 99
    constexpr int VEC_IT = (HALO_H * VEC_ROW + kThreads - 1) / kThreads;   // 4
    const bool fast = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & (4 * sizeof(In) - 1)) == 0);
    float4 pre[VEC_IT];
    auto tile_coords = [&](int tile_id, int& b, int& y0, int& x0) {
        b = tile_id / tiles_img;
        const int rem = tile_id - b * tiles_img;
        const int ty = rem / tiles_x;
        y0 = ty * TH; x0 = (rem - ty * tiles_x) * TW;
    };
    auto prefetch = [&](int tile_id) {
        int b, y0, x0;
        tile_coords(tile_id, b, y0, x0);
        const In* img = x + (size_t)b * H * W * 3;
#pragma unroll
        for (int it = 0; it < VEC_IT; ++it) {
            const int i = it * kThreads + (int)threadIdx.x;
            const int row = i / VEC_ROW, v = i - row * VEC_ROW;
            const int yy = y0 - 1 + row;
            const long long f = (long long)x0 * 3 - 4 + 4 * v;          // first float of the vector inside its row
            pre[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < HALO_H && yy >= 0 && yy < H && f >= 0 && f + 3 < (long long)W * 3)
                pre[it] = load4(img + (size_t)yy * W * 3 + f);
        }
    };
    auto commit = [&]() {                                            // registers -> bf16 shared tile
#pragma unroll
        for (int it = 0; it < VEC_IT; ++it) {
            const int i = it * kThreads + (int)threadIdx.x;
            const int row = i / VEC_ROW, v = i - row * VEC_ROW;
            if (row >= HALO_H) continue;
            const float vals[4] = {pre[it].x, pre[it].y, pre[it].z, pre[it].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int e = 4 * v - 1 + j;
                if (e >= 0 && e < HALO_W * 3) {
                    const int px = e / 3, ch = e - px * 3;
                    tile[row * PITCH + px * 4 + ch] = __float2bfloat16(vals[j]);
                }
            }
        }
    };

    if (fast && first < last) prefetch(first);
    for (int tile_id = first; tile_id < last; ++tile_id) {
        int b, y0, x0;
        tile_coords(tile_id, b, y0, x0);
        if (b != cur_b) { flush(); cur_b = b; }
        __syncthreads();                         // previous tile's readers are done
        if (fast) {
            commit();
        } else {
            // generic path: element by element, zero outside the image
            const In* img = x + (size_t)b * H * W * 3;
            for (int i = threadIdx.x; i < HALO_H * HALO_W * 3; i += kThreads) {
                const int row = i / (HALO_W * 3), e = i - row * (HALO_W * 3);
                const int px = e / 3, ch = e - px * 3;
                const int yy = y0 - 1 + row, xx = x0 - 1 + px;
                float v = 0.0f;
                if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = load1(img + ((size_t)yy * W + xx) * 3 + ch);
                tile[row * PITCH + px * 4 + ch] = __float2bfloat16(v);
            }
        }
        __syncthreads();
        if (fast && tile_id + 1 < last) prefetch(tile_id + 1);
        // ---- warp = tile row; 8 sub-tiles of 16 pixels ----
        const int yo = y0 + warp;
#pragma unroll 2
        for (int sub = 0; sub < TW / 16; ++sub) {
            float acc[2][4];
#pragma unroll
            for (int nh = 0; nh < 2; ++nh) acc[nh][0] = acc[nh][1] = acc[nh][2] = acc[nh][3] = 0.f;
            const int xl = sub * 16;             // halo column of the left neighbour of pixel 0 of the sub-tile
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                // A[m][k] = in[y + r - 1][x_m + s - 1][ch], k = s*4 + ch  ->  word (s, ch pair) of pixel column xl + m + s
                const uint32_t* rowp = reinterpret_cast<const uint32_t*>(tile + (warp + r) * PITCH);
                uint32_t a[4];
                const int w0 = (xl + g) * 2 + t;             // word index: pixel*2 + (k pair within pixel); k = 2t -> s = t>>1
                a[0] = rowp[w0];                             // (m = g,     k = 2t, 2t+1)
                a[1] = rowp[w0 + 16];                        // (m = g + 8, k = 2t, 2t+1)
                a[2] = t < 2 ? rowp[w0 + 4] : 0u;            // (m = g,     k = 8 + 2t, ..): s = 2 only for t < 2
                a[3] = t < 2 ? rowp[w0 + 20] : 0u;           // (m = g + 8, k = 8 + 2t, ..)
                mma_bf16_16816(acc[0], a, bw[r][0][0], bw[r][0][1]);
                mma_bf16_16816(acc[1], a, bw[r][1][0], bw[r][1][1]);
            }
            // ---- epilogue: round to bf16, store, statistics of the rounded values ----
            const int xa = x0 + xl + g, xb = xa + 8;
            const bool row_ok = yo < H;
            uint32_t m[4];                                               // pixel g: nh 0, nh 1; pixel g + 8: nh 0, nh 1
#pragma unroll
            for (int nh = 0; nh < 2; ++nh) {
                const uint32_t p0 = pack2(acc[nh][0], acc[nh][1]);      // pixel g,     channels nh*8 + 2t, +1
                const uint32_t p1 = pack2(acc[nh][2], acc[nh][3]);      // pixel g + 8
                m[nh] = p0; m[2 + nh] = p1;
                if (row_ok && xa < W) {
                    const float v0 = __uint_as_float(p0 << 16), v1 = __uint_as_float(p0 & 0xffff0000u);
                    ssum[2 * nh] += v0; ssq[2 * nh] = fmaf(v0, v0, ssq[2 * nh]);
                    ssum[2 * nh + 1] += v1; ssq[2 * nh + 1] = fmaf(v1, v1, ssq[2 * nh + 1]);
                }
                if (row_ok && xb < W) {
                    const float v0 = __uint_as_float(p1 << 16), v1 = __uint_as_float(p1 & 0xffff0000u);
                    ssum[2 * nh] += v0; ssq[2 * nh] = fmaf(v0, v0, ssq[2 * nh]);
                    ssum[2 * nh + 1] += v1; ssq[2 * nh + 1] = fmaf(v1, v1, ssq[2 * nh + 1]);
                }
            }
            // 4x4 transpose inside the quad (lanes t = 0..3 of one g): lane t ends up with the four words of
            // (pixel, channel half) number t, i.e. 16 contiguous bytes -> one 128-bit store, full 32-byte sectors per pixel
            {
                const bool odd = t & 1;
                const uint32_t s0 = odd ? m[0] : m[1], s1 = odd ? m[2] : m[3];
                const uint32_t r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
                // even lane: {m0, r0 | m2, r1} = rows (t, t+1) of columns 0 and 2; odd lane: {r0, m1 | r1, m3} = columns 1 and 3
                const uint32_t a0 = odd ? r0 : m[0], a1 = odd ? m[1] : r0, b0 = odd ? r1 : m[2], b1 = odd ? m[3] : r1;
                const bool hi = t & 2;
                const uint32_t u0 = hi ? a0 : b0, u1 = hi ? a1 : b1;
                const uint32_t q0 = __shfl_xor_sync(0xffffffffu, u0, 2), q1 = __shfl_xor_sync(0xffffffffu, u1, 2);
                uint4 o;
                if (!hi) { o.x = a0; o.y = a1; o.z = q0; o.w = q1; }     // rows 0,1 own + rows 2,3 received
                else     { o.x = q0; o.y = q1; o.z = b0; o.w = b1; }
                const int xo = (t < 2) ? xa : xb;
                if (row_ok && xo < W)
                    *reinterpret_cast<uint4*>(y + (((size_t)b * H + yo) * W + xo) * 8 + (t & 1) * 4) = o;
            }
        }
    }
    flush();
}

}  // namespace

template <typename In>
static int stem_launch(const In* x, const void* w, void* y, double* stats, int B, int H, int W, cudaStream_t stream) {
    if (!x || !w || !y || B <= 0 || H <= 0 || W <= 0) return RROI_B200_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(y) & 15) || (reinterpret_cast<uintptr_t>(x) & (sizeof(In) - 1))) return RROI_B200_ERR_INVALID_ARG;
    const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH;
    const long long total = (long long)tiles_x * tiles_y * B;
    if (total > 0x7fffffffLL) return RROI_B200_ERR_TOO_LARGE;
    if (stats) {
        const cudaError_t e = pdl::zero_f64(stats, (size_t)B * 32, stream);
        if (e != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
    }
    // persistent: contiguous tile ranges, so a CTA stays inside one image and flushes its sums once or twice
    // two whole waves of resident CTAs (the tile ranges are equal, so a ragged last wave is pure loss)
    int dev = 0, sms = 148, occ = 3;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, stem_conv_kernel<In>, kThreads, 0) != cudaSuccess || sms <= 0 || occ <= 0) {
        (void)cudaGetLastError();
        sms = 148; occ = 3;
    }
    const long long ctas_wanted = 2LL * sms * occ;
    int per = (int)((total + ctas_wanted - 1) / ctas_wanted);
    int grid = (int)((total + per - 1) / per);
    if (grid > sms * occ && grid < ctas_wanted * 15 / 16) {    // second wave under ~88 % full: one tile more per CTA -> a single fuller wave pair
        const int per1 = (int)((total + (long long)sms * occ - 1) / ((long long)sms * occ));
        if ((long long)per1 * sms * occ - total < total / 16) { per = per1; grid = (int)((total + per - 1) / per); }
    }
    (void)pdl::launch(stem_conv_kernel<In>, dim3(grid), dim3(kThreads), 0, stream, x, static_cast<const __nv_bfloat16*>(w), static_cast<uint32_t*>(y), stats,
                                                        B, H, W, tiles_x, tiles_y, per);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { (void)cudaGetLastError(); return RROI_B200_ERR_CUDA; }
    return RROI_B200_OK;
}

extern "C" int fots_b200_stem_conv3x3_c3_c16(const float* x, const void* w, void* y, double* stats, int B, int H, int W,
                                             cudaStream_t stream) {
    return stem_launch<float>(x, w, y, stats, B, H, W, stream);
}

extern "C" int fots_b200_stem_conv3x3_c3_c16_u8(const unsigned char* x, const void* w, void* y, double* stats, int B, int H, int W,
                                                cudaStream_t stream) {
    return stem_launch<uint8_t>(x, w, y, stats, B, H, W, stream);
}
