// lstm_kernels.cu -- the recurrent half of Consumer B (CRNN, /root/reference/tools/models.py:17-33 BidirectionalLSTM,
// :898-909 CRNN.forward): two bidirectional LSTM layers, each followed by a Linear "embedding".  The reference runs
// them through cuDNN (nn.LSTM) in fp32; at the CRNN's sizes (T = 65 steps, 64 sequences, hidden 256) that is 2.2 of the
// 2.85 ms of the whole recogniser, because the recurrence is 65 dependent steps of a tiny matrix product.
//
// Here a layer is three launches:
//   1. fots_b200_gemm_bf16w   G[T*N, 8H] = X[T*N, nin] * W_ih^T + (b_ih + b_hh), both directions at once (no recurrence in
//      it, so it is one dense GEMM over all time steps);
//   2. fots_b200_bilstm_recurrent   ONE persistent kernel for the whole time loop of both directions: a cluster of 8
//      CTAs per direction, CTA r owning hidden units [32r, 32r+32) -- its 128 x 256 slice of W_hh lives in shared memory
//      for all T steps, the cell state in registers; per step every CTA publishes its 64 x 32 slice of h_t in shared
//      memory, one barrier.cluster later all eight CTAs pull the full h_t through distributed shared memory (16-byte
//      ld.shared::cluster), and the gates come out of mma.sync with fp32 accumulation;
//   3. fots_b200_gemm_bf16w   the embedding Linear on [h_fwd, h_bwd].
// Arithmetic: weights in bf16 (as the rest of the B200 inference path stores them), accumulation, gates, cell and hidden
// state in fp32; fp32 activations that feed a tensor-core product (h_{t-1}, the LSTM outputs) are split into bf16
// hi + lo parts and multiplied in two MMAs, so the result matches an fp32 nn.LSTM with the same (bf16-rounded) weights to
// ~1e-5 -- no bf16 rounding of the recurrent state accumulates over the 65 steps.
// mma.sync (not tcgen05) on purpose: M = 64 sequences per step cannot fill a 128-row UMMA tile, and the step is bound
// by the barrier + state exchange, not by the tensor pipe.
#include "../../../include/fots_b200_pipeline.h"
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// x = hi + lo with hi = bf16(x), lo = bf16(x - hi): 16 mantissa bits through two bf16 products
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
    return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// ------------------------------------------------------------------------------------------------ GEMM
// C[M, N] (fp32) = A[M, K] * W[N, K]^T + bias[N].  A: bf16, or fp32 (split into hi + lo, two MMAs per product).
// W: bf16 [N, K] (an nn.Linear / nn.LSTM weight as stored).  K % 32 == 0.  CTA tile 128 x 128, k-step 32, 8 warps
// (4 x 2: 32 x 64 per warp), operands staged through registers into padded shared memory (conflict-free 32-bit
// fragment loads), next k-step's global loads in flight during the MMAs.
constexpr int GB = 128, GK = 32, GP = GK + 8;          // tile edge, k-step, padded row pitch (bf16 elements)

template <bool A_F32>
__global__ void __launch_bounds__(256) gemm_bf16w_kernel(const void* __restrict__ Aptr, const __nv_bfloat16* __restrict__ W,
                                                         const float* __restrict__ bias, float* __restrict__ C, int M, int N, int K) {
    __shared__ __align__(16) __nv_bfloat16 As[(A_F32 ? 2 : 1) * GB * GP];      // hi [, lo]
    __shared__ __align__(16) __nv_bfloat16 Ws[GB * GP];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int m0 = blockIdx.y * GB, n0 = blockIdx.x * GB;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 64;
    float acc[2][8][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;

    // staging: thread -> (row = tid / 2, 16 consecutive k = (tid & 1) * 16) of both tiles
    const int srow = threadIdx.x >> 1, sk = (threadIdx.x & 1) * 16;
    const bool a_ok = m0 + srow < M, w_ok = n0 + srow < N;
    float4 a_f32[4];
    uint4 a_b16[2], w_b16[2];
    auto fetch = [&](int k0) {
        if (A_F32) {
            const float* a = static_cast<const float*>(Aptr) + (size_t)(m0 + srow) * K + k0 + sk;
#pragma unroll
            for (int i = 0; i < 4; ++i) a_f32[i] = a_ok ? __ldg(reinterpret_cast<const float4*>(a) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            const __nv_bfloat16* a = static_cast<const __nv_bfloat16*>(Aptr) + (size_t)(m0 + srow) * K + k0 + sk;
#pragma unroll
            for (int i = 0; i < 2; ++i) a_b16[i] = a_ok ? __ldg(reinterpret_cast<const uint4*>(a) + i) : make_uint4(0, 0, 0, 0);
        }
        const __nv_bfloat16* w = W + (size_t)(n0 + srow) * K + k0 + sk;
#pragma unroll
        for (int i = 0; i < 2; ++i) w_b16[i] = w_ok ? __ldg(reinterpret_cast<const uint4*>(w) + i) : make_uint4(0, 0, 0, 0);
    };
    auto commit = [&]() {
        if (A_F32) {
            const float* f = reinterpret_cast<const float*>(a_f32);
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                __nv_bfloat16 h0, l0, h1, l1;
                split_bf16(f[2 * i], h0, l0);
                split_bf16(f[2 * i + 1], h1, l1);
                hi[i] = pack_bf16x2(h0, h1); lo[i] = pack_bf16x2(l0, l1);
            }
            uint4* dh = reinterpret_cast<uint4*>(As + srow * GP + sk);
            uint4* dl = reinterpret_cast<uint4*>(As + GB * GP + srow * GP + sk);
            dh[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]); dh[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
            dl[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]); dl[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
        } else {
            uint4* d = reinterpret_cast<uint4*>(As + srow * GP + sk);
            d[0] = a_b16[0]; d[1] = a_b16[1];
        }
        uint4* dw = reinterpret_cast<uint4*>(Ws + srow * GP + sk);
        dw[0] = w_b16[0]; dw[1] = w_b16[1];
    };

    fetch(0);
    for (int k0 = 0; k0 < K; k0 += GK) {
        __syncthreads();                                   // the previous k-step's fragment reads are done
        commit();
        __syncthreads();
        if (k0 + GK < K) fetch(k0 + GK);
#pragma unroll
        for (int kk = 0; kk < GK; kk += 16) {
            uint32_t bf[8][2];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const __nv_bfloat16* wp = Ws + (wn + j * 8 + g) * GP + kk + 2 * t;
                bf[j][0] = *reinterpret_cast<const uint32_t*>(wp);
                bf[j][1] = *reinterpret_cast<const uint32_t*>(wp + 8);
            }
#pragma unroll
            for (int part = 0; part < (A_F32 ? 2 : 1); ++part)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const __nv_bfloat16* ap = As + part * GB * GP + (wm + i * 16 + g) * GP + kk + 2 * t;
                    uint32_t af[4];
                    af[0] = *reinterpret_cast<const uint32_t*>(ap);
                    af[1] = *reinterpret_cast<const uint32_t*>(ap + 8 * GP);
                    af[2] = *reinterpret_cast<const uint32_t*>(ap + 8);
                    af[3] = *reinterpret_cast<const uint32_t*>(ap + 8 * GP + 8);
#pragma unroll
                    for (int j = 0; j < 8; ++j) mma_16816(acc[i][j], af, bf[j][0], bf[j][1]);
                }
        }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = n0 + wn + j * 8 + 2 * t;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int row = m0 + wm + i * 16 + g + half * 8;
                if (row >= M) continue;
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    if (col + e < N) C[(size_t)row * N + col + e] = acc[i][j][2 * half + e] + (bias ? __ldg(bias + col + e) : 0.f);
            }
        }
}

// ------------------------------------------------------------------------------------- recurrent kernel
constexpr int LH = 256;                 // hidden size (CRNN: nh = 256)
constexpr int LCL = 8;                  // CTAs per cluster = hidden-unit slices per direction
constexpr int LU = LH / LCL;            // 32 hidden units per CTA
constexpr int LG = 4 * LU;              // 128 gate columns per CTA
constexpr int LNB = 64;                 // sequences per cluster
constexpr int LP = LH + 8;              // padded pitch of the bf16 operand rows (conflict-free fragment loads)
constexpr size_t kLstmSmem = (size_t)LG * LP * 2 + 2 * (size_t)LNB * LP * 2 + 2 * (size_t)LNB * LU * 4;

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t map_to_rank(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ float4 ld_cluster_v4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }
// tanh(x) = 2 sigmoid(2x) - 1 on the fast exponential (abs. error ~1e-6): libm's tanhf is ~5x the instructions, and the gate
// phase sits on the critical path of all 65 steps
__device__ __forceinline__ float tanh_f(float x) { return 2.0f / (1.0f + __expf(-2.0f * x)) - 1.0f; }

// G    fp32 [T, N, 2, 4H]   input projections + both biases, gate order i, f, g, o (nn.LSTM), direction-major
// Whh  bf16 [2, 4H, H]      weight_hh_l0, weight_hh_l0_reverse
// Y    fp32 [T, N, 2H]      h_t of the forward direction in [:, :, :H], of the reverse direction in [:, :, H:]
// grid = (2 * LCL, ceil(N / 64)), clusters of LCL along x: cluster = (direction, block of 64 sequences).
__global__ void __launch_bounds__(256, 1) bilstm_recurrent_kernel(const float* __restrict__ G, const __nv_bfloat16* __restrict__ Whh,
                                                                   float* __restrict__ Y, int T, int N) {
    extern __shared__ __align__(16) uint8_t smem[];
    __nv_bfloat16* const Wsm = reinterpret_cast<__nv_bfloat16*>(smem);                    // [LG][LP]
    __nv_bfloat16* const Hhi = Wsm + LG * LP;                                             // [LNB][LP]
    __nv_bfloat16* const Hlo = Hhi + LNB * LP;                                            // [LNB][LP]
    float* const outbox = reinterpret_cast<float*>(Hlo + LNB * LP);                       // [2][LNB][LU] fp32, double-buffered
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
    const uint32_t r = cluster_rank();
    const int dir = blockIdx.x / LCL;
    const int nb0 = blockIdx.y * LNB;
    const int wm = (warp >> 1) * 16, wn = (warp & 1) * 64;        // warp tile: 16 sequences x 64 gate columns (2 groups of 8 units)

    // W_hh slice -> shared memory.  Local column c = u8 * 32 + gate * 8 + uu  <->  row gate * H + 32 r + 8 u8 + uu of W_hh:
    // the four n8 tiles of one group of 8 units are its i, f, g, o gates, so a thread finds all four gates of "its" unit at
    // the same fragment position of four consecutive accumulator tiles.
    for (int i = threadIdx.x; i < LG * (LH / 8); i += 256) {
        const int c = i / (LH / 8), k8 = i - c * (LH / 8);
        const int u8 = c >> 5, gate = (c >> 3) & 3, uu = c & 7;
        const size_t row = (size_t)dir * 4 * LH + (size_t)gate * LH + r * LU + u8 * 8 + uu;
        *reinterpret_cast<uint4*>(Wsm + c * LP + k8 * 8) = __ldg(reinterpret_cast<const uint4*>(Whh + row * LH) + k8);
    }
    float cst[2][4];                                               // cell state: [unit group][row half x unit pair]
#pragma unroll
    for (int u = 0; u < 2; ++u) cst[u][0] = cst[u][1] = cst[u][2] = cst[u][3] = 0.f;
    const uint32_t outbox_s = (uint32_t)__cvta_generic_to_shared(outbox);
    __syncthreads();
    cluster_arrive();                                              // every CTA of the cluster has started (DSMEM is live)
    cluster_wait();

    for (int step = 0; step < T; ++step) {
        const int tt = dir == 0 ? step : T - 1 - step;
        // ---- accumulators start from the input projection of this step (independent of the recurrence: issued first)
        float acc[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int u8 = (wn >> 5) + (j >> 2), gate = j & 3;
            const int col = dir * 4 * LH + gate * LH + (int)r * LU + u8 * 8 + 2 * t4;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int n = nb0 + wm + g + half * 8;
                float2 v = make_float2(0.f, 0.f);
                if (n < N) v = __ldg(reinterpret_cast<const float2*>(G + ((size_t)tt * N + n) * (8 * LH) + col));
                acc[j][2 * half] = v.x; acc[j][2 * half + 1] = v.y;
            }
        }
        if (step > 0) {
            cluster_wait();                                        // every CTA has published its slice of h_{t-1}
            // ---- pull h_{t-1}: eight 64 x 32 fp32 slices through distributed shared memory -> local bf16 hi / lo operand.
            // All 16 remote loads of a thread are issued before the first use (their latencies overlap).
            const int par = (step - 1) & 1;
            constexpr int kPull = (LCL * LNB * LU / 4) / 256;      // 16 float4 per thread
            float4 v[kPull];
#pragma unroll
            for (int i = 0; i < kPull; ++i) {
                const int idx = i * 256 + threadIdx.x;
                const int src = idx / (LNB * LU / 4), rem = idx - src * (LNB * LU / 4);
                v[i] = ld_cluster_v4(map_to_rank(outbox_s + (uint32_t)((par * LNB * LU + rem * 4) * 4), (uint32_t)src));
            }
#pragma unroll
            for (int i = 0; i < kPull; ++i) {
                const int idx = i * 256 + threadIdx.x;
                const int src = idx / (LNB * LU / 4), rem = idx - src * (LNB * LU / 4);
                const int row = rem / (LU / 4), q = rem - row * (LU / 4);
                __nv_bfloat16 h[4], l[4];
                split_bf16(v[i].x, h[0], l[0]); split_bf16(v[i].y, h[1], l[1]); split_bf16(v[i].z, h[2], l[2]); split_bf16(v[i].w, h[3], l[3]);
                const int k = src * LU + q * 4;
                *reinterpret_cast<uint2*>(Hhi + row * LP + k) = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
                *reinterpret_cast<uint2*>(Hlo + row * LP + k) = make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
            }
            __syncthreads();
            // ---- gates += h_{t-1} * W_hh^T  (hi and lo parts)
#pragma unroll 4
            for (int kk = 0; kk < LH; kk += 16) {
                uint32_t ah[4], al[4];
                const __nv_bfloat16* ap = Hhi + (wm + g) * LP + kk + 2 * t4;
                ah[0] = *reinterpret_cast<const uint32_t*>(ap);           ah[1] = *reinterpret_cast<const uint32_t*>(ap + 8 * LP);
                ah[2] = *reinterpret_cast<const uint32_t*>(ap + 8);       ah[3] = *reinterpret_cast<const uint32_t*>(ap + 8 * LP + 8);
                const __nv_bfloat16* lp = Hlo + (wm + g) * LP + kk + 2 * t4;
                al[0] = *reinterpret_cast<const uint32_t*>(lp);           al[1] = *reinterpret_cast<const uint32_t*>(lp + 8 * LP);
                al[2] = *reinterpret_cast<const uint32_t*>(lp + 8);       al[3] = *reinterpret_cast<const uint32_t*>(lp + 8 * LP + 8);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const __nv_bfloat16* wp = Wsm + (wn + j * 8 + g) * LP + kk + 2 * t4;
                    const uint32_t b0 = *reinterpret_cast<const uint32_t*>(wp), b1 = *reinterpret_cast<const uint32_t*>(wp + 8);
                    mma_16816(acc[j], ah, b0, b1);
                    mma_16816(acc[j], al, b0, b1);
                }
            }
        }
        // ---- gates -> cell -> hidden; publish this CTA's slice of h_t
        const int par = step & 1;
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float ig = sigmoid_f(acc[4 * u + 0][e]), fg = sigmoid_f(acc[4 * u + 1][e]);
                const float gg = tanh_f(acc[4 * u + 2][e]), og = sigmoid_f(acc[4 * u + 3][e]);
                const float c = fg * cst[u][e] + ig * gg;
                cst[u][e] = c;
                const float h = og * tanh_f(c);
                const int row = wm + g + (e >> 1) * 8;
                const int unit = ((wn >> 5) + u) * 8 + 2 * t4 + (e & 1);               // local hidden unit 0..31
                outbox[(par * LNB + row) * LU + unit] = h;
                const int n = nb0 + row;
                if (n < N) Y[((size_t)tt * N + n) * (2 * LH) + dir * LH + (int)r * LU + unit] = h;
            }
        cluster_arrive();                                          // release: the slice is visible to the peers after their wait
    }
    cluster_wait();                                                // nobody may leave while a peer still reads its shared memory
}

int status_of(cudaError_t e) {
    if (e == cudaSuccess) return RROI_B200_OK;
    (void)cudaGetLastError();
    return RROI_B200_ERR_CUDA;
}

}  // namespace

extern "C" int fots_b200_gemm_bf16w(const void* A, int a_is_f32, const void* W, const float* bias, float* C, int M, int N, int K,
                                    cudaStream_t stream) {
    if (M < 0 || N <= 0 || K <= 0 || K % GK != 0 || (M > 0 && (!A || !W || !C))) return RROI_B200_ERR_INVALID_ARG;
    if (((uintptr_t)A | (uintptr_t)W) & 15) return RROI_B200_ERR_INVALID_ARG;
    if (M == 0) return RROI_B200_OK;
    const dim3 grid((N + GB - 1) / GB, (M + GB - 1) / GB);
    if (a_is_f32) gemm_bf16w_kernel<true><<<grid, 256, 0, stream>>>(A, static_cast<const __nv_bfloat16*>(W), bias, C, M, N, K);
    else gemm_bf16w_kernel<false><<<grid, 256, 0, stream>>>(A, static_cast<const __nv_bfloat16*>(W), bias, C, M, N, K);
    return status_of(cudaGetLastError());
}

extern "C" int fots_b200_bilstm_recurrent(const float* G, const void* Whh, float* Y, int T, int N, int H, cudaStream_t stream) {
    if (T < 0 || N < 0 || H != LH || ((T > 0 && N > 0) && (!G || !Whh || !Y))) return RROI_B200_ERR_INVALID_ARG;
    if (((uintptr_t)G & 7) || ((uintptr_t)Whh & 15)) return RROI_B200_ERR_INVALID_ARG;
    if (T == 0 || N == 0) return RROI_B200_OK;
    static bool done[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return RROI_B200_ERR_CUDA;
    if (dev < 0 || dev >= 64 || !done[dev]) {
        const cudaError_t e = cudaFuncSetAttribute(bilstm_recurrent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLstmSmem);
        if (e != cudaSuccess) return status_of(e);
        if (dev >= 0 && dev < 64) done[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * LCL, (unsigned)((N + LNB - 1) / LNB));
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = kLstmSmem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = LCL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return status_of(cudaLaunchKernelEx(&cfg, bilstm_recurrent_kernel, G, static_cast<const __nv_bfloat16*>(Whh), Y, T, N));
}
