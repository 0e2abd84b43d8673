// pipeline_kernels.cu -- the small device-side glue either side of RoIRotate (include/fots_b200_pipeline.h):
// quad -> RoI row, and greedy CTC decode.  Both replace per-box host loops of the reference
// (tools/ocr_utils.py:133-145, :183-186; src/utils.py:93-97) so that the end-to-end path has no host sync.
#include "../../../include/fots_b200_pipeline.h"
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__global__ void boxes_to_rois_kernel(const float* __restrict__ quads, int stride, const int* __restrict__ bidx,
                                     int n, float* __restrict__ rois) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* q = quads + (size_t)i * stride;
    // The reference does this arithmetic on numpy float32 scalars (every + - * rounds to fp32) and only
    // sqrt/atan2 and the degree conversion in Python doubles (tools/ocr_utils.py:136-144).
    const float cxf = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(q[0], q[2]), q[4]), q[6]), 4.0f);
    const float cyf = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(q[1], q[3]), q[5]), q[7]), 4.0f);
    const float dwx = __fsub_rn(q[4], q[2]), dwy = __fsub_rn(q[5], q[3]);
    const float dhx = __fsub_rn(q[2], q[0]), dhy = __fsub_rn(q[3], q[1]);
    const double w = sqrt((double)__fadd_rn(__fmul_rn(dwx, dwx), __fmul_rn(dwy, dwy)));
    const double h = sqrt((double)__fadd_rn(__fmul_rn(dhx, dhx), __fmul_rn(dhy, dhy)));
    const double ang = -atan2((double)dwy, (double)dwx) / 3.1415926535 * 180.0;
    float* r = rois + (size_t)i * 6;
    r[0] = bidx ? (float)bidx[i] : 0.0f;
    r[1] = truncf(cxf);          // int(center[0])
    r[2] = truncf(cyf);
    r[3] = (float)h;
    r[4] = (float)w;
    r[5] = (float)ang;
}

// one CTA per sequence, one thread per time step (coalesced along t for every class)
__global__ void ctc_greedy_kernel(const float* __restrict__ logp, int C, int T, int* __restrict__ ids,
                                  int* __restrict__ lengths) {
    __shared__ int s_best[1024];
    __shared__ int s_warp[32];
    const int n = blockIdx.x, t = threadIdx.x;
    const float* p = logp + (size_t)n * C * T;
    int best = 0;
    if (t < T) {
        float bv = p[t];
        for (int c = 1; c < C; ++c) {
            const float v = p[(size_t)c * T + t];
            if (v > bv) { bv = v; best = c; }      // strict '>' keeps the lowest index on ties
        }
        s_best[t] = best;
    }
    __syncthreads();
    const bool keep = (t < T) && best != 0 && !(t > 0 && s_best[t - 1] == best);
    // block-wide exclusive scan of `keep`
    const unsigned lane = t & 31, wid = t >> 5;
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    const int within = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) s_warp[wid] = __popc(bal);
    __syncthreads();
    if (wid == 0) {
        int v = (lane < (blockDim.x + 31) / 32) ? s_warp[lane] : 0;
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= (unsigned)d) v += o;
        }
        s_warp[lane] = v;        // inclusive
    }
    __syncthreads();
    const int base = wid == 0 ? 0 : s_warp[wid - 1];
    const int total = s_warp[(blockDim.x + 31) / 32 - 1];
    int* out = ids + (size_t)n * T;
    if (keep) out[base + within] = best;
    if (t < T && t >= total) out[t] = 0;           // zero padding
    if (t == 0) lengths[n] = total;
}

// ---- detection decode: nms/adaptor.cpp:76-117 -------------------------------------------------------------------
constexpr int kDecBlock = 256;

__global__ void __launch_bounds__(kDecBlock) decode_count_kernel(const float* __restrict__ segm, int hw, float thr,
                                                                 int* __restrict__ block_counts) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * kDecBlock + threadIdx.x;
    const bool pos = i < hw && segm[(size_t)b * hw + i] > thr;
    const int n = __syncthreads_count(pos);
    if (threadIdx.x == 0) block_counts[b * gridDim.x + blockIdx.x] = n;
}

__global__ void __launch_bounds__(kDecBlock) decode_write_kernel(const float* __restrict__ segm, const float* __restrict__ rbox,
                                                                 const float* __restrict__ angle, int h, int w, float thr,
                                                                 int max_per_image, const int* __restrict__ block_counts,
                                                                 int* __restrict__ counts, int* __restrict__ cand) {
    __shared__ int s_red[kDecBlock / 32];
    __shared__ int s_base;
    const int b = blockIdx.y, hw = h * w, nblk = gridDim.x;
    // exclusive prefix of the per-block counts of this image (raster order = block order)
    int part = 0;
    for (int k = threadIdx.x; k < (int)blockIdx.x; k += kDecBlock) part += block_counts[b * nblk + k];
    for (int d = 16; d > 0; d >>= 1) part += __shfl_down_sync(0xffffffffu, part, d);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < kDecBlock / 32; ++k) t += s_red[k];
        s_base = t;
    }
    __syncthreads();
    const int i = blockIdx.x * kDecBlock + threadIdx.x;
    const float score = i < hw ? segm[(size_t)b * hw + i] : 0.0f;
    const bool pos = i < hw && score > thr;
    const unsigned bal = __ballot_sync(0xffffffffu, pos);
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_red[wid] = __popc(bal);
    __syncthreads();
    int rank = s_base + __popc(bal & ((1u << lane) - 1u));
    for (unsigned k = 0; k < wid; ++k) rank += s_red[k];
    if (blockIdx.x == nblk - 1 && threadIdx.x == kDecBlock - 1) {
        int t = rank + (pos ? 1 : 0);                       // last thread of the last block: total of the image
        counts[b] = t;
    }
    if (!pos || rank >= max_per_image) return;
    const int y = i / w, x = i - y * w;
    const float* r = rbox + (size_t)b * 4 * hw + i;        // planes: top, bottom, left, right
    const float r0 = r[0], r1 = r[hw], r2 = r[2 * (size_t)hw], r3 = r[3 * (size_t)hw];
    const float a_sin = angle[(size_t)b * 2 * hw + i], a_cos = angle[(size_t)b * 2 * hw + hw + i];
    const float sf = 4.0f, prec = 10000.0f;
    const float xp = __fadd_rn((float)x, 0.25f), yp = __fadd_rn((float)y, 0.25f);
    const float pos_r_x = __fmul_rn(__fsub_rn(xp, __fmul_rn(r2, a_cos)), sf);
    const float pos_r_y = __fmul_rn(__fsub_rn(yp, __fmul_rn(r2, a_sin)), sf);
    const float pos_r2_x = __fmul_rn(__fadd_rn(xp, __fmul_rn(r3, a_cos)), sf);
    const float pos_r2_y = __fmul_rn(__fadd_rn(yp, __fmul_rn(r3, a_sin)), sf);
    const float r1s = __fmul_rn(__fmul_rn(r1, a_sin), sf), r1c = __fmul_rn(__fmul_rn(r1, a_cos), sf);
    const float r0s = __fmul_rn(__fmul_rn(r0, a_sin), sf), r0c = __fmul_rn(__fmul_rn(r0, a_cos), sf);
    int* o = cand + ((size_t)b * max_per_image + rank) * 16;
    o[0] = (int)roundf(__fmul_rn(prec, __fsub_rn(pos_r_x, r1s)));
    o[1] = (int)roundf(__fmul_rn(prec, __fadd_rn(pos_r_y, r1c)));
    o[2] = (int)roundf(__fmul_rn(prec, __fadd_rn(pos_r_x, r0s)));
    o[3] = (int)roundf(__fmul_rn(prec, __fsub_rn(pos_r_y, r0c)));
    o[4] = (int)roundf(__fmul_rn(prec, __fadd_rn(pos_r2_x, r0s)));
    o[5] = (int)roundf(__fmul_rn(prec, __fsub_rn(pos_r2_y, r0c)));
    o[6] = (int)roundf(__fmul_rn(prec, __fsub_rn(pos_r2_x, r1s)));
    o[7] = (int)roundf(__fmul_rn(prec, __fadd_rn(pos_r2_y, r1c)));
    const float p_left = expf(__fdiv_rn(-r2, 9.0f)), p_top = expf(__fdiv_rn(-r0, 9.0f));
    const float p_right = expf(__fdiv_rn(-r3, 9.0f)), p_bt = expf(__fdiv_rn(-r1, 9.0f));
    o[8] = __float_as_int(score);
    o[9] = __float_as_int(__fmul_rn(p_left, p_bt));
    o[10] = __float_as_int(__fmul_rn(p_left, p_top));
    o[11] = __float_as_int(__fmul_rn(p_right, p_top));
    o[12] = __float_as_int(__fmul_rn(p_right, p_bt));
    o[13] = x; o[14] = y; o[15] = 0;
}

int status_of(cudaError_t e) {
    if (e == cudaSuccess) return RROI_B200_OK;
    (void)cudaGetLastError();
    return RROI_B200_ERR_CUDA;
}

}  // namespace

extern "C" {

int fots_b200_boxes_to_rois(const float* quads, int quad_stride, const int* batch_idx, int num_boxes,
                            float* rois, cudaStream_t stream) {
    if (num_boxes < 0 || quad_stride < 8 || (num_boxes > 0 && (!quads || !rois))) return RROI_B200_ERR_INVALID_ARG;
    if (num_boxes == 0) return RROI_B200_OK;
    boxes_to_rois_kernel<<<(num_boxes + 127) / 128, 128, 0, stream>>>(quads, quad_stride, batch_idx, num_boxes, rois);
    return status_of(cudaGetLastError());
}

int fots_b200_ctc_greedy(const float* logp, int num_seq, int num_classes, int T, int* ids, int* lengths,
                         cudaStream_t stream) {
    if (num_seq < 0 || num_classes <= 0 || T <= 0 || T > 1024 || (num_seq > 0 && (!logp || !ids || !lengths)))
        return RROI_B200_ERR_INVALID_ARG;
    if (num_seq == 0) return RROI_B200_OK;
    const int threads = ((T + 31) / 32) * 32;
    ctc_greedy_kernel<<<num_seq, threads, 0, stream>>>(logp, num_classes, T, ids, lengths);
    return status_of(cudaGetLastError());
}

int fots_b200_decode_candidates(const float* segm, const float* rbox, const float* angle, int B, int h, int w,
                                float segm_threshold, int max_per_image, int* counts, int* cand, int* scratch,
                                cudaStream_t stream) {
    if (!segm || !rbox || !angle || !counts || !cand || !scratch || B <= 0 || h <= 0 || w <= 0 || max_per_image <= 0 ||
        (long long)h * w > 0x7fffffffLL)
        return RROI_B200_ERR_INVALID_ARG;
    const int hw = h * w;
    const dim3 grid((hw + kDecBlock - 1) / kDecBlock, B);
    decode_count_kernel<<<grid, kDecBlock, 0, stream>>>(segm, hw, segm_threshold, scratch);
    decode_write_kernel<<<grid, kDecBlock, 0, stream>>>(segm, rbox, angle, h, w, segm_threshold, max_per_image, scratch, counts, cand);
    return status_of(cudaGetLastError());
}

}  // extern "C"
