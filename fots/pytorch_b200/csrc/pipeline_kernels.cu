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

}  // extern "C"
