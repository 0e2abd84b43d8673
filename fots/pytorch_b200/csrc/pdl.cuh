// pdl.cuh -- programmatic dependent launch for the kernels of the end-to-end step.
//
// The 8-image step is ~160 kernel launches, half of them on maps of a few MB where a launch costs as much as the data
// (drain of the previous grid -> launch -> prologue -> ramp).  Launched with cudaLaunchAttributeProgrammaticStreamSerialization
// a kernel may be scheduled while its predecessor in the stream is still running; its CTAs do their input-independent
// prologue (barrier / TMEM set-up, constant weights, index arithmetic) and then block in `griddepcontrol.wait` until the
// predecessor has completed and its writes are visible.  Rules every kernel launched through pdl::launch follows:
//   * pdl::trigger() first (lets ITS successor be scheduled as soon as all of this grid's CTAs are resident),
//   * pdl::wait() before the first access to anything another kernel of the stream may have written or may still read
//     (activations, statistics workspaces, outputs) -- constant weights may be read before it.
// Stream capture records these launches as programmatic edges of the CUDA graph.
// MEASURED on B200 (tools/step_time.py, CUDA-graph replay of the end-to-end step): no gain -- graph replay already hides the
// launch latency, what a small launch costs is its own chain of DRAM round trips -- and a LOSS on large batches, where the
// early-resident waiting CTAs get in the way of the running grid: 8-image step 3.787 (on) vs 3.805 ms (off), 32-image
// micro-batch 14.28 (on) vs 13.74 ms (off).  So it is OFF by default (launches are fully stream-serialised and the in-kernel
// instructions are no-ops); FOTS_B200_PDL=1 switches it on (eager, launch-bound callers).
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>

namespace pdl {

__device__ __forceinline__ void wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool enabled() {
    static const bool on = [] { const char* e = getenv("FOTS_B200_PDL"); return e && e[0] == '1'; }();
    return on;
}

// kernel<<<grid, block, smem, stream>>>(args...) with the programmatic-serialisation attribute (+ an optional cluster size)
template <typename... P, typename... A>
inline cudaError_t launch_cluster(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, unsigned cluster_x,
                                  A&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    unsigned n = 0;
    if (cluster_x > 1) {
        at[n].id = cudaLaunchAttributeClusterDimension;
        at[n].val.clusterDim.x = cluster_x; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1;
        ++n;
    }
    if (enabled()) {
        at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = at;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}

template <typename... P, typename... A>
inline cudaError_t launch(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A&&... args) {
    return launch_cluster(kernel, grid, block, smem, stream, 1u, static_cast<A&&>(args)...);
}

// Clearing a small statistics workspace as a KERNEL of the chain (a cudaMemsetAsync node would end the programmatic edges)
static __global__ void zero_f64_kernel(double* p, int n) {
    trigger();
    wait();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = 0.0;
}
inline cudaError_t zero_f64(double* p, size_t n, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    if (n > (size_t)1 << 30) return cudaMemsetAsync(p, 0, n * sizeof(double), stream);
    unsigned blocks = (unsigned)((n + 255) / 256);
    if (blocks > 148u * 4u) blocks = 148u * 4u;
    return launch(zero_f64_kernel, dim3(blocks), dim3(256), 0, stream, p, (int)n);
}

}  // namespace pdl
