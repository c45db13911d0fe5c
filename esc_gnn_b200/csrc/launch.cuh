// Programmatic dependent launch (PDL) for the chains of small dependent kernels of the training step.
//
// Every kernel launched through launch_pdl() starts with pdl_enter(): `griddepcontrol.wait` blocks until the
// preceding grid on the stream has completed and its writes are visible, `griddepcontrol.launch_dependents` then lets
// the NEXT kernel's CTAs be scheduled while this one runs (they park in their own wait).  Nothing before the wait
// touches global memory, so the semantics are those of plain stream order; what is saved is the launch latency
// between two dependent kernels, which at the reference's batch sizes is a large share of a step.
// Stream capture turns these launches into programmatic graph edges.
#pragma once
#include <cuda_runtime.h>

namespace escgnn {

__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
// split form for kernels with a memory-free prologue worth overlapping (barrier init, TMEM allocation)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline int& pdl_enabled() {
    static int on = 1;
    return on;
}

template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// plain stream order (no programmatic edge): for a kernel whose predecessor on the stream is NOT a kernel (a memset / copy
// node) -- griddepcontrol.wait only covers prerequisite GRIDS
template <class... KArgs, class... Args>
inline cudaError_t launch_ordered(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cfg.attrs = nullptr;
    cfg.numAttrs = 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// same, as thread-block clusters of (1, cluster_y, 1) CTAs (gridDim.y must be a multiple of cluster_y)
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, unsigned cluster_y,
                                      Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = cluster_y;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace escgnn
