// pdl.cuh - programmatic dependent launch (PDL) helpers.
//
// The window is ~400 small dependent kernels: launch latency and kernel prologues are a first-order
// cost.  Every kernel therefore (a) signals `griddepcontrol.launch_dependents` on entry so the next
// kernel of its lane can be scheduled and run its prologue (index math, barrier init, TMEM
// allocation, descriptor prefetch) while this one is still executing, and (b) executes
// `griddepcontrol.wait` before its first read of upstream data / first global write, which blocks
// until all prerequisite grids have completed and flushed.  Launches go through cudaLaunchKernelEx
// with cudaLaunchAttributeProgrammaticStreamSerialization; stream capture turns that into
// programmatic edges of the CUDA graph.
#pragma once
#include <cuda_runtime.h>

#include <utility>

namespace rvc {

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// common prologue of the simple kernels: nothing to overlap but the launch latency itself
__device__ __forceinline__ void pdl_enter() { pdl_launch_dependents(); pdl_wait(); }

extern bool g_use_pdl;  // RVC_PDL=0 disables (engine.cu)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = g_use_pdl ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
}

}  // namespace rvc
