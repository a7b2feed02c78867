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

// Lane stamps: %globaltimer written by the first thread of a few marker kernels (STFT start, F0 decode, retrieval gather,
// pitch cache, conv_post) into a per-file device array - where the two front-end lanes and the tail stand inside a
// graph-replayed window WITHOUT event nodes in the graph (rvc_debug_lane_stamps; an event after an op perturbs the lanes).
__device__ __forceinline__ void lane_stamp(unsigned long long* slot) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        *slot = t;
    }
}

extern bool g_use_pdl;  // RVC_PDL=1: every launch (engine.cu)
extern thread_local bool g_pdl_op;   // this op only (RVC_PDL_OPS, engine.cu issue_ops)
// Launch priority of the kernels issued next by this thread (0 = default).  engine.cu raises it for the ops of the
// F0 lanes: the F0 chain is the longest branch of the window and must not queue behind ContentVec's wide grids.
// Set as a launch attribute so that it survives stream capture (kernel-node attribute of the CUDA graph).
extern thread_local int g_launch_priority;

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = (g_use_pdl || g_pdl_op) ? 1 : 0;
    attr[1].id = cudaLaunchAttributePriority;
    attr[1].val.priority = g_launch_priority;
    cfg.attrs = attr; cfg.numAttrs = g_launch_priority != 0 ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
}

}  // namespace rvc
