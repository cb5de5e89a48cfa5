// Shared device/host helpers for libpram_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <math.h>

#define PRAM_OK 0
#define PRAM_ERR_ARG (-1)
#define PRAM_ERR_CUDA (-2)
#define PRAM_ERR_WORKSPACE (-3)
#define PRAM_ERR_UNSUPPORTED (-4)

#define PRAM_API extern "C" __attribute__((visibility("default")))

// Launch-counter: every kernel launch of this library goes through PRAM_LAUNCH so that
// bench.py can report how many of OUR kernels ran inside the timed region (pram_launch_count()).
extern unsigned long long g_pram_launches;

#define PRAM_CHECK_LAUNCH()                                      \
    do {                                                         \
        ++g_pram_launches;                                       \
        cudaError_t e__ = cudaPeekAtLastError();                 \
        if (e__ != cudaSuccess) return PRAM_ERR_CUDA;            \
    } while (0)

#define PRAM_CUDA(call)                                          \
    do {                                                         \
        cudaError_t e__ = (call);                                \
        if (e__ != cudaSuccess) return PRAM_ERR_CUDA;            \
    } while (0)

// Launch predicate (pram_set_launch_predicate): while a device flag is registered, the kernels that take part in a
// device-side early exit (AdaGML's stop test, csrc/adagml_ops.cu) are launched with it and return at once when the flag
// reads 0 -- the device-resident replacement of the reference's host-side `break` (nets/adagml.py:370-372).
extern thread_local const int* g_pram_pred;
__device__ __forceinline__ bool pram_pred_skip(const int* pred) { return pred != nullptr && __ldg(pred) == 0; }

// Programmatic dependent launch (pram_set_pdl): the persistent tensor-core kernels are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so kernel N + 1 of a stream (also inside a captured CUDA graph) is
// scheduled while kernel N drains: its CTAs start on every SM the moment N's CTA there exits, set up their barriers / TMEM /
// descriptors, and only then wait (griddepcontrol.wait) for N to complete and flush.  Every kernel launched that way
// executes pram_pdl_wait() on all threads before its first global-memory access (results AND visibility: the wait is also
// what makes the chain N -> N + 1 -> N + 2 transitive), and calls pram_pdl_trigger() right after it so that its own
// dependents may be scheduled as soon as all of its CTAs are resident.  Both are no-ops for a normally launched kernel.
extern int g_pram_pdl;
__device__ __forceinline__ void pram_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pram_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
static inline cudaError_t pram_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                          Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = g_pram_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#endif

__host__ __device__ static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
