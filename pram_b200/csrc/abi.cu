// Library-level entry points of libpram_b200.so.
#include "common.cuh"
#include <stdlib.h>

unsigned long long g_pram_launches = 0;
thread_local const int* g_pram_pred = nullptr;
// programmatic dependent launch of the persistent tensor-core kernels (common.cuh); PRAM_PDL=0 in the environment disables it
int g_pram_pdl = [] { const char* e = getenv("PRAM_PDL"); return e ? atoi(e) : 1; }();

PRAM_API int pram_set_pdl(int enable) { g_pram_pdl = enable < 0 ? 0 : enable; return PRAM_OK; }
PRAM_API int pram_get_pdl(void) { return g_pram_pdl; }

// flag != NULL: kernels launched by this thread from now on (GEMM, attention, block tail, Linear, LayerNorm, the AdaGML
// control kernels) skip their work when *flag == 0 at execution time; NULL clears the predicate.
PRAM_API int pram_set_launch_predicate(const int* flag) { g_pram_pred = flag; return PRAM_OK; }

PRAM_API int pram_version(void) { return 100; }  // 0.1.0

// Number of kernels this library has launched so far in this process (bench.py: gpu_launches).
PRAM_API unsigned long long pram_launch_count(void) { return g_pram_launches; }

PRAM_API const char* pram_error_string(int code) {
    switch (code) {
        case PRAM_OK: return "ok";
        case PRAM_ERR_ARG: return "invalid argument";
        case PRAM_ERR_CUDA: return "CUDA error";
        case PRAM_ERR_WORKSPACE: return "workspace too small";
        case PRAM_ERR_UNSUPPORTED: return "unsupported configuration";
        default: return "unknown";
    }
}

PRAM_API const char* pram_last_cuda_error(void) { return cudaGetErrorString(cudaGetLastError()); }
