// Shared host-side helpers of the C-ABI library: error reporting.
#pragma once
#include <cuda_runtime.h>

// records msg for selavi_last_error() and returns code (always negative)
int selavi_fail(int code, const char* msg);
// records "<what>: <cudaGetErrorString(e)>" and returns -(1000 + e)
int selavi_cuda_fail(cudaError_t e, const char* what);

#define SV_CUDA_CHECK(expr, what)                         \
    do {                                                  \
        cudaError_t _e = (expr);                          \
        if (_e != cudaSuccess) return selavi_cuda_fail(_e, what); \
    } while (0)
