#include <stdio.h>

#include "../../include/selavi_b200.h"
#include "common.cuh"

static thread_local char g_err[512] = "";

extern "C" int selavi_version(void) { return 100; }
extern "C" const char* selavi_last_error(void) { return g_err; }

int selavi_fail(int code, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}

int selavi_cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    cudaGetLastError();  // clear the sticky-free error state
    return -(1000 + (int)e);
}
