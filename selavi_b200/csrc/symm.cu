// Symmetric (peer-mapped) device buffers for in-kernel NVSwitch P2P exchange.
// One process per GPU: each rank cudaMalloc's a buffer, exports a CUDA IPC handle, the Python side
// all-gathers the 64-byte handles over torch.distributed and every rank maps its peers' buffers.
// (Replaces the NCCL all-gather/barrier pairs of src/sk_utils.py:214-254 for the SK exchange step.)
#include <string.h>

#include "../../include/selavi_b200.h"
#include "common.cuh"

extern "C" int selavi_symm_alloc(size_t bytes, void** ptr_out, unsigned char* handle64) {
    if (!ptr_out || !handle64 || bytes == 0) return selavi_fail(-1, "symm_alloc: bad arguments");
    void* p = nullptr;
    SV_CUDA_CHECK(cudaMalloc(&p, bytes), "symm_alloc: cudaMalloc");
    SV_CUDA_CHECK(cudaMemset(p, 0, bytes), "symm_alloc: cudaMemset");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return selavi_cuda_fail(e, "symm_alloc: cudaIpcGetMemHandle");
    }
    static_assert(sizeof(h) == 64, "CUDA IPC handle is 64 bytes");
    memcpy(handle64, &h, 64);
    *ptr_out = p;
    return 0;
}

extern "C" int selavi_symm_open(const unsigned char* handle64, void** ptr_out) {
    if (!ptr_out || !handle64) return selavi_fail(-1, "symm_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    SV_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "symm_open: cudaIpcOpenMemHandle");
    *ptr_out = p;
    return 0;
}

extern "C" int selavi_symm_close(void* ptr) {
    SV_CUDA_CHECK(cudaIpcCloseMemHandle(ptr), "symm_close: cudaIpcCloseMemHandle");
    return 0;
}

extern "C" int selavi_symm_free(void* ptr) {
    SV_CUDA_CHECK(cudaFree(ptr), "symm_free: cudaFree");
    return 0;
}

extern "C" int selavi_symm_memset(void* ptr, int value, size_t bytes, void* stream) {
    SV_CUDA_CHECK(cudaMemsetAsync(ptr, value, bytes, (cudaStream_t)stream), "symm_memset");
    return 0;
}
