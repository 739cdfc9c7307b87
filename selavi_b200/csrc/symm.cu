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

// ------------------------------------------------------------------------------------------------------------
// One-shot all-reduce (sum) of a small float64 vector over NVSwitch peer memory — the cross-rank exchange of the
// SyncBatchNorm statistics (torch:nn/modules/_functions.py:49-117,144-200 use an NCCL all_gather / all_reduce per
// BN layer; ~140 latency-bound collectives per step).  PUSH model: the single CTA stores its vector into its slot
// of EVERY rank's receive ring (posted stores), raises a flag there, then waits for the other ranks' flags in its
// own memory and sums the slots in rank order (bit-identical result on every rank).  Ring offsets / flag values
// are supplied by the host, which issues the same call sequence on every rank.
namespace {
struct P2PPtrs {
    double* recv[8];
    unsigned long long* flag[8];
};

__global__ void p2p_allreduce_f64_kernel(double* data, int n, int world, int rank, P2PPtrs ptrs, size_t slot_off, int flag_idx,
                                         unsigned long long seqval) {
    const int tid = threadIdx.x;
    for (int idx = tid; idx < world * n; idx += blockDim.x) {
        const int pr = idx / n, k = idx - pr * n;
        ptrs.recv[pr][slot_off + (size_t)rank * n + k] = data[k];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < world) {
        unsigned long long* f = ptrs.flag[tid] + (size_t)flag_idx * world + rank;
        asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(f), "l"(seqval) : "memory");
        const unsigned long long* mine = ptrs.flag[rank] + (size_t)flag_idx * world + tid;
        unsigned long long v;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(mine) : "memory");
        } while (v < seqval);
    }
    __syncthreads();
    const double* slots = ptrs.recv[rank] + slot_off;
    for (int k = tid; k < n; k += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < world; ++r) {
            double v;
            asm volatile("ld.volatile.global.f64 %0, [%1];\n" : "=d"(v) : "l"(slots + (size_t)r * n + k) : "memory");
            s += v;
        }
        data[k] = s;
    }
}
}  // namespace

extern "C" int selavi_p2p_allreduce_f64(double* data, int n, int world, int rank, void* const* peer_recv,
                                        void* const* peer_flag, long long slot_off, int flag_idx, long long seqval,
                                        void* stream) {
    if (!data || n <= 0 || world < 2 || world > 8 || rank < 0 || rank >= world || !peer_recv || !peer_flag || slot_off < 0 ||
        flag_idx < 0 || seqval <= 0)
        return selavi_fail(-1, "p2p_allreduce_f64: bad arguments");
    P2PPtrs p;
    for (int r = 0; r < 8; ++r) {
        p.recv[r] = r < world ? reinterpret_cast<double*>(peer_recv[r]) : nullptr;
        p.flag[r] = r < world ? reinterpret_cast<unsigned long long*>(peer_flag[r]) : nullptr;
    }
    p2p_allreduce_f64_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(data, n, world, rank, p, (size_t)slot_off, flag_idx,
                                                                   (unsigned long long)seqval);
    SV_CUDA_CHECK(cudaGetLastError(), "p2p_allreduce_f64: launch");
    return 0;
}
