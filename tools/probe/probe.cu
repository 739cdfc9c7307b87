// Diagnostics: single-CTA tcgen05.mma probe.  The host supplies raw shared-memory images of the A and B
// operand tiles, the descriptor bits and per-instruction start offsets; the kernel issues the MMAs and dumps
// the 128 x N fp32 accumulator.  Used by tools/umma_probe.py to pin down operand-layout conventions
// (K-major / MN-major, swizzle modes, LBO/SBO meaning) on real hardware.
#include <stdint.h>

// Built into its own library (tools/probe/libselavi_probe.so, see tools/probe/__init__.py): diagnostics are not part of
// the product C ABI.
#include <stdio.h>

#include "../../selavi_b200/csrc/ptx.cuh"

static int probe_fail(int code, const char* msg) {
    fprintf(stderr, "selavi_debug_umma_probe: %s\n", msg);
    return code;
}
#define selavi_fail probe_fail
#define SV_CUDA_CHECK(expr, what)                                              \
    do {                                                                       \
        cudaError_t _e = (expr);                                               \
        if (_e != cudaSuccess) return probe_fail(-(1000 + (int)_e), what);     \
    } while (0)

namespace {
__global__ void __launch_bounds__(160, 1)
umma_probe_kernel(const uint4* a_img, int a_bytes, const uint4* b_img, int b_bytes, uint64_t adesc_base,
                  uint64_t bdesc_base, uint32_t idesc, int n_mma, const uint32_t* a_offs, const uint32_t* b_offs,
                  int N, int kind, float* out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* a_s = smem;
    unsigned char* b_s = smem + ((a_bytes + 1023) & ~1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < a_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(a_s)[i] = a_img[i];
    for (int i = tid; i < b_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(b_s)[i] = b_img[i];
    if (tid == 0) {
        sv::mbar_init(&bar, 1);
        sv::fence_barrier_init();
    }
    if (warp == 4) {
        sv::tmem_alloc(&tmem_slot, 256);
        sv::tmem_relinquish();
    }
    sv::fence_proxy_async();
    sv::tc_fence_before();
    __syncthreads();
    sv::tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (warp == 4 && lane == 0) {
        for (int i = 0; i < n_mma; ++i) {
            const uint64_t da = adesc_base | (uint64_t)(((sv::smem_u32(a_s) + a_offs[i]) & 0x3FFFFu) >> 4);
            const uint64_t db = bdesc_base | (uint64_t)(((sv::smem_u32(b_s) + b_offs[i]) & 0x3FFFFu) >> 4);
            if (kind == 0) sv::umma_tf32(tmem_base, da, db, idesc, i ? 1u : 0u);
            else sv::umma_f16(tmem_base, da, db, idesc, i ? 1u : 0u);
        }
        sv::umma_commit(&bar);
    }
    if (warp < 4) {
        sv::mbar_wait(&bar, 0);
        sv::tc_fence_after();
        for (int u = 0; u < N / 8; ++u) {
            uint32_t v[8];
            sv::tmem_ld8(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(u * 8), v);
            sv::tmem_ld_wait();
            for (int i = 0; i < 8; ++i) out[(size_t)(warp * 32 + lane) * N + u * 8 + i] = __uint_as_float(v[i]);
        }
        sv::tc_fence_before();
    }
    __syncthreads();
    if (warp == 4) {
        sv::tc_fence_after();
        sv::tmem_dealloc(tmem_base, 256);
    }
}
}  // namespace

extern "C" int selavi_debug_umma_probe(const void* a_img, int a_bytes, const void* b_img, int b_bytes,
                                       unsigned long long adesc_base, unsigned long long bdesc_base, unsigned idesc,
                                       int n_mma, const unsigned* a_offs, const unsigned* b_offs, int N, int kind,
                                       float* out, void* stream) {
    if (!a_img || !b_img || !a_offs || !b_offs || !out || (a_bytes & 15) || (b_bytes & 15) || N % 8 || N > 256)
        return selavi_fail(-1, "umma_probe: bad arguments");
    const size_t smem = (size_t)((a_bytes + 1023) & ~1023) + ((b_bytes + 1023) & ~1023) + 2048;
    if (smem > 220 * 1024) return selavi_fail(-1, "umma_probe: images too large");
    SV_CUDA_CHECK(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                  "umma_probe: attr");
    umma_probe_kernel<<<1, 160, smem, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4*>(a_img), a_bytes, reinterpret_cast<const uint4*>(b_img), b_bytes, adesc_base,
        bdesc_base, idesc, n_mma, a_offs, b_offs, N, kind, out);
    SV_CUDA_CHECK(cudaGetLastError(), "umma_probe: launch");
    return 0;
}
