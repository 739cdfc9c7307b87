// Micro-benchmark: issue rate of back-to-back tcgen05.mma (cta_group::1, M = 128, SS mode, operands resident in shared
// memory, no loads, no epilogue) as a function of N, operand kind and descriptor pattern.  Built and run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I selavi_b200/csrc tools/umma_rate.cu -o /tmp/umma_rate
// Prints clocks per MMA and the fraction of the N/2-clock floor (B300_MICROARCH.md: 128*N/256 cycles per K=32-byte step).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#include "ptx.cuh"

struct Cfg {
    int N;          // MMA N
    int kind;       // 0 tf32, 1 f16
    int pattern;    // 0: same descriptors every MMA; 1: 4 k-steps of 32 B inside one SW128 atom row (real loop);
                    // 2: pattern 1 + the A window moves by 128-byte rows every 12 MMAs (tap shifts); 3: like 1, 3 MMAs share k
    int swizzle;    // 2 = SWIZZLE_128B, 0 = none (core matrices 8 x 16 B)
    int iters;
};

__global__ void __launch_bounds__(160, 1) rate_kernel(Cfg c, long long* out_clk) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    // zero operands (values do not matter for timing; zeros avoid NaN paths)
    for (int i = tid; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        sv::mbar_init(&bar, 1);
        sv::fence_barrier_init();
    }
    if (warp == 4) {
        sv::tmem_alloc(&tmem_slot, 512);
        sv::tmem_relinquish();
    }
    sv::fence_proxy_async();
    sv::tc_fence_before();
    __syncthreads();
    sv::tc_fence_after();
    if (warp == 4) {
        const uint32_t tm0 = __shfl_sync(0xffffffffu, tmem_slot, 0);
        const uint32_t idesc = c.kind == 0 ? sv::make_idesc_tf32(128, c.N, 0, 0) : sv::make_idesc_f16(128, c.N, 0, 0, 0, 0);
        const uint32_t a_addr = sv::smem_u32(smem), b_addr = a_addr + 96 * 1024;
        const uint64_t fixed = c.swizzle == 2 ? sv::make_smem_desc_sw128(0, 16, 1024) : sv::make_smem_desc(0, 128, 256, 0);
        const uint64_t da0 = fixed | (uint64_t)((a_addr & 0x3FFFFu) >> 4);
        const uint64_t db0 = fixed | (uint64_t)((b_addr & 0x3FFFFu) >> 4);
        const uint64_t a_lo = 32768 >> 4, b_lo = 32768 >> 4;   // "lo" planes 32 KB further
        long long t0 = 0, t1 = 0;
        for (int rep = 0; rep < 2; ++rep) {   // rep 0 warms up
            t0 = clock64();
            if (sv::elect_one()) {
                for (int i = 0; i < c.iters; ++i) {
                    uint64_t da = da0, db = db0;
                    if (c.pattern >= 1) {   // (pattern 4 alternates tm0/tm1/tm0 | tm0/tm1/tm0: two of three hand-offs switch accumulator)
                        const int k = i & 3;
                        da += 2 * k;
                        db += 2 * k;
                    }
                    if (c.pattern == 2) da += (uint64_t)(((i >> 2) % 9) * 8 * 58);   // row shift of 58 rows x 128 B per "tap"
                    if (c.pattern == 4) {
                        // x3 group, consecutive MMAs alternate between two accumulators (no back-to-back dependency)
                        const uint32_t tm1 = tm0 + (uint32_t)c.N;
                        if (c.kind == 0) {
                            sv::umma_tf32(tm0, da + a_lo, db, idesc, 1u);
                            sv::umma_tf32(tm1, da, db + b_lo, idesc, 1u);
                            sv::umma_tf32(tm0, da, db, idesc, 1u);
                        } else {
                            sv::umma_f16(tm0, da + a_lo, db, idesc, 1u);
                            sv::umma_f16(tm1, da, db + b_lo, idesc, 1u);
                            sv::umma_f16(tm0, da, db, idesc, 1u);
                        }
                    } else if (c.pattern == 3) {
                        // the real x3 group: (a_lo,b_hi) (a_hi,b_lo) (a_hi,b_hi) on the same k
                        if (c.kind == 0) {
                            sv::umma_tf32(tm0, da + a_lo, db, idesc, 1u);
                            sv::umma_tf32(tm0, da, db + b_lo, idesc, 1u);
                            sv::umma_tf32(tm0, da, db, idesc, 1u);
                        } else {
                            sv::umma_f16(tm0, da + a_lo, db, idesc, 1u);
                            sv::umma_f16(tm0, da, db + b_lo, idesc, 1u);
                            sv::umma_f16(tm0, da, db, idesc, 1u);
                        }
                    } else if (c.kind == 0) {
                        sv::umma_tf32(tm0, da, db, idesc, 1u);
                    } else {
                        sv::umma_f16(tm0, da, db, idesc, 1u);
                    }
                }
                sv::umma_commit(&bar);
            }
            __syncwarp();
            sv::mbar_wait(&bar, (uint32_t)rep);
            sv::tc_fence_after();
            t1 = clock64();
        }
        if ((tid & 31) == 0) out_clk[blockIdx.x] = t1 - t0;
    }
    __syncthreads();
    if (warp == 4) {
        sv::tc_fence_after();
        sv::tmem_dealloc(tmem_slot, 512);
    }
}

int main() {
    long long* d_clk;
    cudaMalloc(&d_clk, 148 * sizeof(long long));
    const size_t smem = 200 * 1024;
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int Ns[] = {32, 64, 128, 144, 192, 256};
    const char* pat[] = {"same desc", "4 k-steps", "k-steps+tap shift", "x3 group (hi/lo planes)", "x3 group, 2 accumulators"};
    for (int grid : {148}) {
        for (int kind : {1, 0}) {
            for (int swz : {2, 0}) {
                for (int pattern = 0; pattern < 5; ++pattern) {
                    if (swz == 0 && pattern != 0) continue;
                    printf("grid %3d kind %s swizzle %d pattern '%s':", grid, kind ? "f16 " : "tf32", swz, pat[pattern]);
                    for (int N : Ns) {
                        Cfg c{N, kind, pattern, swz, 2048};
                        rate_kernel<<<grid, 160, smem>>>(c, d_clk);
                        if (cudaDeviceSynchronize() != cudaSuccess) {
                            printf(" launch failed: %s\n", cudaGetErrorString(cudaGetLastError()));
                            return 1;
                        }
                        long long h[148];
                        cudaMemcpy(h, d_clk, grid * sizeof(long long), cudaMemcpyDeviceToHost);
                        long long mx = 0;
                        for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
                        const int n_mma = c.iters * (pattern >= 3 ? 3 : 1);
                        const double per = (double)mx / n_mma;
                        const double floor_clk = kind ? N / 2.0 : N / 2.0;   // per K = 32 bytes in both kinds
                        printf("  N=%d %.1f clk (%.0f%%)", N, per, 100.0 * floor_clk / per);
                    }
                    printf("\n");
                }
            }
        }
    }
    return 0;
}
