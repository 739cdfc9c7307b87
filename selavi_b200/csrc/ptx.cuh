// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk[.tensor]), tcgen05 (UMMA/TMEM).
// Everything here is hand-written; nothing is taken from CUTLASS (its headers were only read as
// documentation of descriptor bit layouts).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sv {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// try_wait with a suspend-time hint (ns): the thread is parked by the hardware until the phase completes or the time is up,
// instead of burning issue slots in a spin loop (ncu on sk_kernel: half of all issued instructions were try_wait spins,
// and the spinning producer slowed the three consumer warps that share its scheduler)
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait_hint(bar, parity, 20000u)) {
    }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 1-D bulk copy global -> shared, completion on mbarrier (bytes multiple of 16, both 16B aligned)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// L2 prefetch of a contiguous global range (bytes multiple of 16, 16B aligned); no completion tracking
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(reinterpret_cast<uint64_t>(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::
            "r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
        "%6, %7}], [%2];\n" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::f16 (fp16/bf16 operands, fp32 accumulate). One thread issues.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::tf32 (fp32 words in smem, low 13 mantissa bits ignored, fp32 accumulate)
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrives when all previously issued MMAs of this thread have completed (implies fence::before).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
// ---------------------------------------------------------------- CTA pairs (cluster of 2, tcgen05 cta_group::2)
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {   // every thread of every CTA of the cluster
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
// arrive (count 1, release at cluster scope) on an mbarrier of any CTA of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(addr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[256 rows: 128 from each CTA's smem] * B[N columns: N/2 from each CTA's smem]; issued by one
// thread of the pair's leader CTA (rank 0), descriptors are shared-memory offsets valid in both CTAs
__device__ __forceinline__ void umma_f16_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the mbarrier at this offset in BOTH CTAs of the pair gets one arrival when the pair's previously issued MMAs retire
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
            smem_u32(bar)),
        "h"(static_cast<uint16_t>(3))
        : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i = lane i of the warp's quadrant)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
        "[%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// Shared-memory matrix descriptor (layout documented in DESIGN.md §UMMA descriptors):
//   bits [0,14)  start address >> 4      bits [16,30) leading-dim byte offset >> 4
//   bits [32,46) stride-dim byte offset >> 4   bits [46,48) version = 1 (Blackwell)
//   bits [61,64) layout type: 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// Same descriptor with an explicit layout type (0 = no swizzle / interleave, 2 = SWIZZLE_128B, 4 = 64B, 6 = 32B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(layout_type & 7u) << 61;
    return d;
}
// Instruction descriptor for kind::tf32 (a/b format 2), fp32 accumulate. major: 0 = K-major, 1 = MN-major.
__host__ __device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N, int amajor, int bmajor) {
    uint32_t d = 0;
    d |= 1u << 4;   // c_format = F32
    d |= 2u << 7;   // a_format = TF32
    d |= 2u << 10;  // b_format = TF32
    d |= static_cast<uint32_t>(amajor) << 15;
    d |= static_cast<uint32_t>(bmajor) << 16;
    d |= static_cast<uint32_t>(N >> 3) << 17;
    d |= static_cast<uint32_t>(M >> 4) << 24;
    return d;
}
// Instruction descriptor for kind::f16, fp32 accumulate.
//   afmt/bfmt: 0 = fp16, 1 = bf16.  amajor/bmajor: 0 = K-major, 1 = MN-major.
__host__ __device__ __forceinline__ uint32_t make_idesc_f16(int M, int N, int afmt, int bfmt, int amajor,
                                                            int bmajor) {
    uint32_t d = 0;
    d |= 1u << 4;                          // c_format = F32
    d |= static_cast<uint32_t>(afmt) << 7;   // a_format
    d |= static_cast<uint32_t>(bfmt) << 10;  // b_format
    d |= static_cast<uint32_t>(amajor) << 15;
    d |= static_cast<uint32_t>(bmajor) << 16;
    d |= static_cast<uint32_t>(N >> 3) << 17;
    d |= static_cast<uint32_t>(M >> 4) << 24;
    return d;
}


// ---------------------------------------------------------------- coalesced accumulator drain
// Bytes of per-warp staging the epilogues below need: 32 rows x 128 B of fp32 columns + 32 row indices.
constexpr int EPI_STAGE_BYTES = 32 * 128 + 32 * 4;

__device__ __forceinline__ void st_shared_f4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float ld_shared_f1(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

// One warp drains GW (32 or 16) fp32 accumulator columns of its 32-lane TMEM quadrant.  TMEM hands every lane ITS row;
// storing that directly makes each st.global.v4 touch 32 different 128-byte lines (measured: the layer-1 kernels were
// bound by exactly this, tools/halo_diag.py "no store").  Here the GW columns go through a 4 KB XOR-swizzled staging
// tile in shared memory, and the global stores (and the accumulate-mode reads) are issued with 4 (GW=32) or 8 (GW=16)
// complete row segments per instruction.  The per-column sums for the following BatchNorm (sum z, sum z^2 over the 32
// rows, invalid rows contribute exact zeros) are read off the same staging tile: lane c owns column c.
//   taddr   : TMEM address of the first column (lane field = quadrant base)
//   stg     : shared-memory address of this warp's staging tile, 128-byte aligned; row_pix = the 32 ints behind it
//             (global row index of each TMEM lane, -1 = no such row), written by the caller + __syncwarp()
//   col0    : destination column of the first accumulator column; columns >= cd are not stored
//   stat    : this warp's [2][256] sum scratch, already offset to the group's first column; nullptr = no statistics
// Statistics of the BatchNorm-BACKWARD pass that follows a data-gradient kernel, fused into its epilogue: the gradient g
// this kernel writes is the upstream gradient of the unit whose raw output is `z` (activation relu(z*scale+shift)), and
// that unit's backward needs sum(g*m) and sum(g*m*zhat) per channel (m = ReLU mask, zhat = (z-mean)*invstd).  Computing
// them here, where g sits in registers, saves the separate reduce pass that reads g and z again from HBM.
struct EpiBwdStat {
    const float* z;        // [pixels][cd]
    const float* scale;    // [cd] each
    const float* shift;
    const float* mean;
    const float* invstd;
};

template <int GW>
__device__ __forceinline__ void epi_drain_group(uint32_t taddr, float oscale, bool row_ok, uint32_t stg, const int* row_pix,
                                                float* __restrict__ dst, int cd, int col0, int accumulate, bool do_store,
                                                float* stat, int lane, const EpiBwdStat* bwd = nullptr) {
    static_assert(GW == 32 || GW == 16, "group width");
    constexpr int CPR = GW / 4;      // 16-byte chunks per row segment
    constexpr int RPI = 32 / CPR;    // row segments per store instruction
    const int j = lane % CPR, rsub = lane / CPR;
    const int col = col0 + j * 4;
    const bool want_bwd = bwd != nullptr && stat != nullptr;
    // fused BatchNorm-backward statistics: this lane's z values are requested FIRST, so that their HBM / L2 latency is
    // hidden behind the TMEM drain and the staging pass (issued inside the store loop they serialise behind the stores:
    // measured 2x on the whole data-gradient kernel)
    float4 zz[CPR];
    if (want_bwd) {
#pragma unroll
        for (int i = 0; i < CPR; ++i) {
            const int pix = row_pix[i * RPI + rsub];
            zz[i] = (pix >= 0 && col < cd) ? __ldg(reinterpret_cast<const float4*>(bwd->z + (size_t)pix * cd + col))
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    uint32_t av[GW];
    if constexpr (GW == 32) tmem_ld32(taddr, av);
    else tmem_ld16(taddr, av);
    tmem_ld_wait();
    const uint32_t my_row = stg + (uint32_t)lane * 128u;
#pragma unroll
    for (int j = 0; j < GW / 4; ++j) {
        float f[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) f[e] = row_ok ? __uint_as_float(av[4 * j + e]) * oscale : 0.f;   // invalid rows: exact zeros
        st_shared_f4(my_row + (uint32_t)((j ^ (lane & 7)) << 4), f[0], f[1], f[2], f[3]);
    }
    __syncwarp();
    if (stat != nullptr && bwd == nullptr && lane < GW) {
        float s1 = 0.f, s2 = 0.f;
        const uint32_t word = (uint32_t)(lane & 3) * 4u;
        const int ch = lane >> 2;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
            const float v = ld_shared_f1(stg + (uint32_t)r * 128u + (uint32_t)((ch ^ (r & 7)) << 4) + word);
            s1 += v;
            s2 = fmaf(v, v, s2);
        }
        stat[lane] = s1;
        stat[256 + lane] = s2;
    }
    float4 b_sc = make_float4(0.f, 0.f, 0.f, 0.f), b_sh = b_sc, b_mu = b_sc, b_is = b_sc;
    float4 t1 = b_sc, t2 = b_sc;     // sum g*m, sum g*m*zhat of this lane's 4 columns over its rows
    if (want_bwd && col < cd) {
        b_sc = __ldg(reinterpret_cast<const float4*>(bwd->scale + col));
        b_sh = __ldg(reinterpret_cast<const float4*>(bwd->shift + col));
        b_mu = __ldg(reinterpret_cast<const float4*>(bwd->mean + col));
        b_is = __ldg(reinterpret_cast<const float4*>(bwd->invstd + col));
    }
#pragma unroll
    for (int i = 0; i < CPR; ++i) {   // 32 / RPI == CPR iterations
        const int rr = i * RPI + rsub;
        const int pix = row_pix[rr];
        float4 v = ld_shared_f4(stg + (uint32_t)rr * 128u + (uint32_t)((j ^ (rr & 7)) << 4));
        if (pix >= 0 && col < cd) {
            if (want_bwd) {
                const float4 q = zz[i];
                const float gx = fmaf(q.x, b_sc.x, b_sh.x) > 0.f ? v.x : 0.f, gy = fmaf(q.y, b_sc.y, b_sh.y) > 0.f ? v.y : 0.f;
                const float gz = fmaf(q.z, b_sc.z, b_sh.z) > 0.f ? v.z : 0.f, gw = fmaf(q.w, b_sc.w, b_sh.w) > 0.f ? v.w : 0.f;
                t1.x += gx; t1.y += gy; t1.z += gz; t1.w += gw;
                t2.x = fmaf(gx, (q.x - b_mu.x) * b_is.x, t2.x);
                t2.y = fmaf(gy, (q.y - b_mu.y) * b_is.y, t2.y);
                t2.z = fmaf(gz, (q.z - b_mu.z) * b_is.z, t2.z);
                t2.w = fmaf(gw, (q.w - b_mu.w) * b_is.w, t2.w);
            }
            if (do_store) {
                float4* dp = reinterpret_cast<float4*>(dst + (size_t)pix * cd + col);
                if (accumulate) {
                    const float4 old = *dp;
                    v.x += old.x;
                    v.y += old.y;
                    v.z += old.z;
                    v.w += old.w;
                }
                *dp = v;
            }
        }
    }
    if (want_bwd) {   // lanes j, j+CPR, j+2*CPR, ... hold the same columns: fold them (fixed order), lane j publishes
#pragma unroll
        for (int o = CPR; o < 32; o <<= 1) {
            t1.x += __shfl_xor_sync(0xffffffffu, t1.x, o); t1.y += __shfl_xor_sync(0xffffffffu, t1.y, o);
            t1.z += __shfl_xor_sync(0xffffffffu, t1.z, o); t1.w += __shfl_xor_sync(0xffffffffu, t1.w, o);
            t2.x += __shfl_xor_sync(0xffffffffu, t2.x, o); t2.y += __shfl_xor_sync(0xffffffffu, t2.y, o);
            t2.z += __shfl_xor_sync(0xffffffffu, t2.z, o); t2.w += __shfl_xor_sync(0xffffffffu, t2.w, o);
        }
        if (lane < CPR) {
            stat[j * 4 + 0] = t1.x; stat[j * 4 + 1] = t1.y; stat[j * 4 + 2] = t1.z; stat[j * 4 + 3] = t1.w;
            stat[256 + j * 4 + 0] = t2.x; stat[256 + j * 4 + 1] = t2.y; stat[256 + j * 4 + 2] = t2.z; stat[256 + j * 4 + 3] = t2.w;
        }
    }
    __syncwarp();
}

// all bnt (multiple of 16) columns of one accumulator; stat = this warp's [2][256] scratch or nullptr
__device__ __forceinline__ void epi_drain_tile(uint32_t taddr, int bnt, float oscale, bool row_ok, uint32_t stg, const int* row_pix,
                                               float* __restrict__ dst, int cd, int n_base, int accumulate, bool do_store,
                                               float* stat, int lane, const EpiBwdStat* bwd = nullptr) {
    int c0 = 0;
    for (; c0 + 32 <= bnt; c0 += 32)
        epi_drain_group<32>(taddr + (uint32_t)c0, oscale, row_ok, stg, row_pix, dst, cd, n_base + c0, accumulate, do_store,
                            stat ? stat + c0 : nullptr, lane, bwd);
    if (c0 < bnt)
        epi_drain_group<16>(taddr + (uint32_t)c0, oscale, row_ok, stg, row_pix, dst, cd, n_base + c0, accumulate, do_store,
                            stat ? stat + c0 : nullptr, lane, bwd);
}

}  // namespace sv
