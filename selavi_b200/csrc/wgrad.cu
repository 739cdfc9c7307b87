// Weight-gradient of the implicit-GEMM convolution on tcgen05 (the contraction runs over PIXELS).
//
//   dW[co, ci, tap] = sum_m  P(src)[pix(m, tap), ci] * dz[m, co]          (m = output pixel of the forward conv)
//
// Replaces cuDNN's wgrad behind every Conv3d/Conv2d of the reference model (model.py:93-121) in
// loss.backward() (main.py:298).  GEMM view: rows = flattened (tap, ci) (128 per CTA), columns = co (BNt),
// K = pixels.  Both operands are channels-last in HBM, i.e. MN-major for this GEMM.  tcgen05.mma.kind::tf32
// returned all-zero accumulators for every MN-major descriptor variant probed on B200 (tools/umma_probe.py,
// profiles/r01_umma_probe.txt), so the loaders transpose 4 pixels x 4 channels in registers and store K-major
// SWIZZLE_128B tiles (a 128-byte row = 32 consecutive pixels of one channel).  The im2col gather, the fused
// BN+ReLU prologue of the previous layer and the tf32 hi/lo split are done in registers like in conv.cu.  The pixel range is split
// over CTAs (split-K); partial tiles go to a [slices][taps*cs][co] buffer that wgrad_reduce sums in a fixed
// order and scatters into the torch weight layout (deterministic, no atomics).
#include <cuda_bf16.h>
#include <cudaTypedefs.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/selavi_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int WG_LOADER_WARPS = 8;
constexpr int WG_MMA_WARP = WG_LOADER_WARPS;
constexpr int WG_THREADS = (WG_LOADER_WARPS + 1) * 32;
constexpr int WG_PIX = 32;                     // pixels per stage (4 MMA K steps)
constexpr int WG_A_BYTES = 128 * 128;          // 128 GEMM rows x 128 B (32 pixels), K-major SWIZZLE_128B

// Split-K schedule: ONE wave of CTAs (148 / (groups x column tiles) slices per group) whenever the layer has that much
// parallelism along the pixels, and the SAME slices for every group of row tiles: the groups then walk the same pixels at
// the same time and the dz stages they all need are shared through L2.  Measured on B200 (tools/wgrad_sweep.py,
// profiles/r02_wgrad_sweep.txt): against round 1's three waves this is 4-15 % faster on every layer (a third of the
// partial-tile traffic, a third of the pipeline fills); dealing slices in proportion to a group's row tiles (equal MMA
// work per CTA, but the groups drift apart) is SLOWER than equal slices by up to 25 %.
constexpr int WG_MAX_GROUPS = 96;
struct WgSplit {
    int ngroups;
    int interleaved;                 // 1: every group has the same slices and CTA ids are slice-major (groups adjacent)
    int ctas_per_ntile;              // sum of slices[g]
    int max_slices;
    short slices[WG_MAX_GROUPS];     // split-K slices of group g
    int kps[WG_MAX_GROUPS];          // pixel stages per slice of group g
    short cta0[WG_MAX_GROUPS + 1];   // first CTA (within a column tile) of group g
};

__device__ __forceinline__ void wg_locate(const WgSplit& sp, int cid, int total_kstages, int& group, int& ntile, int& slice,
                                          int& ks_begin, int& ks_end) {
    ntile = cid / sp.ctas_per_ntile;
    const int r = cid - ntile * sp.ctas_per_ntile;
    int g = 0;
    if (sp.interleaved) {
        // equal slices per group: consecutive CTA ids = the groups of ONE slice, so the CTAs that stream the same dz pixels
        // sit on neighbouring SMs (ncu: with group-major ids every group fetched dz from DRAM separately, 2.32 GB against
        // 1.34 GB algorithmic on layer 1 — the two halves of the wave live on different dies / L2 partitions)
        g = r % sp.ngroups;
        slice = r / sp.ngroups;
    } else {
        while (g + 1 < sp.ngroups && r >= sp.cta0[g + 1]) ++g;
        slice = r - sp.cta0[g];
    }
    group = g;
    ks_begin = slice * sp.kps[g];
    ks_end = ks_begin + sp.kps[g];
    if (ks_end > total_kstages) ks_end = total_kstages;
}

struct WgradParams {
    const float* src;   // forward input (raw), gathered
    const float* dz;    // gradient wrt the conv output [M, cd]
    float* partial;     // [slices][mtiles*128][ntiles*bnt]
    const float* pro_scale;
    const float* pro_shift;
    int nb, ts, hs, ws, cs;
    int td, hd, wd, cd;
    int kt, kh, kw, st, sh, sw, pt, ph, pw;
    int M;
    int mtiles, bnt, ntiles, natom;   // natom = ceil(bnt/32)
    int stages, total_kstages;
    int pro_relu, passes;
    uint32_t tmem_cols;
    WgSplit split;
};

__device__ __forceinline__ uint32_t wg_hi(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }
__device__ __forceinline__ void wg_st4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// 4x4 register transpose: in r[i] = (channel c+0..3) of pixel i  ->  out r[i] = (pixel 0..3) of channel c+i
__device__ __forceinline__ void transpose4(float4 (&r)[4]) {
    const float4 a = r[0], b = r[1], c = r[2], d = r[3];
    r[0] = make_float4(a.x, b.x, c.x, d.x);
    r[1] = make_float4(a.y, b.y, c.y, d.y);
    r[2] = make_float4(a.z, b.z, c.z, d.z);
    r[3] = make_float4(a.w, b.w, c.w, d.w);
}

// store 4 GEMM rows (row0..row0+3), 16-byte chunk `pg` (4 consecutive pixels) of a K-major SW128 tile, hi (+lo)
__device__ __forceinline__ void store_rows4(uint32_t hi_base, uint32_t lo_base, int row0, int pg, const float4 (&r)[4],
                                            bool with_lo) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = row0 + i;
        const uint32_t off = (uint32_t)(row * 128 + ((pg ^ (row & 7)) << 4));
        const uint32_t h0 = wg_hi(r[i].x), h1 = wg_hi(r[i].y), h2 = wg_hi(r[i].z), h3 = wg_hi(r[i].w);
        wg_st4(hi_base + off, h0, h1, h2, h3);
        if (with_lo)
            wg_st4(lo_base + off, __float_as_uint(r[i].x - __uint_as_float(h0)), __float_as_uint(r[i].y - __uint_as_float(h1)),
                   __float_as_uint(r[i].z - __uint_as_float(h2)), __float_as_uint(r[i].w - __uint_as_float(h3)));
    }
}

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_kernel(const WgradParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // stage: A_hi [128 rows x 128B] | A_lo | B_hi [bnt rows x 128B] | B_lo ; a 128B row = 32 pixels (K-major)
    const int b_bytes = p.bnt * 128;
    const int stage_bytes = 2 * WG_A_BYTES + 2 * b_bytes;
    unsigned char* tail = smem + (size_t)p.stages * stage_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
    uint64_t* empty_bar = full_bar + 8;
    uint64_t* accum_bar = empty_bar + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int mt, ntile, slice, ks_begin, ks_end;
    wg_locate(p.split, (int)blockIdx.x, p.total_kstages, mt, ntile, slice, ks_begin, ks_end);   // groups of one row tile
    const int nks = ks_end - ks_begin;   // >= 1 by construction

    if (tid == 0) {
        for (int s = 0; s < p.stages; ++s) {
            sv::mbar_init(&full_bar[s], WG_LOADER_WARPS);
            sv::mbar_init(&empty_bar[s], 1);
        }
        sv::mbar_init(accum_bar, 1);
        sv::fence_barrier_init();
    }
    if (warp == WG_MMA_WARP) {
        sv::tmem_alloc(tmem_slot, p.tmem_cols);
        sv::tmem_relinquish();
    }
    sv::tc_fence_before();
    __syncthreads();
    sv::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int n_off = ntile * p.bnt;

    if (warp < WG_LOADER_WARPS) {
        // thread -> (pixel group pg: 4 consecutive pixels, channel group cg: 4 consecutive GEMM rows)
        const int pg = tid & 7;
        const int cg = tid >> 3;  // 0..31 : A rows 4cg..4cg+3 ; B rows 4cg.. and 4(cg+32)..
        const int C4 = p.cs >> 2;
        const int taps = p.kt * p.kh * p.kw;
        const int Q = mt * 32 + cg;          // flattened K chunk (tap, c4) of this thread's A rows
        const int tap = Q / C4, c4 = Q % C4;
        const bool qvalid = tap < taps;
        const int kw_ = tap % p.kw, kh_ = (tap / p.kw) % p.kh, kt_ = tap / (p.kw * p.kh);
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sf = make_float4(0.f, 0.f, 0.f, 0.f);
        const bool pro = p.pro_scale != nullptr;
        if (pro && qvalid) {
            sc = __ldg(reinterpret_cast<const float4*>(p.pro_scale + c4 * 4));
            sf = __ldg(reinterpret_cast<const float4*>(p.pro_shift + c4 * 4));
        }
        const int nb_blocks = (p.bnt >> 2);   // B channel groups
        int stage = 0;
        uint32_t phase = 0;
        // coordinates of this thread's first pixel of the current stage, advanced incrementally (32 pixels per stage)
        int w0, h0, t0, n0;
        {
            const int mbase = ks_begin * WG_PIX + pg * 4;
            w0 = mbase % p.wd;
            const int t1 = mbase / p.wd;
            h0 = t1 % p.hd;
            const int t2 = t1 / p.hd;
            t0 = t2 % p.td;
            n0 = t2 / p.td;
        }
        int ks_g = ks_begin;   // stage index of the next gather
        const bool hasb0 = cg < nb_blocks && n_off + cg * 4 < p.cd;
        const bool hasb1 = cg + 32 < nb_blocks && n_off + (cg + 32) * 4 < p.cd;
        struct WStage {
            float4 xa[4], xb0[4], xb1[4];
        };
        auto gather = [&](WStage& s) {
            const int mbase = ks_g * WG_PIX + pg * 4;
            int w_ = w0, h_ = h0, t_ = t0, n_ = n0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                s.xa[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                s.xb0[i] = s.xb1[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (mbase + i < p.M) {
                    const int a = t_ * p.st - p.pt + kt_;
                    const int b = h_ * p.sh - p.ph + kh_;
                    const int d = w_ * p.sw - p.pw + kw_;
                    const bool ok = qvalid & (a >= 0) & (a < p.ts) & (b >= 0) & (b < p.hs) & (d >= 0) & (d < p.ws);
                    if (ok) {
                        const size_t pix = (size_t)((n_ * p.ts + a) * p.hs + b) * p.ws + d;
                        float4 x = __ldg(reinterpret_cast<const float4*>(p.src + pix * p.cs + c4 * 4));
                        if (pro) {
                            x.x = fmaf(x.x, sc.x, sf.x);
                            x.y = fmaf(x.y, sc.y, sf.y);
                            x.z = fmaf(x.z, sc.z, sf.z);
                            x.w = fmaf(x.w, sc.w, sf.w);
                            if (p.pro_relu) {
                                x.x = fmaxf(x.x, 0.f);
                                x.y = fmaxf(x.y, 0.f);
                                x.z = fmaxf(x.z, 0.f);
                                x.w = fmaxf(x.w, 0.f);
                            }
                        }
                        s.xa[i] = x;
                    }
                    const float* zrow = p.dz + (size_t)(mbase + i) * p.cd + n_off;
                    if (hasb0) s.xb0[i] = __ldg(reinterpret_cast<const float4*>(zrow + cg * 4));
                    if (hasb1) s.xb1[i] = __ldg(reinterpret_cast<const float4*>(zrow + (cg + 32) * 4));
                    if (++w_ == p.wd) {  // next pixel
                        w_ = 0;
                        if (++h_ == p.hd) {
                            h_ = 0;
                            if (++t_ == p.td) {
                                t_ = 0;
                                ++n_;
                            }
                        }
                    }
                }
            }
            // advance the stage origin by WG_PIX pixels
            ++ks_g;
            w0 += WG_PIX;
            while (w0 >= p.wd) {
                w0 -= p.wd;
                if (++h0 == p.hd) {
                    h0 = 0;
                    if (++t0 == p.td) {
                        t0 = 0;
                        ++n0;
                    }
                }
            }
        };
        auto commit = [&](WStage& s) {
            transpose4(s.xa);
            if (hasb0) transpose4(s.xb0);
            if (hasb1) transpose4(s.xb1);
            sv::mbar_wait(&empty_bar[stage], phase ^ 1);
            const uint32_t a_hi = sv::smem_u32(smem + (size_t)stage * stage_bytes);
            const uint32_t a_lo = a_hi + WG_A_BYTES;
            const uint32_t b_hi = a_lo + WG_A_BYTES;
            const uint32_t b_lo = b_hi + b_bytes;
            const bool with_lo = p.passes == 3;
            store_rows4(a_hi, a_lo, cg * 4, pg, s.xa, with_lo);
            if (cg < nb_blocks) store_rows4(b_hi, b_lo, cg * 4, pg, s.xb0, with_lo);
            if (cg + 32 < nb_blocks) store_rows4(b_hi, b_lo, (cg + 32) * 4, pg, s.xb1, with_lo);
            sv::fence_proxy_async();
            __syncwarp();
            if (lane == 0) sv::mbar_arrive(&full_bar[stage]);
            if (++stage == p.stages) {
                stage = 0;
                phase ^= 1;
            }
        };
        // software pipeline: the loads of stage i+1 are in flight while stage i is transposed, split and stored
        WStage sa, sb;
        gather(sa);
        for (int i = 0; i < nks; i += 2) {
            if (i + 1 < nks) gather(sb);
            commit(sa);
            if (i + 1 < nks) {
                if (i + 2 < nks) gather(sa);
                commit(sb);
            }
        }

        // ---- epilogue: TMEM -> partial[slice][mt*128 + row][ntile*bnt + col]
        sv::mbar_wait(accum_bar, 0);
        sv::tc_fence_after();
        const int quad = warp & 3, half = warp >> 2;
        const int units = p.bnt >> 4;
        const int u_begin = half == 0 ? 0 : (units + 1) / 2;
        const int u_end = half == 0 ? (units + 1) / 2 : units;
        const int row = quad * 32 + lane;
        const int ntot = p.ntiles * p.bnt;
        float* out_row = p.partial + ((size_t)slice * (p.mtiles * 128) + mt * 128 + row) * ntot + n_off;
        for (int u = u_begin; u < u_end; ++u) {
            uint32_t acc[16];
            sv::tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(u * 16), acc);
            sv::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                *reinterpret_cast<float4*>(out_row + u * 16 + i) =
                    make_float4(__uint_as_float(acc[i]), __uint_as_float(acc[i + 1]), __uint_as_float(acc[i + 2]),
                                __uint_as_float(acc[i + 3]));
            }
        }
        sv::tc_fence_before();
    } else {
        if (lane == 0) {
            const uint32_t idesc = sv::make_idesc_tf32(128, p.bnt, 0, 0);  // both operands K-major (K = pixels)
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < nks; ++i) {
                sv::mbar_wait(&full_bar[stage], phase);
                sv::tc_fence_after();
                const uint32_t a_hi = sv::smem_u32(smem + (size_t)stage * stage_bytes);
                const uint32_t a_lo = a_hi + WG_A_BYTES;
                const uint32_t b_hi = a_lo + WG_A_BYTES;
                const uint32_t b_lo = b_hi + b_bytes;
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    const uint64_t da_hi = sv::make_smem_desc_sw128(a_hi + k4 * 32, 16, 1024);
                    const uint64_t db_hi = sv::make_smem_desc_sw128(b_hi + k4 * 32, 16, 1024);
                    if (p.passes == 3) {
                        const uint64_t da_lo = sv::make_smem_desc_sw128(a_lo + k4 * 32, 16, 1024);
                        const uint64_t db_lo = sv::make_smem_desc_sw128(b_lo + k4 * 32, 16, 1024);
                        sv::umma_tf32(tmem_base, da_lo, db_hi, idesc, (i | k4) ? 1u : 0u);
                        sv::umma_tf32(tmem_base, da_hi, db_lo, idesc, 1u);
                        sv::umma_tf32(tmem_base, da_hi, db_hi, idesc, 1u);
                    } else {
                        sv::umma_tf32(tmem_base, da_hi, db_hi, idesc, (i | k4) ? 1u : 0u);
                    }
                }
                sv::umma_commit(&empty_bar[stage]);
                if (++stage == p.stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            sv::umma_commit(accum_bar);
        }
    }
    __syncthreads();
    if (warp == WG_MMA_WARP) {
        sv::tc_fence_after();
        sv::tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------------------------
// bf16x3 variant (default).  Both operands stay MN-major (channels contiguous, exactly as they lie in HBM) —
// supported by tcgen05.mma.kind::f16 (probed: profiles/r01_umma_probe_bf16.txt) but not by kind::tf32.
//   1. split_bf16_kernel (one pass per operand): x = hi + lo, hi = bf16(x), lo = bf16(x - hi); for the conv input it
//      also applies the pending BN+ReLU of the previous layer.  ~2^-17 relative operand error with three MMAs
//      (lo*hi + hi*lo + hi*hi, fp32 accumulate), twice the tf32 MMA rate and half the shared-memory bytes.
//   2. wgrad_bf16_kernel: the loaders are pure 16-byte cp.async copies (8 channels of one pixel, zero-filled outside the
//      image) straight into SWIZZLE_128B MN-major atoms [k-group of 8 pixels][chunk of 64 channels]; no register
//      staging, no conversion, ~6x fewer instructions per stage than the register-staged tf32 loader (ncu:
//      profiles/r01b_wgrad_notes.txt).
__global__ void split_bf16_kernel(const float4* __restrict__ src, const float4* __restrict__ scale,
                                  const float4* __restrict__ shift, int relu, uint2* __restrict__ hi, uint2* __restrict__ lo,
                                  long long total4, int c4n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        float4 x = src[i];
        if (scale) {
            const int c4 = (int)(i % c4n);
            const float4 sc = __ldg(scale + c4), sf = __ldg(shift + c4);
            x.x = fmaf(x.x, sc.x, sf.x);
            x.y = fmaf(x.y, sc.y, sf.y);
            x.z = fmaf(x.z, sc.z, sf.z);
            x.w = fmaf(x.w, sc.w, sf.w);
            if (relu) {
                x.x = fmaxf(x.x, 0.f);
                x.y = fmaxf(x.y, 0.f);
                x.z = fmaxf(x.z, 0.f);
                x.w = fmaxf(x.w, 0.f);
            }
        }
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(x.x, x.y), h1 = __floats2bfloat162_rn(x.z, x.w);
        const __nv_bfloat162 l0 = __floats2bfloat162_rn(x.x - __low2float(h0), x.y - __high2float(h0));
        const __nv_bfloat162 l1 = __floats2bfloat162_rn(x.z - __low2float(h1), x.w - __high2float(h1));
        hi[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
        lo[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
    }
}

struct WgradBf16Params {
    const __nv_bfloat16* a_hi;  // conv input activation, normalised + split  [pixels_in][cs]
    const __nv_bfloat16* a_lo;
    const __nv_bfloat16* z_hi;  // gradient wrt the conv output, split         [M][cd]
    const __nv_bfloat16* z_lo;
    float* partial;
    int nb, ts, hs, ws, cs;
    int td, hd, wd, cd;
    int kt, kh, kw, st, sh, sw, pt, ph, pw;
    int M;
    int mtiles, bnt, ntiles;
    int G, groups;   // G consecutive 128-row tiles per CTA share every dz stage (one accumulator each)
    int stages, total_kstages;
    int passes;
    int use_tma;     // 1: the column operand (dz planes) is fetched with cp.async.bulk.tensor (tensor maps tm_hi / tm_lo)
    uint32_t tmem_cols;
    WgSplit split;
};

constexpr int WB_MAX_G = 3;
constexpr int WB_LOADER_WARPS = 8;
constexpr int WB_MMA_WARP = WB_LOADER_WARPS;
constexpr int WB_THREADS = (WB_LOADER_WARPS + 1) * 32;
constexpr int WB_LTHREADS = WB_LOADER_WARPS * 32;
constexpr int WB_A_BYTES = 4 * 2 * 1024;   // 4 k-groups x 2 chunks of 64 channels (bf16)

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// The column operand B (dz planes [M pixels][cd channels], MN-major) needs no per-element work: each 64-channel chunk of a
// 32-pixel stage (4 SWIZZLE_128B atoms) is one TMA tiled box {64, 32} of a 2-D tensor map, written with the hardware's 128-byte swizzle
// (identical to the (chunk ^ pixel & 7) pattern the cp.async path builds by hand), zero-filled beyond the last pixel /
// channel, completing on the stage's mbarrier.  One thread issues them; the 256 loader threads keep the row operand
// (tap-shifted gathers with per-pixel bounds, which a tiled box cannot express for the strided / padded cases).
__global__ void __launch_bounds__(WB_THREADS, 1) wgrad_bf16_kernel(const WgradBf16Params p, const __grid_constant__ CUtensorMap tm_hi,
                                                                   const __grid_constant__ CUtensorMap tm_lo) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int nchb = (p.bnt + 63) >> 6;              // 64-channel chunks of the B tile
    const int b_bytes = 4 * nchb * 1024;
    const int stage_bytes = p.G * 2 * WB_A_BYTES + 2 * b_bytes;   // G x (A_hi | A_lo) | B_hi | B_lo
    unsigned char* tail = smem + (size_t)p.stages * stage_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
    uint64_t* empty_bar = full_bar + 8;
    uint64_t* accum_bar = empty_bar + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int group, ntile, slice, ks_begin, ks_end;
    wg_locate(p.split, (int)blockIdx.x, p.total_kstages, group, ntile, slice, ks_begin, ks_end);
    const int mt0 = group * p.G;
    const int nmt = p.mtiles - mt0 < p.G ? p.mtiles - mt0 : p.G;   // 128-row tiles of this CTA
    const int nks = ks_end - ks_begin;

    if (tid == 0) {
        for (int s = 0; s < p.stages; ++s) {
            sv::mbar_init(&full_bar[s], WB_LOADER_WARPS + (p.use_tma ? 1 : 0));   // + the TMA issuer's arrive.expect_tx
            sv::mbar_init(&empty_bar[s], 1);
        }
        sv::mbar_init(accum_bar, 1);
        sv::fence_barrier_init();
        if (p.use_tma) {
            sv::tma_prefetch_desc(&tm_hi);
            sv::tma_prefetch_desc(&tm_lo);
        }
    }
    if (warp == WB_MMA_WARP) {
        sv::tmem_alloc(tmem_slot, p.tmem_cols);
        sv::tmem_relinquish();
    }
    sv::tc_fence_before();
    __syncthreads();
    sv::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int n_off = ntile * p.bnt;

    if (warp < WB_LOADER_WARPS) {
        const int C8 = p.cs >> 3;
        const int taps = p.kt * p.kh * p.kw;
        // ---- A units (pixel, 8 GEMM rows): u = tid + 256*j, px = (tid >> 4) + 16*j, g8 = tid & 15; the 8 rows are the
        //      flattened K chunk (tap, c8) — fixed per thread for the whole kernel (cs is a multiple of 8)
        const int g8a = tid & 15;
        int a_coff[WB_MAX_G], a_k[WB_MAX_G];   // per tile of the group: channel offset, packed tap (kt | kh<<8 | kw<<16) or -1
#pragma unroll
        for (int g = 0; g < WB_MAX_G; ++g) {
            const int Q8 = (mt0 + g) * 16 + g8a;
            const int tap = Q8 / C8;
            a_coff[g] = (Q8 % C8) * 8;
            a_k[g] = (g < nmt && tap < taps) ? ((tap / (p.kw * p.kh)) | (((tap / p.kw) % p.kh) << 8) | ((tap % p.kw) << 16)) : -1;
        }
        uint32_t a_soff[2];
        int cw[2], ch_[2], ct[2], cn[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int px = (tid >> 4) + 16 * j;
            a_soff[j] = (uint32_t)(((px >> 3) * 2 + (g8a >> 3)) * 1024 + (px & 7) * 128 + (((g8a & 7) ^ (px & 7)) << 4));
            const int m = ks_begin * WG_PIX + px;
            cw[j] = m % p.wd;
            const int t1 = m / p.wd;
            ch_[j] = t1 % p.hd;
            const int t2 = t1 / p.hd;
            ct[j] = t2 % p.td;
            cn[j] = t2 / p.td;
        }
        // ---- B units (pixel, 8 channels of dz): u = tid + 256*j, px = u / upp, g8 = u % upp
        const int upp = (p.bnt + 7) >> 3;
        int b_px[4];
        uint32_t b_soff[4];
        long long b_goff[4];   // element offset of the unit inside a stage: px*cd + n_off + g8*8 (or -1)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int u = tid + WB_LTHREADS * j;
            const int px = u / upp, g8 = u % upp;
            const int c0 = n_off + g8 * 8;
            const bool ok = px < WG_PIX && c0 < p.cd;
            b_px[j] = ok ? px : -1;
            // B tile layout [64-channel chunk][k-group of 8 pixels][8 rows x 128 B]: the 4 k-groups of a chunk are contiguous, so
            // that ONE TMA box of 64 channels x 32 pixels fills a chunk (descriptor: LBO = 4 KB between chunks, SBO = 1 KB)
            b_soff[j] = (uint32_t)(((g8 >> 3) * 4 + (px >> 3)) * 1024 + (px & 7) * 128 + (((g8 & 7) ^ (px & 7)) << 4));
            b_goff[j] = (long long)px * p.cd + c0;
        }
        int stage = 0;
        uint32_t phase = 0;
        int prev_stage = -1;
        const bool with_lo = p.passes == 3;
        for (int i = 0; i < nks; ++i) {
            const int ks = ks_begin + i;
            sv::mbar_wait(&empty_bar[stage], phase ^ 1);
            const uint32_t a_base = sv::smem_u32(smem + (size_t)stage * stage_bytes);
            const uint32_t b_hi = a_base + (uint32_t)(p.G * 2 * WB_A_BYTES);
            const uint32_t b_lo = b_hi + b_bytes;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int m = ks * WG_PIX + (tid >> 4) + 16 * j;
                const int a0 = ct[j] * p.st - p.pt, b0 = ch_[j] * p.sh - p.ph, d0 = cw[j] * p.sw - p.pw;
#pragma unroll
                for (int g = 0; g < WB_MAX_G; ++g) {
                    if (g < nmt) {
                        const int a = a0 + (a_k[g] & 0xff), b = b0 + ((a_k[g] >> 8) & 0xff), d = d0 + ((a_k[g] >> 16) & 0xff);
                        const bool ok = (a_k[g] >= 0) & (m < p.M) & (a >= 0) & (a < p.ts) & (b >= 0) & (b < p.hs) & (d >= 0) & (d < p.ws);
                        const size_t off = ok ? ((size_t)(((cn[j] * p.ts + a) * p.hs + b) * p.ws + d) * p.cs + a_coff[g]) : 0;
                        const uint32_t nbytes = ok ? 16u : 0u;
                        const uint32_t a_hi = a_base + (uint32_t)(g * 2 * WB_A_BYTES);
                        cp_async16(a_hi + a_soff[j], p.a_hi + off, nbytes);
                        if (with_lo) cp_async16(a_hi + WB_A_BYTES + a_soff[j], p.a_lo + off, nbytes);
                    }
                }
                cw[j] += WG_PIX;  // advance this pixel by one stage
                while (cw[j] >= p.wd) {
                    cw[j] -= p.wd;
                    if (++ch_[j] == p.hd) {
                        ch_[j] = 0;
                        if (++ct[j] == p.td) {
                            ct[j] = 0;
                            ++cn[j];
                        }
                    }
                }
            }
            if (p.use_tma) {
                if (tid == 0) {   // nchb channel chunks x (hi, lo): one 64-channel x 32-pixel box (4 SWIZZLE_128B atoms) each
                    sv::mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(with_lo ? 2 * b_bytes : b_bytes));
                    for (int c = 0; c < nchb; ++c) {
                        const uint32_t off = (uint32_t)(c * 4096);
                        sv::tma_load_2d(smem + (size_t)stage * stage_bytes + (size_t)(p.G * 2 * WB_A_BYTES) + off, &tm_hi,
                                        &full_bar[stage], n_off + c * 64, ks * WG_PIX);
                        if (with_lo)
                            sv::tma_load_2d(smem + (size_t)stage * stage_bytes + (size_t)(p.G * 2 * WB_A_BYTES) + b_bytes + off,
                                            &tm_lo, &full_bar[stage], n_off + c * 64, ks * WG_PIX);
                    }
                }
            }
            const long long zbase = (long long)ks * WG_PIX * p.cd;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (!p.use_tma && b_px[j] >= 0) {
                    const bool ok = ks * WG_PIX + b_px[j] < p.M;
                    const size_t off = ok ? (size_t)(zbase + b_goff[j]) : 0;
                    const uint32_t nbytes = ok ? 16u : 0u;
                    cp_async16(b_hi + b_soff[j], p.z_hi + off, nbytes);
                    if (with_lo) cp_async16(b_lo + b_soff[j], p.z_lo + off, nbytes);
                }
            }
            cp_async_commit();
            if (prev_stage >= 0) {   // the previous stage's copies have landed: publish it to the MMA warp
                cp_async_wait<1>();
                sv::fence_proxy_async();
                __syncwarp();
                if (lane == 0) sv::mbar_arrive(&full_bar[prev_stage]);
            }
            prev_stage = stage;
            if (++stage == p.stages) {
                stage = 0;
                phase ^= 1;
            }
        }
        cp_async_wait<0>();
        sv::fence_proxy_async();
        __syncwarp();
        if (lane == 0 && prev_stage >= 0) sv::mbar_arrive(&full_bar[prev_stage]);

        // ---- epilogue: TMEM -> partial[slice][mt*128 + row][ntile*bnt + col]
        sv::mbar_wait(accum_bar, 0);
        sv::tc_fence_after();
        const int quad = warp & 3, half = warp >> 2;
        const int units = p.bnt >> 4;
        const int u_begin = half == 0 ? 0 : (units + 1) / 2;
        const int u_end = half == 0 ? (units + 1) / 2 : units;
        const int row = quad * 32 + lane;
        const int ntot = p.ntiles * p.bnt;
        for (int g = 0; g < nmt; ++g) {
            float* out_row = p.partial + ((size_t)slice * (p.mtiles * 128) + (mt0 + g) * 128 + row) * ntot + n_off;
            for (int u = u_begin; u < u_end; ++u) {
                uint32_t acc[16];
                sv::tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(g * p.bnt + u * 16), acc);
                sv::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    *reinterpret_cast<float4*>(out_row + u * 16 + i) =
                        make_float4(__uint_as_float(acc[i]), __uint_as_float(acc[i + 1]), __uint_as_float(acc[i + 2]),
                                    __uint_as_float(acc[i + 3]));
                }
            }
        }
        sv::tc_fence_before();
    } else {
        // MMA issuer: converged warp, warp-uniform schedule, one elected lane issues (see conv.cu)
        {
            const uint32_t idesc = sv::make_idesc_f16(128, p.bnt, 1, 1, 1, 1);  // bf16 x bf16, both MN-major
            const uint32_t tm0 = __shfl_sync(0xffffffffu, tmem_base, 0);
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t a_sbo = 2 * 1024, b_sbo = 1024;   // byte stride between k-groups of 8 pixels (A: [k-group][chunk], B: [chunk][k-group])
            const uint64_t a_fixed = sv::make_smem_desc(0, 1024, a_sbo, 2), b_fixed = sv::make_smem_desc(0, 4096, b_sbo, 2);
            for (int i = 0; i < nks; ++i) {
                sv::mbar_wait(&full_bar[stage], phase);
                sv::tc_fence_after();
                const uint32_t a_base = sv::smem_u32(smem + (size_t)stage * stage_bytes);
                // UMMA_K = 16 pixels = 2 k-groups of 8: step k16 advances the address field by 2 * sbo bytes
                const uint64_t da0 = a_fixed | (uint64_t)((a_base & 0x3FFFFu) >> 4);
                const uint64_t db_hi = (b_fixed | (uint64_t)((a_base & 0x3FFFFu) >> 4)) + (uint64_t)((p.G * 2 * WB_A_BYTES) >> 4);
                const uint64_t db_lo = db_hi + (uint64_t)(b_bytes >> 4);
                const uint64_t a_step = (uint64_t)((2 * a_sbo) >> 4), b_step = (uint64_t)((2 * b_sbo) >> 4);
                if (sv::elect_one()) {
                    for (int g = 0; g < nmt; ++g) {
                        const uint64_t da_hi = da0 + (uint64_t)((g * 2 * WB_A_BYTES) >> 4);
                        const uint64_t da_lo = da_hi + (uint64_t)(WB_A_BYTES >> 4);
                        const uint32_t d_tmem = tm0 + (uint32_t)(g * p.bnt);
                        if (p.passes == 3) {
#pragma unroll
                            for (int k16 = 0; k16 < 2; ++k16) {
                                sv::umma_f16(d_tmem, da_lo + k16 * a_step, db_hi + k16 * b_step, idesc, (uint32_t)(i | k16));
                                sv::umma_f16(d_tmem, da_hi + k16 * a_step, db_lo + k16 * b_step, idesc, 1u);
                                sv::umma_f16(d_tmem, da_hi + k16 * a_step, db_hi + k16 * b_step, idesc, 1u);
                            }
                        } else {
#pragma unroll
                            for (int k16 = 0; k16 < 2; ++k16)
                                sv::umma_f16(d_tmem, da_hi + k16 * a_step, db_hi + k16 * b_step, idesc, (uint32_t)(i | k16));
                        }
                    }
                    sv::umma_commit(&empty_bar[stage]);
                }
                __syncwarp();
                if (++stage == p.stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            if (sv::elect_one()) sv::umma_commit(accum_bar);
            __syncwarp();
        }
    }
    __syncthreads();
    if (warp == WB_MMA_WARP) {
        sv::tc_fence_after();
        sv::tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// dW[co][ci][tap] (+)= sum_s partial[s][tap*cs + ci][co]
// swapped (operands exchanged, see wgrad_run): partial[s][(taps-1-tap)*rs + co][ci], rs = channel stride of dz
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, const WgSplit sp, int G, int mg_pad, int ntot, int co, int ci,
                                    int taps, int rs, int swapped, float* __restrict__ dW, int accumulate) {
    const size_t total = (size_t)co * ci * taps;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        // the column index of the partial tiles runs fastest (coalesced reads)
        int o, c, tap;
        size_t row;
        int col;
        if (!swapped) {
            o = (int)(idx % co);
            const size_t r = idx / co;
            c = (int)(r % ci);
            tap = (int)(r / ci);
            row = (size_t)tap * rs + c;
            col = o;
        } else {
            c = (int)(idx % ci);
            const size_t r = idx / ci;
            o = (int)(r % co);
            tap = (int)(r / co);
            row = (size_t)(taps - 1 - tap) * rs + o;
            col = c;
        }
        // four independent partial chains (fixed association: still bit-reproducible) keep 4 loads in flight per thread;
        // with up to ~450 split-K slices a single dependent chain made this kernel latency bound
        const float* src = partial + row * ntot + col;
        const size_t kstride = (size_t)mg_pad * ntot;
        const int slices = sp.slices[(int)(row >> 7) / G];   // the row tile's group decides how many slices exist
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int k = 0;
        for (; k + 4 <= slices; k += 4) {
            s0 += src[(size_t)k * kstride];
            s1 += src[(size_t)(k + 1) * kstride];
            s2 += src[(size_t)(k + 2) * kstride];
            s3 += src[(size_t)(k + 3) * kstride];
        }
        for (; k < slices; ++k) s0 += src[(size_t)k * kstride];
        const float s = (s0 + s1) + (s2 + s3);
        float* dst = dW + ((size_t)o * ci + c) * taps + tap;
        *dst = accumulate ? (*dst + s) : s;
    }
}

void wg_tiles(int n_out, int* bnt, int* ntiles) {
    int nt = (n_out + 255) / 256;
    int per = (n_out + nt - 1) / nt;
    per = (per + 15) & ~15;
    *bnt = per;
    *ntiles = nt;
}

struct WgPlan {
    int mtiles, bnt, ntiles, natom, total_kstages;
    int G, groups;   // bf16 kernel: G consecutive 128-row tiles per CTA (tf32 kernel: always 1 tile per CTA)
    WgSplit split;
};

WgPlan wg_plan(int co, int taps, int cs, long long M, bool bf16) {
    WgPlan pl;
    wg_tiles(co, &pl.bnt, &pl.ntiles);
    pl.natom = (pl.bnt + 31) / 32;
    pl.mtiles = (taps * cs + 127) / 128;
    pl.total_kstages = (int)((M + WG_PIX - 1) / WG_PIX);
    // bf16 kernel: the CTAs of a group of G row tiles share every dz stage (the 4-5 row tiles of the layer-1 convs each
    // re-read the whole gradient tensor from L2 otherwise); G is bounded by TMEM (G * bnt <= 512 columns) and by keeping
    // at least 3 pipeline stages of G * 16 KB + the dz tile in shared memory
    int G = 1;
    if (bf16) {
        const int b_stage = 2 * 4 * ((pl.bnt + 63) / 64) * 1024;
        int gmax = 512 / pl.bnt;
        const int gsmem = ((226 * 1024 - 256) / 3 - b_stage) / (2 * WB_A_BYTES);
        if (gmax > gsmem) gmax = gsmem;
        if (gmax > WB_MAX_G) gmax = WB_MAX_G;
        if (gmax < 1) gmax = 1;
        const int groups = (pl.mtiles + gmax - 1) / gmax;
        G = (pl.mtiles + groups - 1) / groups;   // balanced groups
    }
    pl.G = G;
    pl.groups = (pl.mtiles + G - 1) / G;
    // split-K: one wave of CTAs, equal slices per group (see WgSplit); a layer with more (group, column tile) pairs than
    // SMs runs one slice each
    WgSplit& sp = pl.split;
    sp.ngroups = pl.groups;
    sp.ctas_per_ntile = sp.max_slices = 0;
    sp.interleaved = 0;
    if (pl.groups > WG_MAX_GROUPS) return pl;   // rejected by wgrad_run
    // experiment knobs (tools/wgrad_sweep.py): SELAVI_WGRAD_WAVES = waves of CTAs (default 1), SELAVI_WGRAD_ALIGNED = 1 gives
    // every group the same slices (the groups then walk the same pixels at the same time and share dz through L2)
    static const int waves = getenv("SELAVI_WGRAD_WAVES") ? atoi(getenv("SELAVI_WGRAD_WAVES")) : 1;
    static const int aligned = getenv("SELAVI_WGRAD_ALIGNED") ? atoi(getenv("SELAVI_WGRAD_ALIGNED")) : 1;
    const int budget = (148 * (waves > 0 ? waves : 1)) / pl.ntiles > 0 ? (148 * (waves > 0 ? waves : 1)) / pl.ntiles : 1;
    int used = 0, max_s = 1;
    sp.cta0[0] = 0;
    for (int g = 0; g < pl.groups; ++g) {
        const int nmt = pl.mtiles - g * G < G ? pl.mtiles - g * G : G;
        // largest-remainder style: what is left of the budget, spread over the row tiles that are left
        const int tiles_left = pl.mtiles - g * G;
        int sl = (int)(((long long)(budget - used) * nmt + tiles_left / 2) / tiles_left);
        if (aligned) sl = budget / pl.groups;
        if (sl < 1) sl = 1;
        if (sl > pl.total_kstages) sl = pl.total_kstages;
        if (sl > 1024) sl = 1024;
        const int kps = (pl.total_kstages + sl - 1) / sl;
        sl = (pl.total_kstages + kps - 1) / kps;     // no empty slice
        sp.slices[g] = (short)sl;
        sp.kps[g] = kps;
        used += sl;
        sp.cta0[g + 1] = (short)used;
        if (sl > max_s) max_s = sl;
    }
    sp.ctas_per_ntile = used;
    sp.max_slices = max_s;
    sp.interleaved = 1;
    for (int g = 1; g < pl.groups; ++g)
        if (sp.slices[g] != sp.slices[0]) sp.interleaved = 0;
    if (getenv("SELAVI_WGRAD_GROUP_MAJOR")) sp.interleaved = 0;
    return pl;
}

// Tiling + operand orientation of one weight gradient (host logic, also exported as selavi_conv_wgrad_plan).
// Exchanged operands (bf16 kernel, stride-1 "same" convolutions): sum_px dz[px][co] x[px+tap][ci] = sum_q
// dz[q-tap][co] x[q][ci], i.e. the same kernel with dz as the tap-shifted row operand (rows = (flipped tap, co)) and
// x as the column operand (N = ci).  Pays when that gives fewer / better-shaped MMAs: one M=128 MMA costs about
// max(N/2 + 12, 59) clocks (tools/umma_rate.cu), so the 144->64 temporal convs of layer 1 run as 2 row tiles x N=144
// instead of 4 row tiles x N=64.
WgPlan wg_choose(const int* geom, int ci_real, bool bf16, bool* swapped) {
    const int ts = geom[2], hs = geom[3], ws = geom[4], cs = geom[5], td = geom[6], hd = geom[7], wd = geom[8], cd = geom[9];
    const int kt = geom[10], kh = geom[11], kw = geom[12], st = geom[13], sh = geom[14], sw = geom[15];
    const int pt = geom[16], ph = geom[17], pw = geom[18], co = geom[19];
    const long long M = (long long)geom[1] * td * hd * wd;
    const int taps = kt * kh * kw;
    WgPlan pl = wg_plan(co, taps, cs, M, bf16);
    *swapped = false;
    if (bf16 && st == 1 && sh == 1 && sw == 1 && ts == td && hs == hd && ws == wd && 2 * pt == kt - 1 && 2 * ph == kh - 1 &&
        2 * pw == kw - 1 && !getenv("SELAVI_WGRAD_NOSWAP")) {
        const WgPlan ps = wg_plan(ci_real, taps, cd, M, true);
        auto cost = [](const WgPlan& w) {
            const double t = w.bnt / 2.0 + 12.0;
            return (double)w.mtiles * w.ntiles * (t < 59.0 ? 59.0 : t);
        };
        if (cost(ps) < 0.9 * cost(pl)) {
            *swapped = true;
            pl = ps;
        }
    }
    return pl;
}

}  // namespace

namespace {
inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
}

// workspace: [split-K partial tiles][a_hi][a_lo][z_hi][z_lo] (the bf16 operand copies are only used by the bf16 path)
extern "C" size_t selavi_wgrad_workspace_bytes(const int* geom) {
    if (!geom) return 0;
    const long long M = (long long)geom[1] * geom[6] * geom[7] * geom[8];
    const long long Min = (long long)geom[1] * geom[2] * geom[3] * geom[4];
    const int taps = geom[10] * geom[11] * geom[12];
    const WgPlan pl = wg_plan(geom[19], taps, geom[5], M, true), pl32 = wg_plan(geom[19], taps, geom[5], M, false);
    const int slices = pl.split.max_slices > pl32.split.max_slices ? pl.split.max_slices : pl32.split.max_slices;
    size_t partial = (size_t)slices * pl.mtiles * 128 * pl.ntiles * pl.bnt * sizeof(float);
    {   // operands exchanged (wgrad_run picks it for some stride-1 convs): rows = taps * cd, columns = input channels
        const WgPlan ps = wg_plan(geom[5], taps, geom[9], M, true);
        const size_t alt = (size_t)ps.split.max_slices * ps.mtiles * 128 * ps.ntiles * ps.bnt * sizeof(float);
        if (alt > partial) partial = alt;
    }
    return align256(partial) + 2 * align256((size_t)Min * geom[5] * 2) + 2 * align256((size_t)M * geom[9] * 2) + 256;
}

namespace {

// 2-D tensor map over a bf16 plane [rows][cols] (cols contiguous): box = 64 columns x 32 rows (the pixels of one stage)
bool make_plane_map(CUtensorMap* tm, const void* base, long long rows, int cols) {
    static PFN_cuTensorMapEncodeTiled encode = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
        else
            cudaGetLastError();
    }
    if (!encode || (reinterpret_cast<uintptr_t>(base) & 15) || (cols & 7)) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    const cuuint32_t box[2] = {64, 32};     // one 64-channel chunk of a pixel stage = 4 consecutive SWIZZLE_128B atoms
    const cuuint32_t estr[2] = {1, 1};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int blocks_for(long long n) {
    long long b = (n + 255) / 256;
    return (int)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b));
}

// Common driver.  z_hi/z_lo: pre-split bf16 dz (bf16 path) or null; dz: fp32 dz (tf32 path, or bf16 path without pre-split).
int wgrad_run(const float* src, const float* dz, const void* z_hi_in, const void* z_lo_in, float* dW, const int* geom,
              int ci_real, const float* pro_scale, const float* pro_shift, int pro_relu, void* workspace, int accumulate,
              int passes, cudaStream_t stream, const void* a_hi_in = nullptr, const void* a_lo_in = nullptr) {
    WgradParams p;
    p.src = src;
    p.dz = dz;
    p.partial = reinterpret_cast<float*>(workspace);
    p.pro_scale = pro_scale;
    p.pro_shift = pro_shift;
    p.nb = geom[1]; p.ts = geom[2]; p.hs = geom[3]; p.ws = geom[4]; p.cs = geom[5];
    p.td = geom[6]; p.hd = geom[7]; p.wd = geom[8]; p.cd = geom[9];
    p.kt = geom[10]; p.kh = geom[11]; p.kw = geom[12];
    p.st = geom[13]; p.sh = geom[14]; p.sw = geom[15];
    p.pt = geom[16]; p.ph = geom[17]; p.pw = geom[18];
    const int co = geom[19];
    if ((p.cs & 7) || (p.cd & 7)) return selavi_fail(-1, "conv_wgrad: channel strides must be multiples of 8");
    // passes: 3 = bf16x3 / 1 = bf16 (MN-major kind::f16 kernel); 13 = tf32x3 / 11 = tf32 (K-major transposing kernel)
    if (passes != 1 && passes != 3 && passes != 11 && passes != 13) return selavi_fail(-1, "conv_wgrad: passes must be 1, 3, 11 or 13");
    if ((pro_scale == nullptr) != (pro_shift == nullptr)) return selavi_fail(-1, "conv_wgrad: prologue needs scale and shift");
    const long long M = (long long)p.nb * p.td * p.hd * p.wd;
    const long long Min = (long long)p.nb * p.ts * p.hs * p.ws;
    if (M <= 0 || M > 0x7fffffffLL || Min > 0x7fffffffLL) return selavi_fail(-1, "conv_wgrad: bad pixel count");
    p.M = (int)M;
    const int taps = p.kt * p.kh * p.kw;
    const bool bf16 = passes < 10;
    bool swapped = false;
    const WgPlan pl = wg_choose(geom, ci_real, bf16, &swapped);
    p.mtiles = pl.mtiles; p.bnt = pl.bnt; p.ntiles = pl.ntiles; p.natom = pl.natom;
    p.total_kstages = pl.total_kstages;
    p.split = pl.split;
    if (pl.groups > WG_MAX_GROUPS) return selavi_fail(-1, "conv_wgrad: too many row-tile groups");
    p.pro_relu = pro_relu;
    if (!bf16 && !dz) return selavi_fail(-1, "conv_wgrad: the tf32 path needs the fp32 gradient");
    p.passes = bf16 ? passes : passes - 10;
    uint32_t cols = 32;
    while ((int)cols < pl.G * p.bnt) cols <<= 1;
    p.tmem_cols = cols;
    const int stage_bytes = bf16 ? pl.G * 2 * WB_A_BYTES + 2 * 4 * ((p.bnt + 63) / 64) * 1024 : 2 * WG_A_BYTES + 2 * p.bnt * 128;
    const int tail_bytes = 8 * 8 * 2 + 8 + 8 + 64;
    int stages = (227 * 1024 - 1024 - tail_bytes) / stage_bytes;
    if (stages > 6) stages = 6;
    if (stages < 2) return selavi_fail(-1, "conv_wgrad: tile does not fit shared memory");
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + tail_bytes + 1024;
    dim3 grid(pl.split.ctas_per_ntile * pl.ntiles);
    if (bf16) {
        // operand preparation: normalise + split the conv input (and dz unless it arrives pre-split): one HBM pass each
        const size_t partial_bytes = align256((size_t)pl.split.max_slices * pl.mtiles * 128 * pl.ntiles * pl.bnt * sizeof(float));
        unsigned char* w = reinterpret_cast<unsigned char*>(workspace) + partial_bytes;
        const size_t a_bytes = align256((size_t)Min * p.cs * 2), z_bytes = align256((size_t)M * p.cd * 2);
        WgradBf16Params q;
        const long long a4 = Min * (p.cs / 4), z4 = M * (p.cd / 4);
        if (a_hi_in) {   // the conv input arrives as bf16 hi/lo planes (emitted by selavi_bn_bwd_apply): no split pass
            q.a_hi = reinterpret_cast<const __nv_bfloat16*>(a_hi_in);
            q.a_lo = reinterpret_cast<const __nv_bfloat16*>(a_lo_in);
        } else {
            q.a_hi = reinterpret_cast<const __nv_bfloat16*>(w);
            q.a_lo = reinterpret_cast<const __nv_bfloat16*>(w + a_bytes);
            split_bf16_kernel<<<blocks_for(a4), 256, 0, stream>>>((const float4*)src, (const float4*)pro_scale, (const float4*)pro_shift,
                                                                  pro_relu, (uint2*)q.a_hi, (uint2*)q.a_lo, a4, p.cs / 4);
        }
        if (z_hi_in) {
            q.z_hi = reinterpret_cast<const __nv_bfloat16*>(z_hi_in);
            q.z_lo = reinterpret_cast<const __nv_bfloat16*>(z_lo_in);
        } else {
            if (!dz) return selavi_fail(-1, "conv_wgrad: no gradient given");
            q.z_hi = reinterpret_cast<const __nv_bfloat16*>(w + 2 * a_bytes);
            q.z_lo = reinterpret_cast<const __nv_bfloat16*>(w + 2 * a_bytes + z_bytes);
            split_bf16_kernel<<<blocks_for(z4), 256, 0, stream>>>((const float4*)dz, nullptr, nullptr, 0, (uint2*)q.z_hi,
                                                                  (uint2*)q.z_lo, z4, p.cd / 4);
        }
        SV_CUDA_CHECK(cudaGetLastError(), "conv_wgrad: split launch");
        q.partial = p.partial;
        q.nb = p.nb; q.ts = p.ts; q.hs = p.hs; q.ws = p.ws; q.cs = p.cs;
        q.td = p.td; q.hd = p.hd; q.wd = p.wd; q.cd = p.cd;
        if (swapped) {   // same geometry (stride 1, same padding), row operand = dz, column operand = x
            const __nv_bfloat16 *xh = q.a_hi, *xl = q.a_lo;
            q.a_hi = q.z_hi; q.a_lo = q.z_lo; q.z_hi = xh; q.z_lo = xl;
            q.cs = p.cd; q.cd = p.cs;
        }
        q.kt = p.kt; q.kh = p.kh; q.kw = p.kw; q.st = p.st; q.sh = p.sh; q.sw = p.sw; q.pt = p.pt; q.ph = p.ph; q.pw = p.pw;
        q.M = p.M; q.mtiles = p.mtiles; q.bnt = p.bnt; q.ntiles = p.ntiles;
        q.G = pl.G; q.groups = pl.groups;
        q.stages = p.stages; q.total_kstages = p.total_kstages; q.split = pl.split;
        q.passes = p.passes; q.tmem_cols = p.tmem_cols;
        // column operand through the TMA engine (SELAVI_WGRAD_TMA=0: the cp.async loaders fetch it as in round 1); its rows
        // are the OUTPUT pixels (or, with exchanged operands on a stride-1 "same" conv, the equally many input pixels)
        alignas(64) CUtensorMap tm_hi, tm_lo;
        memset(&tm_hi, 0, sizeof(tm_hi));
        memset(&tm_lo, 0, sizeof(tm_lo));
        static const bool want_tma = !(getenv("SELAVI_WGRAD_TMA") && atoi(getenv("SELAVI_WGRAD_TMA")) == 0);
        q.use_tma = (want_tma && make_plane_map(&tm_hi, q.z_hi, M, q.cd) && make_plane_map(&tm_lo, q.z_lo, M, q.cd)) ? 1 : 0;
        SV_CUDA_CHECK(cudaFuncSetAttribute(wgrad_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                      "conv_wgrad: cudaFuncSetAttribute");
        wgrad_bf16_kernel<<<grid, WB_THREADS, smem, stream>>>(q, tm_hi, tm_lo);
    } else {
        SV_CUDA_CHECK(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                      "conv_wgrad: cudaFuncSetAttribute");
        wgrad_kernel<<<grid, WG_THREADS, smem, stream>>>(p);
    }
    SV_CUDA_CHECK(cudaGetLastError(), "conv_wgrad: launch");
    const size_t total = (size_t)co * ci_real * taps;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    wgrad_reduce_kernel<<<blocks, 256, 0, stream>>>(p.partial, pl.split, pl.G, pl.mtiles * 128, pl.ntiles * pl.bnt, co, ci_real, taps,
                                                    swapped ? p.cd : p.cs, swapped ? 1 : 0, dW, accumulate);
    SV_CUDA_CHECK(cudaGetLastError(), "conv_wgrad: reduce launch");
    return 0;
}

}  // namespace

// host-only query: tiling of the bf16x3 weight gradient for this geometry (row tiles, column tile width, column tiles,
// row tiles per CTA, the largest number of split-K slices any row-tile group gets, whether the operands are exchanged, and
// the number of CTAs launched)
extern "C" int selavi_conv_wgrad_plan(const int* geom, int ci_real, int* mtiles, int* bnt, int* ntiles, int* tiles_per_cta,
                                      int* slices, int* exchanged, int* ctas) {
    if (!geom || ci_real <= 0) return selavi_fail(-1, "conv_wgrad_plan: bad arguments");
    bool sw = false;
    const WgPlan pl = wg_choose(geom, ci_real, true, &sw);
    if (mtiles) *mtiles = pl.mtiles;
    if (bnt) *bnt = pl.bnt;
    if (ntiles) *ntiles = pl.ntiles;
    if (tiles_per_cta) *tiles_per_cta = pl.G;
    if (slices) *slices = pl.split.max_slices;
    if (ctas) *ctas = pl.split.ctas_per_ntile * pl.ntiles;
    if (exchanged) *exchanged = sw ? 1 : 0;
    return 0;
}

// geom: same 20 ints as selavi_conv_gemm with mode 0 (the FORWARD geometry of the convolution); dz is [M, cd].
extern "C" int selavi_conv_wgrad(const float* src, const float* dz, float* dW, const int* geom, int ci_real,
                                 const float* pro_scale, const float* pro_shift, int pro_relu, void* workspace,
                                 int accumulate, int passes, void* stream) {
    if (!src || !dz || !dW || !geom || !workspace) return selavi_fail(-1, "conv_wgrad: null argument");
    return wgrad_run(src, dz, nullptr, nullptr, dW, geom, ci_real, pro_scale, pro_shift, pro_relu, workspace, accumulate, passes,
                     (cudaStream_t)stream);
}

// same, with the gradient already split into bf16 hi/lo planes [M, cd] (written by selavi_bn_bwd_apply / selavi_split_bf16)
extern "C" int selavi_conv_wgrad_bf16(const float* src, const void* z_hi, const void* z_lo, float* dW, const int* geom,
                                      int ci_real, const float* pro_scale, const float* pro_shift, int pro_relu,
                                      void* workspace, int accumulate, int passes, void* stream) {
    if (!src || !z_hi || !z_lo || !dW || !geom || !workspace) return selavi_fail(-1, "conv_wgrad_bf16: null argument");
    if (passes != 1 && passes != 3) return selavi_fail(-1, "conv_wgrad_bf16: passes must be 1 or 3");
    return wgrad_run(src, nullptr, z_hi, z_lo, dW, geom, ci_real, pro_scale, pro_shift, pro_relu, workspace, accumulate, passes,
                     (cudaStream_t)stream);
}

// same, with BOTH operands already split: a_hi/a_lo = bf16 planes of the conv input activation [pixels_in, cs] (emitted by
// selavi_bn_bwd_apply of the unit that produced it), z_hi/z_lo = planes of the gradient wrt the conv output
extern "C" int selavi_conv_wgrad_bf16_planes(const void* a_hi, const void* a_lo, const void* z_hi, const void* z_lo, float* dW,
                                             const int* geom, int ci_real, void* workspace, int accumulate, int passes, void* stream) {
    if (!a_hi || !a_lo || !z_hi || !z_lo || !dW || !geom || !workspace) return selavi_fail(-1, "conv_wgrad_bf16_planes: null argument");
    if (passes != 1 && passes != 3) return selavi_fail(-1, "conv_wgrad_bf16_planes: passes must be 1 or 3");
    return wgrad_run(nullptr, nullptr, z_hi, z_lo, dW, geom, ci_real, nullptr, nullptr, 0, workspace, accumulate, passes,
                     (cudaStream_t)stream, a_hi, a_lo);
}

// x [M, cs] fp32 -> bf16 hi/lo planes, optional fused affine (+ReLU):  hi = bf16(y), lo = bf16(y - hi), y = act(x*scale+shift)
extern "C" int selavi_split_bf16(const float* x, const float* scale, const float* shift, int relu, void* hi, void* lo,
                                 long long M, int cs, void* stream) {
    if (!x || !hi || !lo || M <= 0 || (cs & 3) || ((scale == nullptr) != (shift == nullptr))) return selavi_fail(-1, "split_bf16: bad arguments");
    const long long n4 = M * (cs / 4);
    split_bf16_kernel<<<blocks_for(n4), 256, 0, (cudaStream_t)stream>>>((const float4*)x, (const float4*)scale, (const float4*)shift,
                                                                       relu, (uint2*)hi, (uint2*)lo, n4, cs / 4);
    SV_CUDA_CHECK(cudaGetLastError(), "split_bf16: launch");
    return 0;
}
