// Data gradient of the implicit-GEMM convolution in bf16x3 on tcgen05 (kind::f16), fed by cp.async.
//
//   dx[m, ci] (+)= sum_{tap, co} dz[pix(m, tap), co] * W[co, ci, tap]        m = input pixel of the forward conv
//
// Replaces cuDNN's dgrad in loss.backward() (main.py:298) for every conv of model.py:93-121.  The gradient dz
// arrives already split into bf16 hi/lo planes (written by the BatchNorm-backward kernel, see elem.cu), so the
// A-operand loaders are pure 16-byte cp.async copies (8 channels of one gathered pixel, zero-filled where the
// transposed gather falls outside the image or between strides) into 128B-swizzled K-major tiles; the weights
// are pre-tiled / pre-split bf16 (dgrad_pack_weights_bf16_kernel) and arrive by one TMA bulk copy per stage.
// Three MMAs per k-step (lo*hi + hi*lo + hi*hi, fp32 accumulate in TMEM): ~2^-17 operand error, measured 4e-6
// per layer — the same level as the tf32x3 forward at these K — at twice the tf32 MMA rate and half the
// shared-memory bytes per FLOP.  Persistent CTAs, double-buffered accumulator, dedicated epilogue warps (as conv.cu).
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/selavi_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int DG_BM = 128;
constexpr int DG_EPI_WARPS = 4;
constexpr int DG_LOADER_WARPS = 8;
constexpr int DG_MMA_WARP = DG_EPI_WARPS + DG_LOADER_WARPS;
constexpr int DG_BPROD_WARP = DG_MMA_WARP + 1;
constexpr int DG_THREADS = (DG_BPROD_WARP + 1) * 32;
constexpr int DG_A_BYTES = DG_BM * 128;     // 128 pixels x 64 bf16 channels
constexpr int DG_MAX_TAPS = 64;
constexpr int DG_MAX_STAGES = 6;

// Stride-2 convolutions: an input pixel only receives contributions from the taps whose parity matches
// (dst + pad - k) % stride == 0, i.e. 1/stride of the taps per strided dimension.  Pixels are therefore processed
// in parity classes (up to st*sh*sw of them): a tile holds 128 pixels of ONE class and its K loop runs over that
// class's taps only (no zero-filled MMA work: 2.25 instead of 9 taps on average for a 3x3 stride-2 conv).
constexpr int DG_MAX_CLASSES = 8;
struct DgClass {
    int o[3];        // first dst coordinate of the class per dimension (t, h, w)
    int n[3];        // number of dst coordinates of the class per dimension
    int pixels;      // nb * n[0] * n[1] * n[2]
    int tile_begin;  // first tile (in units of m-tiles) of the class
    int ntaps, kstages;
    long long woff;  // byte offset of the class's packed weights
    unsigned char taps[DG_MAX_TAPS];
};

struct DgradParams {
    int nclasses;
    DgClass cls[DG_MAX_CLASSES];
    const __nv_bfloat16* z_hi;   // [pixels_out][cs] gradient wrt the conv output, split
    const __nv_bfloat16* z_lo;
    float* dst;                  // [M = pixels_in][cd]
    const unsigned char* wpack;  // [ntiles][kstages][2][BNt][128B] bf16
    int nb, ts, hs, ws, cs;      // gathered tensor (dz) geometry: forward OUTPUT dims
    int td, hd, wd, cd;          // dx geometry: forward INPUT dims
    int kt, kh, kw, st, sh, sw, pt, ph, pw;
    int M, m_tiles, bnt, ntiles, stages, accumulate, passes;
    uint32_t tmem_cols;
};

__device__ __forceinline__ void dg_cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

__global__ void __launch_bounds__(DG_THREADS, 1) dgrad_bf16_kernel(const DgradParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int b_tile_bytes = p.bnt * 128;
    const int stage_bytes = 2 * DG_A_BYTES + 2 * b_tile_bytes;   // A_hi | A_lo | B_hi | B_lo
    unsigned char* tail = smem + (size_t)p.stages * stage_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
    uint64_t* empty_bar = full_bar + DG_MAX_STAGES;
    uint64_t* tfull_bar = empty_bar + DG_MAX_STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    int* tap_dt = reinterpret_cast<int*>(tmem_slot + 2);
    int* tap_off = tap_dt + DG_MAX_TAPS;
    unsigned char* s_stage = reinterpret_cast<unsigned char*>(   // [EPI_WARPS] x (32 rows x 128 B | 32 row indices)
        (reinterpret_cast<uintptr_t>(tap_off + DG_MAX_TAPS) + 127) & ~(uintptr_t)127);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int taps = p.kt * p.kh * p.kw;
    const int total_tiles = p.m_tiles * p.ntiles;

    if (tid == 0) {
        for (int s = 0; s < p.stages; ++s) {
            sv::mbar_init(&full_bar[s], DG_LOADER_WARPS + 1);
            sv::mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            sv::mbar_init(&tfull_bar[a], 1);
            sv::mbar_init(&tempty_bar[a], DG_EPI_WARPS);
        }
        sv::fence_barrier_init();
    }
    for (int t = tid; t < taps; t += DG_THREADS) {
        const int kw_ = t % p.kw, kh_ = (t / p.kw) % p.kh, kt_ = t / (p.kw * p.kh);
        tap_dt[t] = kt_ | (kh_ << 8) | (kw_ << 16);
        // transposed gather: src = (dst + pad - k) / stride == floor((dst+pad)/stride) - (k >> log2(stride)) when valid
        const int qt = p.st == 2 ? (kt_ >> 1) : kt_, qh = p.sh == 2 ? (kh_ >> 1) : kh_, qw = p.sw == 2 ? (kw_ >> 1) : kw_;
        tap_off[t] = -((qt * p.hs + qh) * p.ws + qw);
    }
    if (warp == DG_MMA_WARP) {
        sv::tmem_alloc(tmem_slot, p.tmem_cols);
        sv::tmem_relinquish();
    }
    sv::tc_fence_before();
    __syncthreads();
    sv::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= DG_EPI_WARPS && warp < DG_MMA_WARP) {
        // ------------------------------------------------------------------ A loaders: cp.async gathers
        const int ltid = tid - DG_EPI_WARPS * 32;
        const int c = ltid & 7;    // 16-byte chunk (8 channels) within the 128B K row
        const int r0 = ltid >> 3;  // rows r0 + 32*j
        const int C8 = p.cs >> 3;
        const uint32_t sw_off = (uint32_t)((c ^ (r0 & 7)) << 4);
        const bool with_lo = p.passes == 3;
        int stage = 0, prev_stage = -1;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int m_tile = tile / p.ntiles;
            int ci = 0;
            while (ci + 1 < p.nclasses && m_tile >= p.cls[ci + 1].tile_begin) ++ci;
            const DgClass& cl = p.cls[ci];
            const int m0 = (m_tile - cl.tile_begin) * DG_BM;   // first row of the tile inside its class
            const int kstages = cl.kstages;
            int pb[4];
            uint32_t vm[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int m = m0 + r0 + 32 * j;
                pb[j] = 0;
                vm[j] = 0;
                if (m < cl.pixels) {
                    const int iw = m % cl.n[2];
                    const int t1 = m / cl.n[2];
                    const int ih = t1 % cl.n[1];
                    const int t2 = t1 / cl.n[1];
                    const int it_ = t2 % cl.n[0];
                    const int n_ = t2 / cl.n[0];
                    const int w_ = cl.o[2] + iw * p.sw, h_ = cl.o[1] + ih * p.sh, t_ = cl.o[0] + it_ * p.st;
                    const int at = t_ + p.pt, ah = h_ + p.ph, aw = w_ + p.pw;
                    uint32_t mt = 0, mh = 0, mw = 0;
                    for (int k = 0; k < p.kt; ++k) {
                        int u = at - k;
                        bool ok = u >= 0;
                        if (p.st == 2) { ok &= !(u & 1); u >>= 1; }
                        mt |= (uint32_t)(ok & (u < p.ts)) << k;
                    }
                    for (int k = 0; k < p.kh; ++k) {
                        int u = ah - k;
                        bool ok = u >= 0;
                        if (p.sh == 2) { ok &= !(u & 1); u >>= 1; }
                        mh |= (uint32_t)(ok & (u < p.hs)) << k;
                    }
                    for (int k = 0; k < p.kw; ++k) {
                        int u = aw - k;
                        bool ok = u >= 0;
                        if (p.sw == 2) { ok &= !(u & 1); u >>= 1; }
                        mw |= (uint32_t)(ok & (u < p.ws)) << k;
                    }
                    const int bt = p.st == 2 ? (at >> 1) : at, bh = p.sh == 2 ? (ah >> 1) : ah, bw = p.sw == 2 ? (aw >> 1) : aw;
                    pb[j] = ((n_ * p.ts + bt) * p.hs + bh) * p.ws + bw;
                    vm[j] = mt | (mh << 8) | (mw << 16) | 0x80000000u;
                }
            }
            int tap = 0, c8 = c;  // flattened K chunk q = 8*ks + c -> (index in the class's tap list, c8)
            while (c8 >= C8) {
                c8 -= C8;
                ++tap;
            }
            for (int ks = 0; ks < kstages; ++ks) {
                sv::mbar_wait(&empty_bar[stage], phase ^ 1);
                const uint32_t a_hi = sv::smem_u32(smem + (size_t)stage * stage_bytes);
                const uint32_t a_lo = a_hi + DG_A_BYTES;
                int pk = 0, off = 0;
                const bool kvalid = tap < cl.ntaps;
                if (kvalid) {
                    const int tp = cl.taps[tap];
                    pk = tap_dt[tp];
                    off = tap_off[tp];
                }
                const int s_t = pk & 255, s_h = 8 + ((pk >> 8) & 255), s_w = 16 + ((pk >> 16) & 255);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t m_ = vm[j];
                    const bool ok = kvalid && ((m_ >> 31) & (m_ >> s_t) & (m_ >> s_h) & (m_ >> s_w) & 1u);
                    const size_t goff = ok ? ((size_t)(pb[j] + off) * p.cs + c8 * 8) : 0;
                    const uint32_t nbytes = ok ? 16u : 0u;
                    const uint32_t row_off = (uint32_t)((r0 + 32 * j) * 128) + sw_off;
                    dg_cp_async16(a_hi + row_off, p.z_hi + goff, nbytes);
                    if (with_lo) dg_cp_async16(a_lo + row_off, p.z_lo + goff, nbytes);
                }
                asm volatile("cp.async.commit_group;\n" ::: "memory");
                if (prev_stage >= 0) {   // the previous stage's copies have landed: publish it
                    asm volatile("cp.async.wait_group 1;\n" ::: "memory");
                    sv::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) sv::mbar_arrive(&full_bar[prev_stage]);
                }
                prev_stage = stage;
                c8 += 8;
                while (c8 >= C8 && tap < cl.ntaps) {
                    c8 -= C8;
                    ++tap;
                }
                if (++stage == p.stages) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        sv::fence_proxy_async();
        __syncwarp();
        if (lane == 0 && prev_stage >= 0) sv::mbar_arrive(&full_bar[prev_stage]);
    } else if (warp < DG_EPI_WARPS) {
        // ------------------------------------------------------------------ epilogue
        const int quad = warp;
        unsigned char* my_stage = s_stage + (size_t)warp * sv::EPI_STAGE_BYTES;
        const uint32_t stg = sv::smem_u32(my_stage);
        int* row_pix = reinterpret_cast<int*>(my_stage + 32 * 128);
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const int m_tile = tile / p.ntiles, ntile = tile % p.ntiles;
            int ci = 0;
            while (ci + 1 < p.nclasses && m_tile >= p.cls[ci + 1].tile_begin) ++ci;
            const DgClass& cl = p.cls[ci];
            const int mc = (m_tile - cl.tile_begin) * DG_BM + quad * 32 + lane;   // row inside the class
            const bool row_ok = mc < cl.pixels;
            int m = -1;
            if (row_ok) {
                const int iw = mc % cl.n[2];
                const int t1 = mc / cl.n[2];
                const int ih = t1 % cl.n[1];
                const int t2 = t1 / cl.n[1];
                const int it_ = t2 % cl.n[0];
                const int n_ = t2 / cl.n[0];
                m = ((n_ * p.td + cl.o[0] + it_ * p.st) * p.hd + cl.o[1] + ih * p.sh) * p.wd + cl.o[2] + iw * p.sw;
            }
            row_pix[lane] = m;
            __syncwarp();
            const int n_base = ntile * p.bnt;
            sv::mbar_wait(&tfull_bar[acc], (uint32_t)((it >> 1) & 1));
            sv::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc * (p.tmem_cols >> 1);
            // coalesced drain through the per-warp staging tile, see sv::epi_drain_group
            sv::epi_drain_tile(taddr, p.bnt, 1.f, row_ok, stg, row_pix, p.dst, p.cd, n_base, p.accumulate, true, nullptr, lane);
            sv::tc_fence_before();
            __syncwarp();
            if (lane == 0) sv::mbar_arrive(&tempty_bar[acc]);
        }
    } else if (warp == DG_MMA_WARP) {
        // MMA issuer: converged warp, warp-uniform schedule, one elected lane issues (see conv.cu)
        {
            const uint32_t idesc = sv::make_idesc_f16(DG_BM, p.bnt, 1, 1, 0, 0);  // bf16 x bf16, K-major
            const uint32_t tm0 = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint64_t desc_fixed = sv::make_smem_desc_sw128(0, 16, 1024);
            int stage = 0, it = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                const int m_tile = tile / p.ntiles;
                int ci = 0;
                while (ci + 1 < p.nclasses && m_tile >= p.cls[ci + 1].tile_begin) ++ci;
                const int kstages = p.cls[ci].kstages;
                sv::mbar_wait(&tempty_bar[acc], (uint32_t)(((it >> 1) & 1) ^ 1));
                sv::tc_fence_after();
                const uint32_t d_tmem = tm0 + (uint32_t)acc * (p.tmem_cols >> 1);
                for (int ks = 0; ks < kstages; ++ks) {
                    sv::mbar_wait(&full_bar[stage], phase);
                    sv::tc_fence_after();
                    const uint32_t a_hi = sv::smem_u32(smem + (size_t)stage * stage_bytes);
                    // k-step k4 (K = 16 bf16 = 32 bytes) adds 2 to the descriptor's 16-byte-unit address field
                    const uint64_t da_hi = desc_fixed | (uint64_t)((a_hi & 0x3FFFFu) >> 4);
                    const uint64_t da_lo = da_hi + (uint64_t)(DG_A_BYTES >> 4);
                    const uint64_t db_hi = da_lo + (uint64_t)(DG_A_BYTES >> 4);
                    const uint64_t db_lo = db_hi + (uint64_t)(b_tile_bytes >> 4);
                    if (sv::elect_one()) {
                        if (p.passes == 3) {
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4) {
                                sv::umma_f16(d_tmem, da_lo + 2 * k4, db_hi + 2 * k4, idesc, (uint32_t)(ks | k4));
                                sv::umma_f16(d_tmem, da_hi + 2 * k4, db_lo + 2 * k4, idesc, 1u);
                                sv::umma_f16(d_tmem, da_hi + 2 * k4, db_hi + 2 * k4, idesc, 1u);
                            }
                        } else {
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4)
                                sv::umma_f16(d_tmem, da_hi + 2 * k4, db_hi + 2 * k4, idesc, (uint32_t)(ks | k4));
                        }
                        sv::umma_commit(&empty_bar[stage]);
                    }
                    __syncwarp();
                    if (++stage == p.stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (sv::elect_one()) sv::umma_commit(&tfull_bar[acc]);
                __syncwarp();
            }
        }
    } else {
        if (lane == 0) {   // B producer: one TMA bulk copy per stage
            const uint32_t bytes = (uint32_t)(p.passes == 3 ? 2 * b_tile_bytes : b_tile_bytes);
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int ntile = tile % p.ntiles, m_tile = tile / p.ntiles;
                int ci = 0;
                while (ci + 1 < p.nclasses && m_tile >= p.cls[ci + 1].tile_begin) ++ci;
                const int kstages = p.cls[ci].kstages;
                const unsigned char* wsrc = p.wpack + p.cls[ci].woff + (size_t)ntile * kstages * 2 * b_tile_bytes;
                for (int ks = 0; ks < kstages; ++ks) {
                    sv::mbar_wait(&empty_bar[stage], phase ^ 1);
                    sv::mbar_arrive_expect_tx(&full_bar[stage], bytes);
                    sv::bulk_g2s(smem + (size_t)stage * stage_bytes + 2 * DG_A_BYTES, wsrc + (size_t)ks * 2 * b_tile_bytes,
                                 bytes, &full_bar[stage]);
                    if (++stage == p.stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    }
    __syncthreads();
    if (warp == DG_MMA_WARP) {
        sv::tc_fence_after();
        sv::tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// W [co][ci][taps] -> B operand of the data gradient in bf16 hi/lo: n = ci, k = tap*cs + co (cs = padded co),
// tiles [ntile][kstage of 64 k][hi|lo][bnt rows][128 B], 128B-swizzled.
struct DgTapList {
    int ntaps;
    unsigned char taps[DG_MAX_TAPS];
};
__global__ void dgrad_pack_weights_bf16_kernel(const float* __restrict__ W, int co, int ci, int taps_total, const DgTapList tl,
                                               int cs, int bnt, int ntiles, int kstages, __nv_bfloat16* __restrict__ out) {
    const size_t total = (size_t)ntiles * kstages * bnt * 64;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int e = (int)(idx & 7);
        const int c = (int)((idx >> 3) & 7);
        const int n = (int)((idx >> 6) % bnt);
        const size_t blk = (idx >> 6) / bnt;
        const int ks = (int)(blk % kstages);
        const int nt = (int)(blk / kstages);
        const int k = ks * 64 + c * 8 + e;
        const int tj = k / cs, kc = k % cs;
        const int nn = nt * bnt + n;
        float val = 0.f;
        if (tj < tl.ntaps && nn < ci && kc < co) val = W[((size_t)kc * ci + nn) * taps_total + tl.taps[tj]];
        const __nv_bfloat16 hi = __float2bfloat16_rn(val);
        const __nv_bfloat16 lo = __float2bfloat16_rn(val - __bfloat162float(hi));
        __nv_bfloat16* base = out + blk * (size_t)(2 * bnt * 64);
        const int pos = n * 64 + ((c ^ (n & 7)) << 3) + e;
        base[pos] = hi;
        base[bnt * 64 + pos] = lo;
    }
}

void dg_tiles(int n_out, int* bnt, int* ntiles) {
    int nt = (n_out + 255) / 256;
    int per = (n_out + nt - 1) / nt;
    per = (per + 15) & ~15;
    *bnt = per;
    *ntiles = nt;
}

struct DgPlan {
    int nclasses, bnt, ntiles, m_tiles;
    DgClass cls[DG_MAX_CLASSES];
    size_t wbytes;
};

// geom in dgrad mode: [1, nb, ts,hs,ws,cs (dz: forward OUTPUT dims), td,hd,wd,cd (dx: forward INPUT dims), kt,kh,kw, st,sh,sw, pt,ph,pw, ci]
int dg_plan(const int* g, DgPlan* pl) {
    const int nb = g[1], D[3] = {g[6], g[7], g[8]}, K[3] = {g[10], g[11], g[12]}, S[3] = {g[13], g[14], g[15]}, P[3] = {g[16], g[17], g[18]};
    const int cs = g[5], n_out = g[19];
    if (K[0] * K[1] * K[2] > DG_MAX_TAPS) return -1;
    dg_tiles(n_out, &pl->bnt, &pl->ntiles);
    pl->nclasses = 0;
    int tile = 0;
    size_t woff = 0;
    for (int rt = 0; rt < S[0]; ++rt)
        for (int rh = 0; rh < S[1]; ++rh)
            for (int rw = 0; rw < S[2]; ++rw) {
                const int r[3] = {rt, rh, rw};
                DgClass c;
                c.pixels = nb;
                for (int d = 0; d < 3; ++d) {
                    c.o[d] = ((r[d] - P[d]) % S[d] + S[d]) % S[d];      // dst = o + S*i  <=>  (dst + pad) % S == r
                    c.n[d] = c.o[d] < D[d] ? (D[d] - c.o[d] + S[d] - 1) / S[d] : 0;
                    c.pixels *= c.n[d];
                }
                if (c.pixels == 0) continue;
                c.ntaps = 0;
                for (int kt = 0; kt < K[0]; ++kt)
                    for (int kh = 0; kh < K[1]; ++kh)
                        for (int kw = 0; kw < K[2]; ++kw)
                            if (kt % S[0] == rt && kh % S[1] == rh && kw % S[2] == rw)
                                c.taps[c.ntaps++] = (unsigned char)((kt * K[1] + kh) * K[2] + kw);
                // a class without taps still has to be written (zeros): give it one K stage of zero weights
                c.kstages = c.ntaps > 0 ? (c.ntaps * (cs >> 3) + 7) / 8 : 1;
                c.tile_begin = tile;
                c.woff = (long long)woff;
                tile += (c.pixels + DG_BM - 1) / DG_BM;
                woff += (size_t)pl->ntiles * c.kstages * 2 * pl->bnt * 128;
                pl->cls[pl->nclasses++] = c;
            }
    pl->m_tiles = tile;
    pl->wbytes = woff;
    return 0;
}

}  // namespace

extern "C" size_t selavi_dgrad_wpack_bytes(const int* geom) {
    DgPlan pl;
    if (!geom || dg_plan(geom, &pl) != 0) return 0;
    return pl.wbytes;
}

extern "C" int selavi_dgrad_pack_weights(const float* W, const int* geom, int co, void* wpack, void* stream) {
    if (!W || !wpack || !geom || co <= 0) return selavi_fail(-1, "dgrad_pack_weights: bad arguments");
    DgPlan pl;
    if (dg_plan(geom, &pl) != 0) return selavi_fail(-1, "dgrad_pack_weights: kernel too large");
    const int ci = geom[19], cs = geom[5], taps = geom[10] * geom[11] * geom[12];
    if (cs & 7) return selavi_fail(-1, "dgrad_pack_weights: channel stride must be a multiple of 8");
    for (int c = 0; c < pl.nclasses; ++c) {
        DgTapList tl;
        tl.ntaps = pl.cls[c].ntaps;
        for (int j = 0; j < tl.ntaps; ++j) tl.taps[j] = pl.cls[c].taps[j];
        const size_t total = (size_t)pl.ntiles * pl.cls[c].kstages * pl.bnt * 64;
        int blocks = (int)((total + 255) / 256);
        if (blocks > 148 * 16) blocks = 148 * 16;
        dgrad_pack_weights_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
            W, co, ci, taps, tl, cs, pl.bnt, pl.ntiles, pl.cls[c].kstages,
            reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<unsigned char*>(wpack) + pl.cls[c].woff));
    }
    SV_CUDA_CHECK(cudaGetLastError(), "dgrad_pack_weights: launch");
    return 0;
}

// geom: the 20-int geometry of selavi_conv_gemm in mode 1 (gathered tensor = dz with the forward OUTPUT dims, dst = dx)
extern "C" int selavi_conv_dgrad_bf16(const void* z_hi, const void* z_lo, float* dx, const void* wpack, const int* geom,
                                      int accumulate, int passes, void* stream) {
    if (!z_hi || !z_lo || !dx || !wpack || !geom) return selavi_fail(-1, "conv_dgrad_bf16: null argument");
    if (geom[0] != 1) return selavi_fail(-1, "conv_dgrad_bf16: geometry must be in dgrad mode");
    DgradParams p;
    p.z_hi = reinterpret_cast<const __nv_bfloat16*>(z_hi);
    p.z_lo = reinterpret_cast<const __nv_bfloat16*>(z_lo);
    p.dst = dx;
    p.wpack = reinterpret_cast<const unsigned char*>(wpack);
    p.nb = geom[1]; p.ts = geom[2]; p.hs = geom[3]; p.ws = geom[4]; p.cs = geom[5];
    p.td = geom[6]; p.hd = geom[7]; p.wd = geom[8]; p.cd = geom[9];
    p.kt = geom[10]; p.kh = geom[11]; p.kw = geom[12];
    p.st = geom[13]; p.sh = geom[14]; p.sw = geom[15];
    p.pt = geom[16]; p.ph = geom[17]; p.pw = geom[18];
    if ((p.cs & 7) || (p.cd & 3)) return selavi_fail(-1, "conv_dgrad_bf16: bad channel strides");
    if (p.kt * p.kh * p.kw > DG_MAX_TAPS || p.kt > 8 || p.kh > 8 || p.kw > 8) return selavi_fail(-1, "conv_dgrad_bf16: kernel too large");
    if ((p.st != 1 && p.st != 2) || (p.sh != 1 && p.sh != 2) || (p.sw != 1 && p.sw != 2)) return selavi_fail(-1, "conv_dgrad_bf16: stride must be 1 or 2");
    if (passes != 1 && passes != 3) return selavi_fail(-1, "conv_dgrad_bf16: passes must be 1 or 3");
    const long long M = (long long)p.nb * p.td * p.hd * p.wd;
    if (M <= 0 || M > 0x7fffffffLL || (long long)p.nb * p.ts * p.hs * p.ws > 0x7fffffffLL) return selavi_fail(-1, "conv_dgrad_bf16: bad pixel count");
    p.M = (int)M;
    DgPlan pl;
    if (dg_plan(geom, &pl) != 0) return selavi_fail(-1, "conv_dgrad_bf16: kernel too large");
    p.bnt = pl.bnt; p.ntiles = pl.ntiles; p.m_tiles = pl.m_tiles; p.nclasses = pl.nclasses;
    for (int c = 0; c < pl.nclasses; ++c) p.cls[c] = pl.cls[c];
    if (p.ntiles * p.bnt < p.cd) return selavi_fail(-1, "conv_dgrad_bf16: cd exceeds the tiled channel range");
    p.accumulate = accumulate;
    p.passes = passes;
    uint32_t cols = 32;
    while ((int)cols < 2 * p.bnt) cols <<= 1;
    p.tmem_cols = cols;
    const int stage_bytes = 2 * DG_A_BYTES + 2 * p.bnt * 128;
    const int tail_bytes = (2 * DG_MAX_STAGES + 4) * 8 + 8 + 2 * DG_MAX_TAPS * 4 + DG_EPI_WARPS * sv::EPI_STAGE_BYTES + 128 + 64;
    int stages = (227 * 1024 - 1024 - tail_bytes) / stage_bytes;
    if (stages > DG_MAX_STAGES) stages = DG_MAX_STAGES;
    if (stages < 2) return selavi_fail(-1, "conv_dgrad_bf16: tile does not fit shared memory");
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + tail_bytes + 1024;
    SV_CUDA_CHECK(cudaFuncSetAttribute(dgrad_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                  "conv_dgrad_bf16: cudaFuncSetAttribute");
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int total_tiles = p.m_tiles * p.ntiles;
    dgrad_bf16_kernel<<<total_tiles < sms ? total_tiles : sms, DG_THREADS, smem, (cudaStream_t)stream>>>(p);
    SV_CUDA_CHECK(cudaGetLastError(), "conv_dgrad_bf16: launch");
    return 0;
}
