// Tap-reuse ("halo") implicit-GEMM forward convolution on tcgen05 in fp16x3, for the stride-1 factorised convs of
// R(2+1)D (tv:video/resnet.py:45-61: 1x3x3 spatial and 3x1x1 temporal, padding 1) and the 3x3 stride-1 convs of
// the audio ResNet (tv:resnet.py:59-105) — the layers that dominate the forward of model.py:93-121.
//
// conv.cu gathers (and BN-normalises, splits, stores) every input element once PER TAP: 3x / 9x the loader work
// and L2->SM traffic of the tensor itself, which is what bounds the N=64..144 layers of layer1.  Here a CTA
// stages the input window of its 128 output pixels ONCE per 64-channel chunk as a linear array of 128-byte rows
// (K-major, SWIZZLE_128B) and every tap is the same shared-memory tile read through a row-shifted descriptor:
//   temporal: tile = 8 frames x 16 positions, slot row = (frame+1)*16 + position, tap kt shifts by 16*kt rows;
//   spatial:  tile = 128 consecutive positions of the row-padded frame (pitch WP = W+2), slot row q = padded
//             input index, tap (kh,kw) shifts by kh*WP + kw rows (the 2 junk columns per row are masked out).
// Operands are fp16 hi/lo pairs (hi = fp16(x), lo = fp16(x - hi): 22 significant bits like the tf32x3 split, at
// twice the MMA rate and half the shared-memory bytes); activations are pre-scaled by 2^4 and weights by 2^8 so
// that the lo parts stay clear of the fp16 subnormal range, the epilogue multiplies by 2^-12 (all exact).
// Loaders fuse the previous layer's train-mode BatchNorm+ReLU; the epilogue emits the BN partial sums (as conv.cu).
//
// The same kernel computes the DATA GRADIENT of these convolutions (loss.backward(), main.py:298): for stride 1 it
// is the same convolution of dz with the taps flipped and the weight matrix transposed.  dz arrives as bf16 hi/lo
// planes (written by the BatchNorm-backward kernel), so in that mode the loaders are pure 16-byte cp.async copies
// and the MMAs run in bf16x3 without scaling (gradients need the bf16 exponent range).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "../../include/selavi_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int HL_EPI_WARPS = 4;
constexpr int HL_LOADER_WARPS = 8;
constexpr int HL_MMA_WARP = HL_EPI_WARPS + HL_LOADER_WARPS;
constexpr int HL_BPROD_WARP = HL_MMA_WARP + 1;
constexpr int HL_THREADS = (HL_BPROD_WARP + 1) * 32;
constexpr int HL_MAX_A = 3;
constexpr int HL_MAX_B = 16;
constexpr float HL_ASCALE = 16.f;
constexpr float HL_WSCALE = 256.f;
constexpr float HL_OSCALE = 1.f / (16.f * 256.f);

struct HaloParams {
    const float* src;            // src_kind 0: fp32 activations (normalised / split to fp16 hi/lo by the loaders)
    const __nv_bfloat16* src_hi; // src_kind 1: pre-split bf16 planes (copied by cp.async)
    const __nv_bfloat16* src_lo;
    float* dst;
    const unsigned char* wpack;  // [ntile][kchunk][tap][hi|lo][bnt][128 B] fp16
    const float* pro_scale;      // [cs] or null
    const float* pro_shift;
    float* stats;                // [m_tiles][2][ntiles*bnt] or null
    sv::EpiBwdStat bwd;          // data gradient: the following BatchNorm-backward statistics instead of sum / sum of squares
    int has_bwd;
    int mode;                    // 0 temporal 3x1x1, 1 spatial 1x3x3
    int nb, T, H, W, cs, cd;
    int S, WP;
    uint32_t wp_magic;           // ceil(2^32 / WP)
    int TB, SB, OB;
    int rows, rows_alloc;
    int m_tiles, bnt, ntiles, kchunks, taps;
    int n_a, n_b;
    int pro_relu, use_base_off;
    int src_kind, accumulate;
    int dbg;                     // timing diagnostics only (results invalid): 2 no B reloads, 4 no A gathers, 8 no stores
    float oscale;
    uint32_t tmem_cols;
};

__device__ __forceinline__ void hl_cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

__device__ __forceinline__ void hl_st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void hl_named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint32_t hl_h2_bits(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

struct HlBatch {
    float4 v[4][2];
    uint32_t ok;
};

// PAIR: the two CTAs of a cluster form a tcgen05 CTA pair (cta_group::2).  Each CTA still owns one 128-pixel tile (its own
// input window, loaders, TMEM accumulators and epilogue) but the pair shares every weight tile: each CTA stages HALF of
// its rows (N/2 output channels) and the leader CTA issues M=256 MMAs that read both halves and write both CTAs' TMEM.
// That halves the weight bytes streamed into each SM and the shared-memory bytes the tensor core reads for B.
template <bool PAIR>
__global__ void __launch_bounds__(HL_THREADS, 1) conv_halo_kernel(const HaloParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int a_plane = p.rows_alloc * 128;
    const int a_slot_bytes = 2 * a_plane;
    const int b_plane = (PAIR ? p.bnt >> 1 : p.bnt) * 128;   // rows of the weight tile staged by this CTA
    const int b_slot_bytes = 2 * b_plane;                     // hi | lo
    unsigned char* a_base = smem;
    unsigned char* b_base = a_base + (size_t)p.n_a * a_slot_bytes;
    unsigned char* tail = b_base + (size_t)p.n_b * b_slot_bytes;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(tail);
    uint64_t* a_empty = a_full + HL_MAX_A;
    uint64_t* b_full = a_empty + HL_MAX_A;
    uint64_t* b_empty = b_full + HL_MAX_B;
    uint64_t* b_peer = b_empty + HL_MAX_B;    // PAIR, leader only: the peer CTA's half of slot s has landed
    uint64_t* tfull_bar = b_peer + HL_MAX_B;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* s_stat = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~(uintptr_t)15);   // [EPI_WARPS][2][256]
    float* s_pro = s_stat + HL_EPI_WARPS * 512;                 // [2][kchunks*64]: scale*16 | shift*16
    unsigned char* s_stage = reinterpret_cast<unsigned char*>(     // [EPI_WARPS] x (32 rows x 128 B | 32 row indices)
        (reinterpret_cast<uintptr_t>(s_pro + 2 * p.kchunks * 64) + 127) & ~(uintptr_t)127);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ctab = p.kchunks * 64;
    // work items: (m_tile, n_tile), or for a CTA pair (pair of consecutive m_tiles, n_tile); CTA `rank` of the pair takes
    // m_tile 2*m_pair + rank (an odd tail recomputes the last tile and discards it: both CTAs run the same schedule)
    const uint32_t rank = PAIR ? sv::cluster_ctarank() : 0u;
    const int nwork = PAIR ? ((p.m_tiles + 1) >> 1) * p.ntiles : p.m_tiles * p.ntiles;
    const int w_first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int w_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    auto m_tile_of = [&](int w, bool& valid) {
        int mt = w / p.ntiles;
        valid = true;
        if (PAIR) {
            mt = 2 * mt + (int)rank;
            valid = mt < p.m_tiles;
            if (!valid) mt = p.m_tiles - 1;
        }
        return mt;
    };
    const int npair = PAIR ? 2 : 1;

    if (tid == 0) {
        for (int s = 0; s < p.n_a; ++s) {
            sv::mbar_init(&a_full[s], HL_LOADER_WARPS * npair);   // PAIR: both CTAs' loaders arrive at the leader's
            sv::mbar_init(&a_empty[s], 1);
        }
        for (int s = 0; s < p.n_b; ++s) {
            sv::mbar_init(&b_full[s], 1);
            sv::mbar_init(&b_empty[s], 1);
            sv::mbar_init(&b_peer[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            sv::mbar_init(&tfull_bar[a], 1);
            sv::mbar_init(&tempty_bar[a], HL_EPI_WARPS * npair);
        }
        sv::fence_barrier_init();
    }
    for (int i = tid; i < ctab; i += HL_THREADS) {
        float sc = HL_ASCALE, sf = 0.f;
        if (p.pro_scale != nullptr) {
            sc = i < p.cs ? p.pro_scale[i] * HL_ASCALE : 0.f;
            sf = i < p.cs ? p.pro_shift[i] * HL_ASCALE : 0.f;
        }
        s_pro[i] = sc;
        s_pro[ctab + i] = sf;
    }
    if (PAIR) {   // both CTAs' barriers are initialised (and both CTAs are running) before anything crosses the pair
        __syncthreads();
        sv::cluster_sync_all();
    }
    if (warp == HL_MMA_WARP) {
        if (PAIR) {
            sv::tmem_alloc_2cta(tmem_slot, p.tmem_cols);
            sv::tmem_relinquish_2cta();
        } else {
            sv::tmem_alloc(tmem_slot, p.tmem_cols);
            sv::tmem_relinquish();
        }
    }
    sv::tc_fence_before();
    __syncthreads();
    sv::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // barrier arrivals that the pair's leader (rank 0) consumes
    auto arrive_leader = [&](uint64_t* bar) {
        if (PAIR) sv::mbar_arrive_cluster(sv::mapa_u32(sv::smem_u32(bar), 0u));
        else sv::mbar_arrive(bar);
    };

    if (warp >= HL_EPI_WARPS && warp < HL_MMA_WARP) {
        // ------------------------------------------------------------------ A loaders (256 threads)
        const int ltid = tid - HL_EPI_WARPS * 32;
        const int c = ltid & 7;        // 16-byte chunk (8 fp16 channels) of the 128-byte row
        const int rbase = ltid >> 3;   // slot rows rbase + 32*j
        const uint32_t sw_off = (uint32_t)((c ^ (rbase & 7)) << 4);
        const int nrow_thr = p.rows_alloc >> 5;
        const bool relu = p.pro_relu != 0;
        int pb[8];
        auto setup = [&](int w) {
            bool valid_;
            const int m_tile = m_tile_of(w, valid_);
            if (p.mode == 0) {
                const int sb = m_tile % p.SB;
                const int t1 = m_tile / p.SB;
                const int tb = t1 % p.TB;
                const int n = t1 / p.TB;
                const int t0 = tb * 8 - 1, s0 = sb * 16;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int r = rbase + 32 * j;
                    const int t = t0 + (r >> 4), s = s0 + (r & 15);
                    const bool ok = (r < p.rows) & (t >= 0) & (t < p.T) & (s < p.S);
                    pb[j] = ok ? (n * p.T + t) * p.S + s : -1;
                }
            } else {
                const int ob = m_tile % p.OB;
                const int f = m_tile / p.OB;
                const int i0 = ob * 128;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int q = rbase + 32 * j;
                    const int i = i0 + q;
                    const int hp = (int)__umulhi((uint32_t)i, p.wp_magic);
                    const int wp = i - hp * p.WP;
                    const bool ok = (q < p.rows) & (hp >= 1) & (hp <= p.H) & (wp >= 1) & (wp <= p.W);
                    pb[j] = ok ? (f * p.H + hp - 1) * p.W + wp - 1 : -1;
                }
            }
        };
        // 16-byte column c of chunk kc is read by the MMAs only when it lies inside the chunk's valid k16 steps
        auto col_used = [&](int kc) { return c * 8 < ((p.cs - kc * 64 + 15) & ~15); };
        auto gather = [&](HlBatch& b, const int jbase, int kc) {
            const int ch = kc * 64 + c * 8;
            const bool chan_ok = ch < p.cs;
            b.ok = 0;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int j = jbase + jj;
                b.v[jj][0] = make_float4(0.f, 0.f, 0.f, 0.f);
                b.v[jj][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j < nrow_thr && pb[j] >= 0 && chan_ok) {
                    const float4* g = reinterpret_cast<const float4*>(p.src + (size_t)pb[j] * p.cs + ch);
                    b.v[jj][0] = __ldg(g);
                    b.v[jj][1] = __ldg(g + 1);
                    b.ok |= 1u << jj;
                }
            }
        };
        auto commit = [&](const HlBatch& b, const int jbase, int kc, uint32_t a_hi) {
            if (!col_used(kc)) return;
            const uint32_t a_lo = a_hi + (uint32_t)a_plane;
            const int ch = kc * 64 + c * 8;
            const float4 sc0 = *reinterpret_cast<const float4*>(s_pro + ch);
            const float4 sc1 = *reinterpret_cast<const float4*>(s_pro + ch + 4);
            const float4 sf0 = *reinterpret_cast<const float4*>(s_pro + ctab + ch);
            const float4 sf1 = *reinterpret_cast<const float4*>(s_pro + ctab + ch + 4);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int j = jbase + jj;
                if (j < nrow_thr) {
                    float x[8] = {b.v[jj][0].x, b.v[jj][0].y, b.v[jj][0].z, b.v[jj][0].w,
                                  b.v[jj][1].x, b.v[jj][1].y, b.v[jj][1].z, b.v[jj][1].w};
                    if ((b.ok >> jj) & 1u) {
                        x[0] = fmaf(x[0], sc0.x, sf0.x);
                        x[1] = fmaf(x[1], sc0.y, sf0.y);
                        x[2] = fmaf(x[2], sc0.z, sf0.z);
                        x[3] = fmaf(x[3], sc0.w, sf0.w);
                        x[4] = fmaf(x[4], sc1.x, sf1.x);
                        x[5] = fmaf(x[5], sc1.y, sf1.y);
                        x[6] = fmaf(x[6], sc1.z, sf1.z);
                        x[7] = fmaf(x[7], sc1.w, sf1.w);
                        if (relu) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) x[e] = fmaxf(x[e], 0.f);
                        }
                    }
                    uint32_t hb[4], lb[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const __half2 h = __floats2half2_rn(x[2 * e], x[2 * e + 1]);
                        const float2 hf = __half22float2(h);
                        const __half2 l = __floats2half2_rn(x[2 * e] - hf.x, x[2 * e + 1] - hf.y);
                        hb[e] = hl_h2_bits(h);
                        lb[e] = hl_h2_bits(l);
                    }
                    const uint32_t row_off = (uint32_t)((rbase + 32 * j) * 128) + sw_off;
                    hl_st_shared_v4(a_hi + row_off, hb[0], hb[1], hb[2], hb[3]);
                    hl_st_shared_v4(a_lo + row_off, lb[0], lb[1], lb[2], lb[3]);
                }
            }
        };
        int slot = 0;
        uint32_t phase = 0;
        int tile = w_first, kc = 0;   // tile = work item index
        if (p.src_kind == 1) {
            // pre-split bf16 planes: the slot is filled by zero-filling 16-byte cp.async copies, published one slot late
            int prev_slot = -1;
            while (tile < nwork) {
                if (kc == 0) setup(tile);
                sv::mbar_wait(&a_empty[slot], phase ^ 1);
                const uint32_t a_hi = sv::smem_u32(a_base + (size_t)slot * a_slot_bytes);
                const uint32_t a_lo = a_hi + (uint32_t)a_plane;
                const int ch = kc * 64 + c * 8;
                const bool chan_ok = ch < p.cs;
                const bool used = col_used(kc);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (j < nrow_thr && used && !(p.dbg & 4)) {
                        const bool ok = chan_ok && pb[j] >= 0;
                        const size_t goff = ok ? (size_t)pb[j] * p.cs + ch : 0;
                        const uint32_t nbytes = ok ? 16u : 0u;
                        const uint32_t row_off = (uint32_t)((rbase + 32 * j) * 128) + sw_off;
                        hl_cp_async16(a_hi + row_off, p.src_hi + goff, nbytes);
                        hl_cp_async16(a_lo + row_off, p.src_lo + goff, nbytes);
                    }
                }
                asm volatile("cp.async.commit_group;\n" ::: "memory");
                if (prev_slot >= 0) {
                    asm volatile("cp.async.wait_group 1;\n" ::: "memory");
                    sv::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) arrive_leader(&a_full[prev_slot]);
                }
                prev_slot = slot;
                if (++slot == p.n_a) {
                    slot = 0;
                    phase ^= 1;
                }
                if (++kc == p.kchunks) {
                    kc = 0;
                    tile += w_step;
                }
            }
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
            sv::fence_proxy_async();
            __syncwarp();
            if (lane == 0 && prev_slot >= 0) arrive_leader(&a_full[prev_slot]);
        }
        HlBatch sa, sb;
        const bool no_a = (p.dbg & 4) != 0;
        if (p.src_kind == 1) tile = nwork;
        if (tile < nwork) {
            setup(tile);
            if (!no_a) gather(sa, 0, 0);
        }
        while (tile < nwork) {
            if (!no_a) gather(sb, 4, kc);   // rows 4..7 of this slot in flight while rows 0..3 are converted
            sv::mbar_wait(&a_empty[slot], phase ^ 1);
            const uint32_t a_hi = sv::smem_u32(a_base + (size_t)slot * a_slot_bytes);
            if (!no_a) commit(sa, 0, kc, a_hi);
            int ntile = tile, nkc = kc + 1;
            if (nkc == p.kchunks) {
                nkc = 0;
                ntile += w_step;
            }
            if (ntile < nwork) {
                if (ntile != tile) setup(ntile);
                if (!no_a) gather(sa, 0, nkc);   // first rows of the NEXT slot in flight while rows 4..7 are converted
            }
            if (!no_a) commit(sb, 4, kc, a_hi);
            sv::fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) arrive_leader(&a_full[slot]);
            if (++slot == p.n_a) {
                slot = 0;
                phase ^= 1;
            }
            tile = ntile;
            kc = nkc;
        }
    } else if (warp < HL_EPI_WARPS) {
        // ------------------------------------------------------------------ epilogue (4 warps, one TMEM quadrant each)
        const int quad = warp;
        const int ctot = p.ntiles * p.bnt;
        float* my_stat = s_stat + (size_t)warp * 512;
        unsigned char* my_stage = s_stage + (size_t)warp * sv::EPI_STAGE_BYTES;
        const uint32_t stg = sv::smem_u32(my_stage);
        int* row_pix = reinterpret_cast<int*>(my_stage + 32 * 128);
        int it = 0;
        for (int w = w_first; w < nwork; w += w_step, ++it) {
            const int acc = it & 1;
            bool valid;
            const int m_tile = m_tile_of(w, valid), ntile = w % p.ntiles;
            const int m = quad * 32 + lane;
            bool row_ok;
            int pix;
            if (p.mode == 0) {
                const int sb = m_tile % p.SB;
                const int t1 = m_tile / p.SB;
                const int tb = t1 % p.TB;
                const int n = t1 / p.TB;
                const int t = tb * 8 + (m >> 4), s = sb * 16 + (m & 15);
                row_ok = (t < p.T) & (s < p.S);
                pix = (n * p.T + t) * p.S + s;
            } else {
                const int ob = m_tile % p.OB;
                const int f = m_tile / p.OB;
                const int o = ob * 128 + m;
                const int h = (int)__umulhi((uint32_t)o, p.wp_magic);
                const int w = o - h * p.WP;
                row_ok = (h < p.H) & (w < p.W);
                pix = (f * p.H + h) * p.W + w;
            }
            row_pix[lane] = row_ok ? pix : -1;
            __syncwarp();
            const int n_base = ntile * p.bnt;
            sv::mbar_wait(&tfull_bar[acc], (uint32_t)((it >> 1) & 1));
            sv::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc * (p.tmem_cols >> 1);
            sv::epi_drain_tile(taddr, p.bnt, p.oscale, row_ok, stg, row_pix, p.dst, p.cd, n_base, p.accumulate,
                               valid && !(p.dbg & 8), p.stats != nullptr ? my_stat : nullptr, lane, p.has_bwd ? &p.bwd : nullptr);
            sv::tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_leader(&tempty_bar[acc]);
            if (p.stats != nullptr) {
                hl_named_bar_sync(1, HL_EPI_WARPS * 32);
                for (int i = tid; i < 2 * p.bnt && valid; i += HL_EPI_WARPS * 32) {
                    const int which = i >= p.bnt, col = which ? i - p.bnt : i;
                    float tsum = 0.f;
#pragma unroll
                    for (int q = 0; q < HL_EPI_WARPS; ++q) tsum += s_stat[(size_t)q * 512 + which * 256 + col];
                    p.stats[((size_t)m_tile * 2 + which) * ctot + n_base + col] = tsum;
                }
                hl_named_bar_sync(1, HL_EPI_WARPS * 32);
            }
        }
    } else if (warp == HL_MMA_WARP) {
        // ------------------------------------------------------------------ MMA issuer
        // The whole warp walks the schedule converged (all values warp-uniform, so descriptors live in uniform
        // registers) and one elected lane issues: under a divergent `if (lane == 0)` the compiler has to wrap every
        // tcgen05.mma in an elect / register-broadcast / branch loop, which made the ISSUE the bottleneck (~180 clk per
        // MMA on every layer); tools/umma_rate.cu: N/2 clk per MMA needs a tight unrolled issue sequence.
        if (PAIR && rank != 0) {
            // peer CTA of a pair: its MMAs are issued by the leader.  This warp only relays "my half of weight slot s has
            // landed" (a local TMA completion) to the leader's b_peer[s].
            int sb = 0;
            uint32_t pbp = 0;
            for (int w = w_first; w < nwork; w += w_step) {
                const int nslots = p.kchunks * p.taps;
                for (int i = 0; i < nslots; ++i) {
                    sv::mbar_wait(&b_full[sb], pbp);
                    if (lane == 0) sv::mbar_arrive_cluster(sv::mapa_u32(sv::smem_u32(&b_peer[sb]), 0u));
                    __syncwarp();
                    if (++sb == p.n_b) {
                        sb = 0;
                        pbp ^= 1;
                    }
                }
            }
        } else {
            const int fmt = p.src_kind == 1 ? 1 : 0;   // fp16 x fp16 (forward) or bf16 x bf16 (data gradient), K-major
            const uint32_t idesc = sv::make_idesc_f16(PAIR ? 256 : 128, p.bnt, fmt, fmt, 0, 0);
            const uint32_t tm0 = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint64_t desc_fixed = sv::make_smem_desc_sw128(0, 16, 1024);
            auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t accumulate) {
                if (PAIR) sv::umma_f16_2cta(d, da, db, idesc, accumulate);
                else sv::umma_f16(d, da, db, idesc, accumulate);
            };
            auto commit = [&](uint64_t* bar) {   // (PAIR: the barrier at this offset in both CTAs)
                if (PAIR) sv::umma_commit_2cta(bar);
                else sv::umma_commit(bar);
            };
            int sa = 0, sb = 0, it = 0;
            uint32_t pa = 0, pbp = 0;
            for (int w = w_first; w < nwork; w += w_step, ++it) {
                const int acc = it & 1;
                sv::mbar_wait(&tempty_bar[acc], (uint32_t)(((it >> 1) & 1) ^ 1));
                sv::tc_fence_after();
                const uint32_t d_tmem = tm0 + (uint32_t)acc * (p.tmem_cols >> 1);
                for (int kc = 0; kc < p.kchunks; ++kc) {
                    int nk16 = (p.cs - kc * 64 + 15) >> 4;
                    if (nk16 > 4) nk16 = 4;
                    sv::mbar_wait(&a_full[sa], pa);
                    sv::tc_fence_after();
                    const uint32_t a_hi = sv::smem_u32(a_base + (size_t)sa * a_slot_bytes);
                    for (int tap = 0; tap < p.taps; ++tap) {
                        const int shift_rows = p.mode == 0 ? tap * 16 : (tap / 3) * p.WP + (tap % 3);
                        const uint32_t a_tap = a_hi + (uint32_t)shift_rows * 128u;
                        const uint32_t started = (uint32_t)(kc | tap);   // 0 only for the opening MMA group of a tile
                        // descriptors of k16 step 0; step k adds 32 bytes = 2 to the 16-byte-unit address field
                        uint64_t da_hi = desc_fixed | (uint64_t)((a_tap & 0x3FFFFu) >> 4);
                        if (p.use_base_off) da_hi |= (uint64_t)((a_tap >> 7) & 7u) << 49;
                        const uint64_t da_lo = da_hi + (uint64_t)(a_plane >> 4);
                        sv::mbar_wait(&b_full[sb], pbp);
                        if (PAIR) sv::mbar_wait(&b_peer[sb], pbp);
                        sv::tc_fence_after();
                        const uint32_t b_hi = sv::smem_u32(b_base + (size_t)sb * b_slot_bytes);
                        const uint64_t db_hi = desc_fixed | (uint64_t)((b_hi & 0x3FFFFu) >> 4);
                        const uint64_t db_lo = db_hi + (uint64_t)(b_plane >> 4);
                        if (sv::elect_one()) {
                            if (nk16 == 4) {   // full 64-channel chunk: 12 MMAs back to back, descriptor offsets folded
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    mma(d_tmem, da_lo + 2 * k, db_hi + 2 * k, started | (uint32_t)k);
                                    mma(d_tmem, da_hi + 2 * k, db_lo + 2 * k, 1u);
                                    mma(d_tmem, da_hi + 2 * k, db_hi + 2 * k, 1u);
                                }
                            } else {
                                for (int k = 0; k < nk16; ++k) {
                                    mma(d_tmem, da_lo + 2 * k, db_hi + 2 * k, started | (uint32_t)k);
                                    mma(d_tmem, da_hi + 2 * k, db_lo + 2 * k, 1u);
                                    mma(d_tmem, da_hi + 2 * k, db_hi + 2 * k, 1u);
                                }
                            }
                            commit(&b_empty[sb]);
                        }
                        __syncwarp();
                        if (++sb == p.n_b) {
                            sb = 0;
                            pbp ^= 1;
                        }
                    }
                    if (sv::elect_one()) commit(&a_empty[sa]);
                    __syncwarp();
                    if (++sa == p.n_a) {
                        sa = 0;
                        pa ^= 1;
                    }
                }
                if (sv::elect_one()) commit(&tfull_bar[acc]);
                __syncwarp();
            }
        }
    } else {
        // ------------------------------------------------------------------ B producer (one thread, TMA bulk)
        if (lane == 0) {
            int sb = 0;
            uint32_t pbp = 0;
            const size_t px_bytes = p.src_kind == 1 ? (size_t)p.cs * 2 : (size_t)p.cs * 4;
            auto prefetch_range = [&](size_t pix0, int npix) {   // [pix0, pix0+npix) of the source tensor(s) -> L2
                size_t off = pix0 * px_bytes, left = (size_t)npix * px_bytes;
                while (left > 0) {
                    const uint32_t n = left > 32768 ? 32768u : (uint32_t)left;
                    if (p.src_kind == 1) {
                        sv::bulk_prefetch_l2(reinterpret_cast<const unsigned char*>(p.src_hi) + off, n);
                        sv::bulk_prefetch_l2(reinterpret_cast<const unsigned char*>(p.src_lo) + off, n);
                    } else {
                        sv::bulk_prefetch_l2(reinterpret_cast<const unsigned char*>(p.src) + off, n);
                    }
                    off += n;
                    left -= n;
                }
            };
            // the input window of a tile is a few contiguous pixel ranges: pull the NEXT tile's window into L2 while
            // this one is processed, so that the loaders' gathers see L2 rather than DRAM latency
            auto prefetch_tile = [&](int w) {
                bool valid_;
                const int m_tile = m_tile_of(w, valid_);
                if (p.mode == 0) {
                    const int sb = m_tile % p.SB;
                    const int t1 = m_tile / p.SB;
                    const int tb = t1 % p.TB;
                    const int n = t1 / p.TB;
                    const int s0 = sb * 16;
                    const int ns = p.S - s0 < 16 ? p.S - s0 : 16;
                    for (int t = tb * 8 - 1; t <= tb * 8 + 8; ++t)
                        if (t >= 0 && t < p.T) prefetch_range((size_t)(n * p.T + t) * p.S + s0, ns);
                } else {
                    const int ob = m_tile % p.OB;
                    const int f = m_tile / p.OB;
                    int h0 = (ob * 128) / p.WP - 1, h1 = (ob * 128 + p.rows - 1) / p.WP - 1;
                    if (h0 < 0) h0 = 0;
                    if (h1 > p.H - 1) h1 = p.H - 1;
                    if (h1 >= h0) prefetch_range((size_t)(f * p.H + h0) * p.W, (h1 - h0 + 1) * p.W);
                }
            };
            const size_t full_plane = (size_t)p.bnt * 128;          // one hi (or lo) plane of a weight tile in HBM
            const size_t half_off = PAIR ? (size_t)rank * b_plane : 0;   // this CTA's rows inside the plane
            for (int w = w_first; w < nwork; w += w_step) {
                // (measured: +7 % on the register-staged fp32 loaders, slightly negative on the cp.async-fed bf16 planes)
                if (p.src_kind == 0 && w + w_step < nwork && (p.ntiles == 1 || w % p.ntiles == 0)) prefetch_tile(w + w_step);
                const int ntile = w % p.ntiles;
                const int nslots = p.kchunks * p.taps;
                const unsigned char* wsrc = p.wpack + (size_t)ntile * nslots * 2 * full_plane;
                for (int i = 0; i < nslots; ++i) {
                    sv::mbar_wait(&b_empty[sb], pbp ^ 1);
                    if ((p.dbg & 2) && (w != w_first || i >= p.n_b)) {
                        sv::mbar_arrive(&b_full[sb]);
                    } else {
                        unsigned char* dst = b_base + (size_t)sb * b_slot_bytes;
                        const unsigned char* src = wsrc + (size_t)i * 2 * full_plane + half_off;
                        sv::mbar_arrive_expect_tx(&b_full[sb], (uint32_t)b_slot_bytes);
                        if (PAIR) {   // rows [rank*N/2, (rank+1)*N/2) of the hi plane, then of the lo plane
                            sv::bulk_g2s(dst, src, (uint32_t)b_plane, &b_full[sb]);
                            sv::bulk_g2s(dst + b_plane, src + full_plane, (uint32_t)b_plane, &b_full[sb]);
                        } else {
                            sv::bulk_g2s(dst, src, (uint32_t)b_slot_bytes, &b_full[sb]);
                        }
                    }
                    if (++sb == p.n_b) {
                        sb = 0;
                        pbp ^= 1;
                    }
                }
            }
        }
    }
    __syncthreads();
    if (PAIR) sv::cluster_sync_all();   // the peer may still be signalling this CTA's barriers / reading its shared memory
    if (warp == HL_MMA_WARP) {
        sv::tc_fence_after();
        if (PAIR) sv::tmem_dealloc_2cta(tmem_base, p.tmem_cols);
        else sv::tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// W (torch layout [co][ci][taps]) -> fp16 hi/lo B tiles [ntile][kchunk][tap][hi|lo][bnt rows][128 B], scaled by 2^8.
//   mode 0 (forward):       n = co, k = ci, tap as is
//   mode 1 (data gradient): n = ci, k = co, tap flipped (taps-1-tap), bf16 hi/lo, unscaled
__global__ void conv_halo_pack_kernel(const float* __restrict__ W, int mode, int co, int ci, int taps, int bnt, int ntiles,
                                      int kchunks, uint16_t* __restrict__ out) {
    const size_t total = (size_t)ntiles * kchunks * taps * bnt * 64;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int e = (int)(idx & 7);
        const int c = (int)((idx >> 3) & 7);
        const int n = (int)((idx >> 6) % bnt);
        const size_t blk = (idx >> 6) / bnt;  // (ntile*kchunks + kc)*taps + tap
        const int tap = (int)(blk % taps);
        const int kc = (int)((blk / taps) % kchunks);
        const int nt = (int)(blk / taps / kchunks);
        const int k = kc * 64 + c * 8 + e;
        const int nn = nt * bnt + n;
        float val = 0.f;
        uint16_t hi, lo;
        if (mode == 0) {
            if (nn < co && k < ci) val = W[((size_t)nn * ci + k) * taps + tap] * HL_WSCALE;
            const __half h = __float2half_rn(val);
            const __half l = __float2half_rn(val - __half2float(h));
            hi = __half_as_ushort(h);
            lo = __half_as_ushort(l);
        } else {
            if (nn < ci && k < co) val = W[((size_t)k * ci + nn) * taps + (taps - 1 - tap)];
            const __nv_bfloat16 h = __float2bfloat16_rn(val);
            const __nv_bfloat16 l = __float2bfloat16_rn(val - __bfloat162float(h));
            hi = __bfloat16_as_ushort(h);
            lo = __bfloat16_as_ushort(l);
        }
        uint16_t* base = out + blk * (size_t)(2 * bnt * 64);
        const int pos = n * 64 + ((c ^ (n & 7)) << 3) + e;
        base[pos] = hi;
        base[bnt * 64 + pos] = lo;
    }
}

struct HaloPlan {
    int mode, nb, T, H, W, cs, cd, co, ci;
    int S, WP, TB, SB, OB, rows, rows_alloc, m_tiles, bnt, ntiles, kchunks, taps, n_a, n_b;
    size_t smem, wbytes;
    int pair_ok, n_a2, n_b2;     // CTA-pair variant (cta_group::2): each CTA stages half of every weight tile
    size_t smem2;
};

// geom: the 20-int geometry of selavi_conv_gemm (mode 0 forward, mode 1 data gradient: for these stride-1 "same"
// convolutions both are the same gather with cs / cd swapped).  Returns 0 when the halo kernel supports it, 1 otherwise.
int hl_plan(const int* g, HaloPlan* pl) {
    if (g[0] != 0 && g[0] != 1) return 1;
    const int nb = g[1], ts = g[2], hs = g[3], ws = g[4], cs = g[5], td = g[6], hd = g[7], wd = g[8], cd = g[9];
    const int kt = g[10], kh = g[11], kw = g[12], st = g[13], sh = g[14], sw = g[15], pt = g[16], ph = g[17], pw = g[18];
    if (st != 1 || sh != 1 || sw != 1) return 1;
    if (ts != td || hs != hd || ws != wd) return 1;
    if ((cs & 7) || (cd & 7) || cs <= 0 || cd <= 0) return 1;
    int mode;
    if (kt == 3 && kh == 1 && kw == 1 && pt == 1 && ph == 0 && pw == 0) mode = 0;
    else if (kt == 1 && kh == 3 && kw == 3 && pt == 0 && ph == 1 && pw == 1) mode = 1;
    else return 1;
    pl->mode = mode;
    pl->nb = nb; pl->T = ts; pl->H = hs; pl->W = ws; pl->cs = cs; pl->cd = cd; pl->co = g[19];
    pl->S = hs * ws;
    pl->WP = ws + 2;
    pl->taps = mode == 0 ? 3 : 9;
    pl->kchunks = (cs + 63) / 64;
    const long long pixels = (long long)nb * ts * hs * ws;
    if (pixels <= 0 || pixels > 0x7fffffffLL) return 1;
    if (mode == 0) {
        pl->TB = (ts + 7) / 8;
        pl->SB = (pl->S + 15) / 16;
        pl->OB = 0;
        pl->rows = 160;
        const long long mt = (long long)nb * pl->TB * pl->SB;
        if (mt > 0x3fffffffLL) return 1;
        pl->m_tiles = (int)mt;
    } else {
        if (pl->WP > 62) return 1;   // slot rows 128 + 2*WP + 2 must fit 256
        if ((long long)(hs + 2) * pl->WP + 256 > 60000) return 1;   // exactness range of the magic division
        pl->TB = pl->SB = 0;
        pl->OB = (hs * pl->WP + 127) / 128;
        pl->rows = 128 + 2 * pl->WP + 2;
        const long long mt = (long long)nb * ts * pl->OB;
        if (mt > 0x3fffffffLL) return 1;
        pl->m_tiles = (int)mt;
    }
    pl->rows_alloc = (pl->rows + 31) & ~31;
    const int a_slot = 2 * pl->rows_alloc * 128;
    const int fixed = (2 * HL_MAX_A + 3 * HL_MAX_B + 4) * 8 + 16 + HL_EPI_WARPS * 512 * 4 + 2 * pl->kchunks * 64 * 4 +
                      HL_EPI_WARPS * sv::EPI_STAGE_BYTES + 128 + 16 + 1024 + 64;
    const int budget = 227 * 1024 - fixed;
    for (int nt = (cd + 255) / 256; nt <= 16; ++nt) {
        int per = (cd + nt - 1) / nt;
        per = (per + 15) & ~15;
        const int b_slot = 2 * per * 128;
        int n_a = 2;
        int n_b = (budget - n_a * a_slot) / b_slot;
        if (n_b < 2) continue;
        if (n_b > HL_MAX_B / 2) n_b = HL_MAX_B / 2;
        if (n_b >= 4 && budget - 3 * a_slot - 4 * b_slot >= 0) {   // room for a third A slot
            n_a = 3;
            n_b = (budget - n_a * a_slot) / b_slot;
            if (n_b > HL_MAX_B / 2) n_b = HL_MAX_B / 2;
        }
        pl->bnt = per;
        pl->ntiles = nt;
        pl->n_a = n_a;
        pl->n_b = n_b;
        pl->smem = (size_t)n_a * a_slot + (size_t)n_b * b_slot + fixed;
        pl->wbytes = (size_t)nt * pl->kchunks * pl->taps * b_slot;
        // CTA pair: same tiling, half-size weight slots (N = per must be a multiple of 16 for M = 256: it is)
        pl->pair_ok = pl->m_tiles >= 2 ? 1 : 0;
        {
            const int b_half = b_slot / 2;
            int na2 = 2, nb2 = (budget - na2 * a_slot) / b_half;
            if (nb2 > HL_MAX_B) nb2 = HL_MAX_B;
            if (nb2 >= 6 && budget - 3 * a_slot - 6 * b_half >= 0) {
                na2 = 3;
                nb2 = (budget - na2 * a_slot) / b_half;
                if (nb2 > HL_MAX_B) nb2 = HL_MAX_B;
            }
            if (nb2 < 2) pl->pair_ok = 0;
            pl->n_a2 = na2;
            pl->n_b2 = nb2;
            pl->smem2 = (size_t)na2 * a_slot + (size_t)nb2 * b_half + fixed;
        }
        return 0;
    }
    return 1;
}

}  // namespace

extern "C" int selavi_conv_halo_plan(const int* geom, int* m_tiles, int* bnt, int* ntiles, size_t* wpack_bytes) {
    if (!geom) return selavi_fail(-1, "conv_halo_plan: null geometry");
    HaloPlan pl;
    if (hl_plan(geom, &pl) != 0) return 1;
    if (m_tiles) *m_tiles = pl.m_tiles;
    if (bnt) *bnt = pl.bnt;
    if (ntiles) *ntiles = pl.ntiles;
    if (wpack_bytes) *wpack_bytes = pl.wbytes;
    return 0;
}

extern "C" int selavi_conv_halo_pack_weights(const float* W, const int* geom, int k_real, void* wpack, void* stream) {
    if (!W || !geom || !wpack || k_real <= 0) return selavi_fail(-1, "conv_halo_pack_weights: bad arguments");
    HaloPlan pl;
    if (hl_plan(geom, &pl) != 0) return selavi_fail(-1, "conv_halo_pack_weights: geometry not supported by the halo kernel");
    const size_t total = (size_t)pl.ntiles * pl.kchunks * pl.taps * pl.bnt * 64;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    // forward: W[co = n_out][ci = k_real];  data gradient: W[co = k_real][ci = n_out]
    const int co = geom[0] == 0 ? pl.co : k_real, ci = geom[0] == 0 ? k_real : pl.co;
    conv_halo_pack_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(W, geom[0], co, ci, pl.taps, pl.bnt, pl.ntiles, pl.kchunks,
                                                                    reinterpret_cast<uint16_t*>(wpack));
    SV_CUDA_CHECK(cudaGetLastError(), "conv_halo_pack_weights: launch");
    return 0;
}

static int hl_launch(const HaloPlan& pl, HaloParams& p, int flags, void* stream);

extern "C" int selavi_conv_halo_fwd(const float* src, float* dst, const void* wpack, const int* geom, const float* pro_scale,
                                    const float* pro_shift, int pro_relu, float* stats_partial, int flags, void* stream) {
    if (!src || !dst || !wpack || !geom) return selavi_fail(-1, "conv_halo_fwd: null argument");
    if (geom[0] != 0) return selavi_fail(-1, "conv_halo_fwd: geometry must be in forward mode");
    if ((pro_scale == nullptr) != (pro_shift == nullptr)) return selavi_fail(-1, "conv_halo_fwd: prologue needs scale and shift");
    HaloPlan pl;
    if (hl_plan(geom, &pl) != 0) return selavi_fail(-1, "conv_halo_fwd: geometry not supported by the halo kernel");
    HaloParams p;
    p.src = src; p.src_hi = nullptr; p.src_lo = nullptr; p.dst = dst; p.wpack = reinterpret_cast<const unsigned char*>(wpack);
    p.pro_scale = pro_scale; p.pro_shift = pro_shift; p.stats = stats_partial;
    p.pro_relu = pro_relu;
    p.src_kind = 0; p.accumulate = 0; p.oscale = HL_OSCALE;
    p.has_bwd = 0;
    p.bwd = sv::EpiBwdStat{nullptr, nullptr, nullptr, nullptr, nullptr};
    return hl_launch(pl, p, flags, stream);
}

static int hl_dgrad(const void* z_hi, const void* z_lo, float* dx, const void* wpack, const int* geom, int accumulate,
                    const sv::EpiBwdStat* bwd, float* stats_partial, int flags, void* stream);

extern "C" int selavi_conv_halo_dgrad(const void* z_hi, const void* z_lo, float* dx, const void* wpack, const int* geom,
                                      int accumulate, int flags, void* stream) {
    return hl_dgrad(z_hi, z_lo, dx, wpack, geom, accumulate, nullptr, nullptr, flags, stream);
}

// Same, and the epilogue also emits the per-tile partial sums of the BatchNorm-backward pass of the unit that produced the
// activation dx is the gradient of: stats_partial [m_tiles][2][ntiles*bnt] = (sum g*m, sum g*m*zhat) with g = dx,
// m = (bn_z*bn_scale+bn_shift > 0), zhat = (bn_z-bn_mean)*bn_invstd; bn_z [pixels_in, cis] fp32, the vectors [cis].
// Reduce with selavi_bn_reduce_partials; replaces selavi_bn_bwd_reduce (mask_mode 2) for that unit.  accumulate must be 0.
extern "C" int selavi_conv_halo_dgrad_bnstats(const void* z_hi, const void* z_lo, float* dx, const void* wpack, const int* geom,
                                              const float* bn_z, const float* bn_scale, const float* bn_shift,
                                              const float* bn_mean, const float* bn_invstd, float* stats_partial, int flags,
                                              void* stream) {
    if (!bn_z || !bn_scale || !bn_shift || !bn_mean || !bn_invstd || !stats_partial) return selavi_fail(-1, "conv_halo_dgrad_bnstats: null argument");
    const sv::EpiBwdStat bwd{bn_z, bn_scale, bn_shift, bn_mean, bn_invstd};
    return hl_dgrad(z_hi, z_lo, dx, wpack, geom, 0, &bwd, stats_partial, flags, stream);
}

static int hl_dgrad(const void* z_hi, const void* z_lo, float* dx, const void* wpack, const int* geom, int accumulate,
                    const sv::EpiBwdStat* bwd, float* stats_partial, int flags, void* stream) {
    if (!z_hi || !z_lo || !dx || !wpack || !geom) return selavi_fail(-1, "conv_halo_dgrad: null argument");
    if (geom[0] != 1) return selavi_fail(-1, "conv_halo_dgrad: geometry must be in dgrad mode");
    HaloPlan pl;
    if (hl_plan(geom, &pl) != 0) return selavi_fail(-1, "conv_halo_dgrad: geometry not supported by the halo kernel");
    HaloParams p;
    p.src = nullptr;
    p.src_hi = reinterpret_cast<const __nv_bfloat16*>(z_hi);
    p.src_lo = reinterpret_cast<const __nv_bfloat16*>(z_lo);
    p.dst = dx; p.wpack = reinterpret_cast<const unsigned char*>(wpack);
    p.pro_scale = nullptr; p.pro_shift = nullptr; p.stats = stats_partial;
    p.pro_relu = 0;
    p.src_kind = 1; p.accumulate = accumulate ? 1 : 0; p.oscale = 1.f;
    p.has_bwd = bwd != nullptr ? 1 : 0;
    p.bwd = bwd ? *bwd : sv::EpiBwdStat{nullptr, nullptr, nullptr, nullptr, nullptr};
    return hl_launch(pl, p, flags, stream);
}

static int hl_launch(const HaloPlan& pl, HaloParams& p, int flags, void* stream) {
    p.mode = pl.mode; p.nb = pl.nb; p.T = pl.T; p.H = pl.H; p.W = pl.W; p.cs = pl.cs; p.cd = pl.cd;
    p.S = pl.S; p.WP = pl.WP;
    p.wp_magic = (uint32_t)((0x100000000ULL + (unsigned)pl.WP - 1) / (unsigned)pl.WP);
    p.TB = pl.TB; p.SB = pl.SB; p.OB = pl.OB; p.rows = pl.rows; p.rows_alloc = pl.rows_alloc;
    p.m_tiles = pl.m_tiles; p.bnt = pl.bnt; p.ntiles = pl.ntiles; p.kchunks = pl.kchunks; p.taps = pl.taps;
    p.n_a = pl.n_a; p.n_b = pl.n_b;
    p.use_base_off = (flags & 1) ? 1 : 0;
    p.dbg = flags & (2 | 4 | 8);
    // flag 32: CTA pairs (cluster of 2, tcgen05 cta_group::2)
    const bool pair = (flags & 32) && pl.pair_ok;
    if (pair) {
        p.n_a = pl.n_a2;
        p.n_b = pl.n_b2;
    }
    uint32_t cols = 32;
    while ((int)cols < 2 * p.bnt) cols <<= 1;
    p.tmem_cols = cols;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (pair) {
        SV_CUDA_CHECK(cudaFuncSetAttribute(conv_halo_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem2),
                      "conv_halo: cudaFuncSetAttribute");
        const int nwork = ((p.m_tiles + 1) / 2) * p.ntiles;
        int npairs = sms / 2;
        if (npairs > nwork) npairs = nwork;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * npairs);
        cfg.blockDim = dim3(HL_THREADS);
        cfg.dynamicSmemBytes = pl.smem2;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        SV_CUDA_CHECK(cudaLaunchKernelEx(&cfg, conv_halo_kernel<true>, p), "conv_halo: pair launch");
        return 0;
    }
    SV_CUDA_CHECK(cudaFuncSetAttribute(conv_halo_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem),
                  "conv_halo: cudaFuncSetAttribute");
    const int total_tiles = p.m_tiles * p.ntiles;
    const int grid = total_tiles < sms ? total_tiles : sms;
    conv_halo_kernel<false><<<grid, HL_THREADS, pl.smem, (cudaStream_t)stream>>>(p);
    SV_CUDA_CHECK(cudaGetLastError(), "conv_halo: launch");
    return 0;
}
