// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 / TMEM), forward and data-gradient.
//
// Replaces the cuDNN calls behind torchvision's Conv2Plus1D / BasicBlock / stem / downsample convolutions
// (tv:video/resnet.py:45-61,184-195,276-281; tv:resnet.py:59-105) that the reference model is made of
// (model.py:93-121).  Activations are channels-last fp32 ([N,T,H,W,Cs], Cs = channels padded to 4).
//
//   dst[m, n] = sum_{tap, c} P(src)[pix(m, tap), c] * W[n, tap, c]        m = output pixel, n = output channel
//
// One CTA computes a 128-pixel x BNt-channel tile (BNt <= 256, fp32 accumulator in TMEM).
//   * A operand (pixels x K): gathered by 8 loader warps with 16-byte loads (im2col on the fly, zero padding,
//     stride, or the transposed gather of the data gradient), passed through an optional fused prologue
//     P(x) = relu(x*scale[c] + shift[c])  — the train-mode BatchNorm+ReLU of the PREVIOUS layer, so normalised
//     activations are never written to HBM — split into tf32 hi/lo words and stored into 128B-swizzled
//     K-major tiles in shared memory.
//   * B operand (weights): pre-tiled, pre-swizzled and pre-split on the device once per optimizer step
//     (conv_pack_weights), fetched with one 1-D TMA bulk copy per stage.
//   * MMA: one thread issues tcgen05.mma.kind::tf32; `passes`=3 computes hi*hi + hi*lo + lo*hi (fp32-class
//     accuracy, error ~2^-21; needed because train-mode BN amplifies operand rounding ~80x),
//     `passes`=1 is plain tf32.
//   * Epilogue: TMEM -> registers -> global (float4), plus per-tile per-channel sum / sum-of-squares partials
//     for the following BatchNorm (deterministic two-stage reduction, no atomics).
//
// F16 variant (passes = 6, forward of the strided / 7x7 / 1x1 convolutions): the same kernel with fp16 hi/lo operands
// (hi = fp16(16 x), lo = fp16(16 x - hi); weights scaled by 2^8, epilogue by 2^-12: all exact, 22 significant bits like
// tf32x3 and like the tap-reuse kernel conv_halo.cu) on tcgen05.mma.kind::f16: a 128-byte row holds 64 K elements
// instead of 32, i.e. half the pipeline stages, half the shared-memory bytes and half the MMA instructions.
#include <cuda_fp16.h>
#include <stdint.h>

#include "../../include/selavi_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int BM = 128;            // pixels per tile (UMMA M)
constexpr int BK = 32;             // fp32 elements per K stage (one 128B swizzle row); the fp16 variant holds 64
constexpr float F16_ASCALE = 16.f;
constexpr float F16_WSCALE = 256.f;
constexpr float F16_OSCALE = 1.f / (16.f * 256.f);
constexpr int EPI_WARPS = 4;       // warps 0..3  : epilogue (TMEM lane quadrant = warp id)
constexpr int LOADER_WARPS = 8;    // warps 4..19 : A-operand gather (4 per scheduler: the gather streams are latency-bound)
constexpr int ROWS_PER = 1024 / (LOADER_WARPS * 32);   // tile rows per loader thread
constexpr int ROW_STEP = LOADER_WARPS * 4;             // row distance between a thread's rows
constexpr int MMA_WARP = EPI_WARPS + LOADER_WARPS;
constexpr int BPROD_WARP = MMA_WARP + 1;
constexpr int CONV_THREADS = (BPROD_WARP + 1) * 32;
constexpr int A_TILE_BYTES = BM * 128;
constexpr int MAX_TAPS = 64;
constexpr int MAX_STAGES = 6;

struct ConvParams {
    const float* src;
    float* dst;
    const unsigned char* wpack;  // [ntiles][kstages][2][BNt][128B]
    const float* pro_scale;      // [cs] or null
    const float* pro_shift;
    float* stats;                // [m_tiles][2][ntiles*BNt] or null
    int mode;                    // 0 fwd, 1 dgrad
    int nb, ts, hs, ws, cs;      // src geometry
    int td, hd, wd, cd;          // dst geometry
    int kt, kh, kw, st, sh, sw, pt, ph, pw;
    int M;                       // nb*td*hd*wd
    int m_tiles;
    int kstages;                 // ceil(taps * cs/4 / 8)
    int bnt, ntiles, stages;
    int pro_relu, accumulate, passes;
    uint32_t tmem_cols;          // 2 accumulators of bnt columns, power of two
};

// hi = x rounded to tf32 (round half up on the magnitude), so that lo = x - hi is exact, |lo| <= 2^-12 |x| and unbiased
__device__ __forceinline__ uint32_t tf32_hi(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

// One gathered K stage of a loader thread: 4 rows x one 16-byte shared-memory chunk (4 fp32 channels, or 8 channels of the
// fp16 variant = NV float4 loads per row), plus the prologue affine of those channels.
template <int NV>
struct AStage {
    float4 v[ROWS_PER][NV];
    int ch;           // first channel of the chunk (the prologue affine is fetched at commit time: L1-resident, saves registers)
    uint32_t okmask;  // bit j: row j valid for this tap (prologue applies, otherwise exact zero)
};

__device__ __forceinline__ uint32_t h2_bits(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

// Persistent kernel: CTA b processes tiles b, b+gridDim.x, ... (tile = m_tile * ntiles + n_tile).
template <bool F16>
__global__ void __launch_bounds__(CONV_THREADS, 1) conv_igemm_kernel(const ConvParams p) {
    constexpr int NV = F16 ? 2 : 1;      // float4 loads per (row, chunk)
    constexpr int CW = 4 * NV;           // channels per 16-byte shared-memory chunk
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // carve: [stages x (A_hi | A_lo | B_hi | B_lo)] [barriers] [tap tables] [stat scratch]
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int b_tile_bytes = p.bnt * 128;
    const int stage_bytes = 2 * A_TILE_BYTES + 2 * b_tile_bytes;
    unsigned char* tail = smem + (size_t)p.stages * stage_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);        // [MAX_STAGES]
    uint64_t* empty_bar = full_bar + MAX_STAGES;                    // [MAX_STAGES]
    uint64_t* tfull_bar = empty_bar + MAX_STAGES;                   // [2] accumulator ready
    uint64_t* tempty_bar = tfull_bar + 2;                           // [2] accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    int* tap_dt = reinterpret_cast<int*>(tmem_slot + 2);            // [MAX_TAPS] packed (kt | kh<<8 | kw<<16)
    int* tap_off = tap_dt + MAX_TAPS;                               // [MAX_TAPS] pixel offset of the tap
    float* s_stat = reinterpret_cast<float*>(tap_off + MAX_TAPS);   // [EPI_WARPS][2][256]
    unsigned char* s_stage = reinterpret_cast<unsigned char*>(         // [EPI_WARPS] x (32 rows x 128 B | 32 row indices)
        (reinterpret_cast<uintptr_t>(s_stat + EPI_WARPS * 512) + 127) & ~(uintptr_t)127);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int taps = p.kt * p.kh * p.kw;
    const int total_tiles = p.m_tiles * p.ntiles;

    if (tid == 0) {
        for (int s = 0; s < p.stages; ++s) {
            sv::mbar_init(&full_bar[s], LOADER_WARPS + 1);
            sv::mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            sv::mbar_init(&tfull_bar[a], 1);
            sv::mbar_init(&tempty_bar[a], EPI_WARPS);
        }
        sv::fence_barrier_init();
    }
    for (int t = tid; t < taps; t += CONV_THREADS) {
        const int kw_ = t % p.kw, kh_ = (t / p.kw) % p.kh, kt_ = t / (p.kw * p.kh);
        tap_dt[t] = kt_ | (kh_ << 8) | (kw_ << 16);
        if (p.mode == 0) {
            tap_off[t] = (kt_ * p.hs + kh_) * p.ws + kw_;
        } else {
            // transposed gather: src = (dst + pad - k) / stride  ==  floor((dst+pad)/stride) - (k >> log2(stride)) when valid
            const int qt = p.st == 2 ? (kt_ >> 1) : kt_, qh = p.sh == 2 ? (kh_ >> 1) : kh_, qw = p.sw == 2 ? (kw_ >> 1) : kw_;
            tap_off[t] = -((qt * p.hs + qh) * p.ws + qw);
        }
    }
    if (warp == MMA_WARP) {
        sv::tmem_alloc(tmem_slot, p.tmem_cols);
        sv::tmem_relinquish();
    }
    sv::tc_fence_before();
    __syncthreads();
    sv::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= EPI_WARPS && warp < MMA_WARP) {
        // ------------------------------------------------------------------ A loaders (256 threads)
        const int ltid = tid - EPI_WARPS * 32;
        const int c = ltid & 7;    // 16-byte chunk within the 128B K row
        const int r0 = ltid >> 3;  // rows r0 + ROW_STEP*j
        const int C4 = p.cs / CW;      // chunk units per pixel (cs is a multiple of 8)
        const uint32_t sw_off = (uint32_t)((c ^ (r0 & 7)) << 4);
        const bool pro = p.pro_scale != nullptr;
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int m0 = (tile / p.ntiles) * BM;
            // per-row gather base (pixel index of tap 0) and per-dimension tap validity bits
            int pb[ROWS_PER];
            uint32_t vm[ROWS_PER];
#pragma unroll
            for (int j = 0; j < ROWS_PER; ++j) {
                const int m = m0 + r0 + ROW_STEP * j;
                pb[j] = 0;
                vm[j] = 0;
                if (m < p.M) {
                    const int w_ = m % p.wd;
                    const int t1 = m / p.wd;
                    const int h_ = t1 % p.hd;
                    const int t2 = t1 / p.hd;
                    const int t_ = t2 % p.td;
                    const int n_ = t2 / p.td;
                    int bt, bh, bw;
                    uint32_t mt = 0, mh = 0, mw = 0;
                    if (p.mode == 0) {
                        bt = t_ * p.st - p.pt;
                        bh = h_ * p.sh - p.ph;
                        bw = w_ * p.sw - p.pw;
                        for (int k = 0; k < p.kt; ++k) mt |= (uint32_t)((bt + k >= 0) & (bt + k < p.ts)) << k;
                        for (int k = 0; k < p.kh; ++k) mh |= (uint32_t)((bh + k >= 0) & (bh + k < p.hs)) << k;
                        for (int k = 0; k < p.kw; ++k) mw |= (uint32_t)((bw + k >= 0) & (bw + k < p.ws)) << k;
                    } else {
                        const int at = t_ + p.pt, ah = h_ + p.ph, aw = w_ + p.pw;
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
                        bt = p.st == 2 ? (at >> 1) : at;
                        bh = p.sh == 2 ? (ah >> 1) : ah;
                        bw = p.sw == 2 ? (aw >> 1) : aw;
                    }
                    pb[j] = ((n_ * p.ts + bt) * p.hs + bh) * p.ws + bw;
                    vm[j] = mt | (mh << 8) | (mw << 16) | 0x80000000u;
                }
            }
            int tap = 0, c4 = c;  // flattened K chunk q = 8*ks + c  ->  (tap, c4)
            while (c4 >= C4) {
                c4 -= C4;
                ++tap;
            }
            // issue the gathers of one K stage into registers (no dependence on the smem ring)
            auto gather = [&](AStage<NV>& a) {
                a.okmask = 0;
                a.ch = 0;
#pragma unroll
                for (int j = 0; j < ROWS_PER; ++j)
#pragma unroll
                    for (int v = 0; v < NV; ++v) a.v[j][v] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (tap < taps) {
                    const int pk = tap_dt[tap];
                    const int off = tap_off[tap];
                    const int s_t = pk & 255, s_h = 8 + ((pk >> 8) & 255), s_w = 16 + ((pk >> 16) & 255);
#pragma unroll
                    for (int j = 0; j < ROWS_PER; ++j) {
                        const uint32_t m_ = vm[j];
                        if ((m_ >> 31) & (m_ >> s_t) & (m_ >> s_h) & (m_ >> s_w) & 1u) {
                            const float4* g = reinterpret_cast<const float4*>(p.src + (size_t)(pb[j] + off) * p.cs + c4 * CW);
#pragma unroll
                            for (int v = 0; v < NV; ++v) a.v[j][v] = __ldg(g + v);
                            a.okmask |= 1u << j;
                        }
                    }
                    a.ch = c4 * CW;
                }
                c4 += 8;  // advance the K position by 8 chunks
                while (c4 >= C4 && tap < taps) {
                    c4 -= C4;
                    ++tap;
                }
            };
            // prologue + hi/lo split + swizzled store of one gathered stage, then signal the MMA warp
            auto commit = [&](const AStage<NV>& a) {
                sv::mbar_wait(&empty_bar[stage], phase ^ 1);
                const uint32_t a_hi = sv::smem_u32(smem + (size_t)stage * stage_bytes);
                const uint32_t a_lo = a_hi + A_TILE_BYTES;
#pragma unroll
                for (int j = 0; j < ROWS_PER; ++j) {
                    float x[CW];
#pragma unroll
                    for (int v = 0; v < NV; ++v) {
                        float4 t = a.v[j][v];
                        if (pro && ((a.okmask >> j) & 1u)) {
                            const float4 sc = __ldg(reinterpret_cast<const float4*>(p.pro_scale + a.ch) + v);
                            const float4 sf = __ldg(reinterpret_cast<const float4*>(p.pro_shift + a.ch) + v);
                            t.x = fmaf(t.x, sc.x, sf.x);
                            t.y = fmaf(t.y, sc.y, sf.y);
                            t.z = fmaf(t.z, sc.z, sf.z);
                            t.w = fmaf(t.w, sc.w, sf.w);
                            if (p.pro_relu) {
                                t.x = fmaxf(t.x, 0.f);
                                t.y = fmaxf(t.y, 0.f);
                                t.z = fmaxf(t.z, 0.f);
                                t.w = fmaxf(t.w, 0.f);
                            }
                        }
                        x[4 * v] = t.x;
                        x[4 * v + 1] = t.y;
                        x[4 * v + 2] = t.z;
                        x[4 * v + 3] = t.w;
                    }
                    const uint32_t row_off = (uint32_t)((r0 + ROW_STEP * j) * 128) + sw_off;
                    if constexpr (F16) {
                        uint32_t hb[4], lb[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float x0 = x[2 * e] * F16_ASCALE, x1 = x[2 * e + 1] * F16_ASCALE;   // exact (power of two)
                            const __half2 h = __floats2half2_rn(x0, x1);
                            const float2 hf = __half22float2(h);
                            hb[e] = h2_bits(h);
                            lb[e] = h2_bits(__floats2half2_rn(x0 - hf.x, x1 - hf.y));
                        }
                        st_shared_v4(a_hi + row_off, hb[0], hb[1], hb[2], hb[3]);
                        st_shared_v4(a_lo + row_off, lb[0], lb[1], lb[2], lb[3]);
                    } else {
                        const uint32_t h0 = tf32_hi(x[0]), h1 = tf32_hi(x[1]), h2 = tf32_hi(x[2]), h3 = tf32_hi(x[3]);
                        st_shared_v4(a_hi + row_off, h0, h1, h2, h3);
                        if (p.passes == 3) {
                            st_shared_v4(a_lo + row_off, __float_as_uint(x[0] - __uint_as_float(h0)),
                                         __float_as_uint(x[1] - __uint_as_float(h1)), __float_as_uint(x[2] - __uint_as_float(h2)),
                                         __float_as_uint(x[3] - __uint_as_float(h3)));
                        }
                    }
                }
                sv::fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) sv::mbar_arrive(&full_bar[stage]);
                if (++stage == p.stages) {
                    stage = 0;
                    phase ^= 1;
                }
            };
            // software pipeline: the gathers of stage ks+1 are in flight while stage ks is converted and stored
            AStage<NV> sa, sb;
            gather(sa);
            for (int ks = 0; ks < p.kstages; ks += 2) {
                if (ks + 1 < p.kstages) gather(sb);
                commit(sa);
                if (ks + 1 < p.kstages) {
                    if (ks + 2 < p.kstages) gather(sa);
                    commit(sb);
                }
            }
        }
    } else if (warp < EPI_WARPS) {
        // ------------------------------------------------------------------ epilogue (4 warps, one TMEM quadrant each)
        const int quad = warp;
        const int ctot = p.ntiles * p.bnt;
        float* my_stat = s_stat + (size_t)warp * 512;
        unsigned char* my_stage = s_stage + (size_t)warp * sv::EPI_STAGE_BYTES;
        const uint32_t stg = sv::smem_u32(my_stage);
        int* row_pix = reinterpret_cast<int*>(my_stage + 32 * 128);
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const int m_tile = tile / p.ntiles, ntile = tile % p.ntiles;
            const int m = m_tile * BM + quad * 32 + lane;
            const bool row_ok = m < p.M;
            const int n_base = ntile * p.bnt;
            row_pix[lane] = row_ok ? m : -1;
            __syncwarp();
            sv::mbar_wait(&tfull_bar[acc], (uint32_t)((it >> 1) & 1));
            sv::tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc * (p.tmem_cols >> 1);
            // coalesced drain through the per-warp staging tile (+ the BN column sums), see sv::epi_drain_group
            sv::epi_drain_tile(taddr, p.bnt, F16 ? F16_OSCALE : 1.f, row_ok, stg, row_pix, p.dst, p.cd, n_base, p.accumulate, true,
                               p.stats != nullptr ? my_stat : nullptr, lane);
            // accumulator drained: hand the TMEM buffer back to the MMA warp
            sv::tc_fence_before();
            __syncwarp();
            if (lane == 0) sv::mbar_arrive(&tempty_bar[acc]);
            if (p.stats != nullptr) {
                named_bar_sync(1, EPI_WARPS * 32);
                for (int i = tid; i < 2 * p.bnt; i += EPI_WARPS * 32) {
                    const int which = i >= p.bnt, col = which ? i - p.bnt : i;
                    float tsum = 0.f;
#pragma unroll
                    for (int q = 0; q < EPI_WARPS; ++q) tsum += s_stat[(size_t)q * 512 + which * 256 + col];
                    p.stats[((size_t)m_tile * 2 + which) * ctot + n_base + col] = tsum;
                }
                named_bar_sync(1, EPI_WARPS * 32);
            }
        }
    } else if (warp == MMA_WARP) {
        // ------------------------------------------------------------------ MMA issuer
        // the warp stays converged (warp-uniform schedule => descriptors in uniform registers), one elected lane issues;
        // a divergent `if (lane == 0)` makes the compiler wrap each tcgen05.mma in an elect/broadcast/branch loop
        {
            const uint32_t idesc = F16 ? sv::make_idesc_f16(BM, p.bnt, 0, 0, 0, 0) : sv::make_idesc_tf32(BM, p.bnt, 0, 0);
            const uint32_t tm0 = __shfl_sync(0xffffffffu, tmem_base, 0);
            const uint64_t desc_fixed = sv::make_smem_desc_sw128(0, 16, 1024);
            int stage = 0, it = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                sv::mbar_wait(&tempty_bar[acc], (uint32_t)(((it >> 1) & 1) ^ 1));
                sv::tc_fence_after();
                const uint32_t d_tmem = tm0 + (uint32_t)acc * (p.tmem_cols >> 1);
                for (int ks = 0; ks < p.kstages; ++ks) {
                    sv::mbar_wait(&full_bar[stage], phase);
                    sv::tc_fence_after();
                    const uint32_t a_hi = sv::smem_u32(smem + (size_t)stage * stage_bytes);
                    // k-step k4 adds 32 bytes = 2 to the descriptor's 16-byte-unit address field
                    const uint64_t da_hi = desc_fixed | (uint64_t)((a_hi & 0x3FFFFu) >> 4);
                    const uint64_t da_lo = da_hi + (uint64_t)(A_TILE_BYTES >> 4);
                    const uint64_t db_hi = da_lo + (uint64_t)(A_TILE_BYTES >> 4);
                    const uint64_t db_lo = db_hi + (uint64_t)(b_tile_bytes >> 4);
                    if (sv::elect_one()) {
                        if constexpr (F16) {   // 4 k16 steps of 32 bytes, three MMAs each (lo*hi + hi*lo + hi*hi)
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4) {
                                sv::umma_f16(d_tmem, da_lo + 2 * k4, db_hi + 2 * k4, idesc, (uint32_t)(ks | k4));
                                sv::umma_f16(d_tmem, da_hi + 2 * k4, db_lo + 2 * k4, idesc, 1u);
                                sv::umma_f16(d_tmem, da_hi + 2 * k4, db_hi + 2 * k4, idesc, 1u);
                            }
                        } else if (p.passes == 3) {
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4) {
                                sv::umma_tf32(d_tmem, da_lo + 2 * k4, db_hi + 2 * k4, idesc, (uint32_t)(ks | k4));
                                sv::umma_tf32(d_tmem, da_hi + 2 * k4, db_lo + 2 * k4, idesc, 1u);
                                sv::umma_tf32(d_tmem, da_hi + 2 * k4, db_hi + 2 * k4, idesc, 1u);
                            }
                        } else {
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4)
                                sv::umma_tf32(d_tmem, da_hi + 2 * k4, db_hi + 2 * k4, idesc, (uint32_t)(ks | k4));
                        }
                        sv::umma_commit(&empty_bar[stage]);  // smem slot free once these MMAs retire
                    }
                    __syncwarp();
                    if (++stage == p.stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (sv::elect_one()) sv::umma_commit(&tfull_bar[acc]);  // accumulator complete
                __syncwarp();
            }
        }
    } else {
        // ------------------------------------------------------------------ B producer (one thread, TMA bulk)
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)((F16 || p.passes == 3) ? 2 * b_tile_bytes : b_tile_bytes);
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int ntile = tile % p.ntiles;
                const unsigned char* wsrc = p.wpack + (size_t)ntile * p.kstages * 2 * b_tile_bytes;
                for (int ks = 0; ks < p.kstages; ++ks) {
                    sv::mbar_wait(&empty_bar[stage], phase ^ 1);
                    sv::mbar_arrive_expect_tx(&full_bar[stage], bytes);
                    sv::bulk_g2s(smem + (size_t)stage * stage_bytes + 2 * A_TILE_BYTES,
                                 wsrc + (size_t)ks * 2 * b_tile_bytes, bytes, &full_bar[stage]);
                    if (++stage == p.stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    }
    __syncthreads();
    if (warp == MMA_WARP) {
        sv::tc_fence_after();
        sv::tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// W (torch layout [co][ci][taps]) -> pre-tiled, pre-swizzled, hi/lo split B operand.
//   mode 0 (fwd):   n = co, k = tap*cs + ci   (cs = padded ci)
//   mode 1 (dgrad): n = ci, k = tap*cs + co   (cs = padded co)
__global__ void conv_pack_weights_kernel(const float* __restrict__ W, int mode, int co, int ci, int taps, int cs,
                                         int bnt, int ntiles, int kstages, float* __restrict__ out) {
    const size_t total = (size_t)ntiles * kstages * bnt * 32;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int e = (int)(idx & 3);
        const int c = (int)((idx >> 2) & 7);
        const int n = (int)((idx >> 5) % bnt);
        const size_t blk = (idx >> 5) / bnt;  // ntile*kstages + ks
        const int ks = (int)(blk % kstages);
        const int nt = (int)(blk / kstages);
        const int k = ks * 32 + c * 4 + e;
        const int tap = k / cs, kc = k % cs;
        const int nn = nt * bnt + n;
        float val = 0.f;
        if (tap < taps) {
            if (mode == 0) {
                if (nn < co && kc < ci) val = W[((size_t)nn * ci + kc) * taps + tap];
            } else {
                if (nn < ci && kc < co) val = W[((size_t)kc * ci + nn) * taps + tap];
            }
        }
        const float hi = __uint_as_float((__float_as_uint(val) + 0x1000u) & 0xFFFFE000u);
        const float lo = val - hi;
        float* base = out + blk * (size_t)(2 * bnt * 32);
        const int pos = n * 32 + ((c ^ (n & 7)) << 2) + e;
        base[pos] = hi;
        base[bnt * 32 + pos] = lo;
    }
}

// fp16x3 forward variant: [ntile][kstage of 64 k][hi|lo][bnt rows][128 B], values scaled by 2^8; k = tap*cs + ci
__global__ void conv_pack_weights_f16_kernel(const float* __restrict__ W, int co, int ci, int taps, int cs, int bnt, int ntiles,
                                             int kstages, uint16_t* __restrict__ out) {
    const size_t total = (size_t)ntiles * kstages * bnt * 64;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int e = (int)(idx & 7);
        const int c = (int)((idx >> 3) & 7);
        const int n = (int)((idx >> 6) % bnt);
        const size_t blk = (idx >> 6) / bnt;  // ntile*kstages + ks
        const int ks = (int)(blk % kstages);
        const int nt = (int)(blk / kstages);
        const int k = ks * 64 + c * 8 + e;
        const int tap = k / cs, kc = k % cs;
        const int nn = nt * bnt + n;
        float val = 0.f;
        if (tap < taps && nn < co && kc < ci) val = W[((size_t)nn * ci + kc) * taps + tap] * F16_WSCALE;
        const __half h = __float2half_rn(val);
        const __half l = __float2half_rn(val - __half2float(h));
        uint16_t* base = out + blk * (size_t)(2 * bnt * 64);
        const int pos = n * 64 + ((c ^ (n & 7)) << 3) + e;
        base[pos] = __half_as_ushort(h);
        base[bnt * 64 + pos] = __half_as_ushort(l);
    }
}

void pick_tiles(int n_out, int* bnt, int* ntiles) {
    int nt = (n_out + 255) / 256;
    int per = (n_out + nt - 1) / nt;
    per = (per + 15) & ~15;
    *bnt = per;
    *ntiles = nt;
}

}  // namespace

extern "C" int selavi_conv_tiles(int n_out, int* bnt, int* ntiles) {
    if (n_out <= 0 || !bnt || !ntiles) return selavi_fail(-1, "conv_tiles: bad arguments");
    pick_tiles(n_out, bnt, ntiles);
    return 0;
}

extern "C" size_t selavi_conv_wpack_bytes(int n_out, int k_total) {
    int bnt, nt;
    pick_tiles(n_out, &bnt, &nt);
    const int kstages = (k_total + BK - 1) / BK;
    return (size_t)nt * kstages * 2 * bnt * 128;
}

extern "C" size_t selavi_conv_wpack_bytes_f16(int n_out, int k_total) {
    int bnt, nt;
    pick_tiles(n_out, &bnt, &nt);
    const int kstages = (k_total + 2 * BK - 1) / (2 * BK);
    return (size_t)nt * kstages * 2 * bnt * 128;
}

extern "C" int selavi_conv_pack_weights(const float* W, int mode, int co, int ci, int taps, int cs, void* wpack,
                                        void* stream) {
    if (!W || !wpack || co <= 0 || ci <= 0 || taps <= 0 || (cs & 3)) return selavi_fail(-1, "conv_pack_weights: bad arguments");
    int bnt, nt;
    if (mode == 2) {   // forward, fp16x3 operands (selavi_conv_gemm passes = 6)
        if (cs & 7) return selavi_fail(-1, "conv_pack_weights: the fp16x3 variant needs channel strides in multiples of 8");
        pick_tiles(co, &bnt, &nt);
        const int kstages = (taps * cs + 2 * BK - 1) / (2 * BK);
        const size_t total = (size_t)nt * kstages * bnt * 64;
        int blocks = (int)((total + 255) / 256);
        if (blocks > 148 * 16) blocks = 148 * 16;
        conv_pack_weights_f16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(W, co, ci, taps, cs, bnt, nt, kstages,
                                                                               reinterpret_cast<uint16_t*>(wpack));
        SV_CUDA_CHECK(cudaGetLastError(), "conv_pack_weights: launch");
        return 0;
    }
    if (mode != 0 && mode != 1) return selavi_fail(-1, "conv_pack_weights: mode must be 0, 1 or 2");
    pick_tiles(mode == 0 ? co : ci, &bnt, &nt);
    const int kstages = (taps * cs + BK - 1) / BK;
    const size_t total = (size_t)nt * kstages * bnt * 32;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    conv_pack_weights_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(W, mode, co, ci, taps, cs, bnt, nt, kstages,
                                                                       reinterpret_cast<float*>(wpack));
    SV_CUDA_CHECK(cudaGetLastError(), "conv_pack_weights: launch");
    return 0;
}

// geom: [mode, nb, ts, hs, ws, cs, td, hd, wd, cd, kt, kh, kw, st, sh, sw, pt, ph, pw, n_out]
extern "C" int selavi_conv_gemm(const float* src, float* dst, const void* wpack, const int* geom, const float* pro_scale,
                                const float* pro_shift, int pro_relu, float* stats_partial, int accumulate, int passes,
                                void* stream) {
    if (!src || !dst || !wpack || !geom) return selavi_fail(-1, "conv_gemm: null argument");
    ConvParams p;
    p.src = src;
    p.dst = dst;
    p.wpack = reinterpret_cast<const unsigned char*>(wpack);
    p.pro_scale = pro_scale;
    p.pro_shift = pro_shift;
    p.stats = stats_partial;
    p.mode = geom[0];
    p.nb = geom[1]; p.ts = geom[2]; p.hs = geom[3]; p.ws = geom[4]; p.cs = geom[5];
    p.td = geom[6]; p.hd = geom[7]; p.wd = geom[8]; p.cd = geom[9];
    p.kt = geom[10]; p.kh = geom[11]; p.kw = geom[12];
    p.st = geom[13]; p.sh = geom[14]; p.sw = geom[15];
    p.pt = geom[16]; p.ph = geom[17]; p.pw = geom[18];
    const int n_out = geom[19];
    if ((p.cs & 3) || (p.cd & 3) || p.cs <= 0 || p.cd <= 0) return selavi_fail(-1, "conv_gemm: channel strides must be multiples of 4");
    if (p.kt * p.kh * p.kw > MAX_TAPS) return selavi_fail(-1, "conv_gemm: too many taps");
    if (p.kt > 8 || p.kh > 8 || p.kw > 8) return selavi_fail(-1, "conv_gemm: kernel extent above 8 not supported");
    if ((p.st != 1 && p.st != 2) || (p.sh != 1 && p.sh != 2) || (p.sw != 1 && p.sw != 2)) return selavi_fail(-1, "conv_gemm: stride must be 1 or 2");
    if (passes != 1 && passes != 3 && passes != 6) return selavi_fail(-1, "conv_gemm: passes must be 1, 3 (tf32) or 6 (fp16x3)");
    const bool f16 = passes == 6;
    if (f16 && (p.mode != 0 || (p.cs & 7))) return selavi_fail(-1, "conv_gemm: the fp16x3 variant is forward only, channel stride multiple of 8");
    if ((pro_scale == nullptr) != (pro_shift == nullptr)) return selavi_fail(-1, "conv_gemm: prologue needs scale and shift");
    const long long M = (long long)p.nb * p.td * p.hd * p.wd;
    if (M <= 0 || M > 0x7fffffffLL || (long long)p.nb * p.ts * p.hs * p.ws > 0x7fffffffLL) return selavi_fail(-1, "conv_gemm: bad pixel count");
    p.M = (int)M;
    pick_tiles(n_out, &p.bnt, &p.ntiles);
    // padded destination channels [n_out, cd) are written as exact zeros by the (zero) weight tile rows
    if (p.ntiles * p.bnt < p.cd) return selavi_fail(-1, "conv_gemm: cd exceeds the tiled channel range");
    p.kstages = f16 ? (p.kt * p.kh * p.kw * (p.cs >> 3) + 7) / 8 : (p.kt * p.kh * p.kw * (p.cs >> 2) + 7) / 8;
    p.m_tiles = (p.M + BM - 1) / BM;
    p.pro_relu = pro_relu;
    p.accumulate = accumulate;
    p.passes = f16 ? 3 : passes;
    uint32_t cols = 32;
    while ((int)cols < 2 * p.bnt) cols <<= 1;  // two accumulators (epilogue of tile i overlaps the MMAs of tile i+1)
    p.tmem_cols = cols;
    const int stage_bytes = 2 * A_TILE_BYTES + 2 * p.bnt * 128;
    const int tail_bytes = (2 * MAX_STAGES + 4) * 8 + 8 + 2 * MAX_TAPS * 4 + EPI_WARPS * 512 * 4 + EPI_WARPS * sv::EPI_STAGE_BYTES + 128 + 64;
    int stages = (227 * 1024 - 1024 - tail_bytes) / stage_bytes;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages > p.kstages) stages = p.kstages < 1 ? 1 : p.kstages;
    if (stages < 2 && p.kstages > 1) return selavi_fail(-1, "conv_gemm: tile does not fit shared memory");
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + tail_bytes + 1024;
    SV_CUDA_CHECK(f16 ? cudaFuncSetAttribute(conv_igemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                      : cudaFuncSetAttribute(conv_igemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                  "conv_gemm: cudaFuncSetAttribute");
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int total_tiles = p.m_tiles * p.ntiles;
    const int grid = total_tiles < sms ? total_tiles : sms;  // persistent: one CTA per SM
    if (f16) conv_igemm_kernel<true><<<grid, CONV_THREADS, smem, (cudaStream_t)stream>>>(p);
    else conv_igemm_kernel<false><<<grid, CONV_THREADS, smem, (cudaStream_t)stream>>>(p);
    SV_CUDA_CHECK(cudaGetLastError(), "conv_gemm: launch");
    return 0;
}
