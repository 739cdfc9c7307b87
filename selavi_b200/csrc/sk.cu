// Sinkhorn-Knopp pseudo-label assignment, one persistent cooperative kernel (sm_100a).
//
// Replaces the solver of the reference `optimize_L_sk_gpu` (src/sk_utils.py:359-422):
//   PS.pow_(lamb/2); while err>0.1 and it<2000: alpha = r/(beta^T PS)^T; beta' = c/(PS alpha); ...
//   newL = argmax(beta*PS*alpha, 1); cost = -(1/lamb) * nansum(log PS[n,newL[n]]) / N
//
// B200 design (HBM-bound, 0.25 flop/B): the reference streams the N x K float64 matrix twice per
// iteration (two DGEMVs).  Here one iteration reads every element ONCE: a warp that has a row in
// shared memory computes rowdot_n = sum_k PS[n,k]*alpha[k], beta'_n = c/rowdot_n and immediately
// accumulates the NEXT iteration's column sums colsum_k += beta'_n*PS[n,k] in registers.  Rows are
// staged global->shared with 1-D TMA bulk copies (cp.async.bulk + mbarrier tx-count pipeline); the
// cross-CTA reduction of the K column sums goes through a [grid][Ks] partial buffer and ONE grid
// barrier per iteration, summed by every CTA in a fixed order (bit-reproducible run to run).
// Odd iterations walk the CTA's row range backwards so the tail of the previous pass is still in
// the 126 MB L2.  Convergence (err = sum |beta/beta' - 1| every `check_every` iterations) is decided
// on-device with the reference's exact stopping rule; no host sync inside the loop.
//
// Multi-GPU (rows sharded over ranks): `peer_sum[r]` point at every rank's receive buffer (NVSwitch
// P2P-mapped memory); once per iteration CTA 0 of every rank PUSHES its reduced [colsum.., err] vector
// into its slot of every rank's buffer (posted stores over NVLink) and raises a flag there; all CTAs poll
// and sum LOCAL memory in rank order (same summation order everywhere => identical alpha on all ranks).
#include <math.h>
#include <stdint.h>

#include "../../include/selavi_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int SK_WARPS = 15;                     // consumer warps
constexpr int SK_THREADS = (SK_WARPS + 1) * 32;  // + 1 producer warp
constexpr int SK_STAGES = 4;
constexpr int SK_MAX_WORLD = 8;
constexpr int SK_GMAX = 160;                     // >= SM count

struct SkArgs {
    double* PS;          // [N_local, K] row-major, mutated in place (pow)
    long long n_local;   // rows on this rank
    long long n_global;  // rows over all ranks (c = 1/n_global, beta0 = 1/n_global)
    int K;
    int rows_per_stage;  // even
    double pow_exp;      // 0.5*lamb
    int use_dist;        // 0: 'default' (uniform r); 1: argsort matching with kdist
    double* kdist;       // [K] in: target sizes; out: permuted (reference mutates args.dist[hc] in place)
    double* r;           // [Kp] workspace (persists between calls for do_prep=0)
    double* alpha;       // [K] out
    double* beta;        // [N_local] out
    long long* labels;   // [N_local] out
    double* part;        // [2][G][Ks]  per-CTA partial column sums (+ misc group), double-buffered
    double* part_raw;    // [G][Ks]     column sums of the un-powered matrix (argsort key)
    unsigned* bar;       // grid barrier counter (zeroed by the host wrapper)
    int* state;          // [0] = parity of the part buffer holding the next iteration's column sums
    int max_iters;
    int check_every;
    double tol;
    int stop_on_converge;  // 1: reference rule; 0: run exactly max_iters (cfg-5 microbench)
    int do_prep;           // 1: pow + initial sums; 0: PS already powered, continue from workspace state
    int do_final;          // 1: labels + cost
    int* iters_out;
    double* err_out;
    double* cost_out;  // local nansum(log PS[n,L_n]) contribution (host divides)
    int world;
    int rank;
    double* peer_sum[SK_MAX_WORLD];     // [2][world][Ks]  receive slots of each rank (slot r written by rank r)
    unsigned* peer_flag[SK_MAX_WORLD];  // [world] epoch flags of each rank (zeroed by the host before the call)
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];\n" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_cg_f64(const double* p) {
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];\n" : "=d"(v) : "l"(p));
    return v;
}

// Monotonic-counter grid barrier (kernel is launched cooperatively, all CTAs are co-resident).
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned& epoch) {
    // generic-proxy global writes of this pass (PS in the prep pass) are later read by TMA (async proxy)
    asm volatile("fence.proxy.async;\n" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += 1;
        __threadfence();
        atomicAdd(ctr, 1u);
        const unsigned target = epoch * gridDim.x;
        while (ld_acquire_u32(ctr) < target) {
        }
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

enum { MODE_PREP = 0, MODE_ITER = 1, MODE_FINAL = 2 };

// KPL = ceil(K/32) column groups per lane; group index KPL is the "misc" group (lane 0: err, lane 1: cost)
template <int KPL>
__global__ void __launch_bounds__(SK_THREADS, 1) sk_kernel(const SkArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int Kp = KPL * 32;
    constexpr int Ks = Kp + 32;
    const int K = a.K;
    const int R = a.rows_per_stage;
    const size_t stage_bytes = (size_t)R * K * sizeof(double);
    const size_t stage_stride = (stage_bytes + 127) & ~(size_t)127;
    double* s_alpha = reinterpret_cast<double*>(smem_raw);            // [Ks]
    double* s_vec = s_alpha + Ks;                                     // [Ks]  reduced colsum (+misc)
    double* s_r = s_vec + Ks;                                         // [Ks]
    double* s_red = s_r + Ks;                                         // [SK_WARPS+1][Ks]
    double* s_misc = s_red + (SK_WARPS + 1) * Ks;                     // [8]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_misc + 8);     // [STAGES]
    uint64_t* empty_bar = full_bar + SK_STAGES;                       // [STAGES]
    unsigned char* s_rows = reinterpret_cast<unsigned char*>(empty_bar + SK_STAGES);
    s_rows = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(s_rows) + 127) & ~(uintptr_t)127);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x, cta = blockIdx.x;
    const bool is_producer = (warp == SK_WARPS);

    if (tid == 0) {
        for (int s = 0; s < SK_STAGES; ++s) {
            sv::mbar_init(&full_bar[s], 1);
            sv::mbar_init(&empty_bar[s], SK_WARPS);
        }
        sv::fence_barrier_init();
    }
    __syncthreads();

    // row range of this CTA (even-sized chunks so every bulk copy source is 16-byte aligned)
    long long rows_per_cta = (a.n_local + G - 1) / G;
    rows_per_cta = (rows_per_cta + 1) & ~1LL;
    const long long row_begin = (long long)cta * rows_per_cta < a.n_local ? (long long)cta * rows_per_cta : a.n_local;
    const long long row_end = row_begin + rows_per_cta < a.n_local ? row_begin + rows_per_cta : a.n_local;
    const long long my_rows = row_end - row_begin;
    const int n_chunks = (int)((my_rows + R - 1) / R);

    const double cN = 1.0 / (double)a.n_global;
    unsigned epoch = 0;    // grid barrier epoch
    unsigned xepoch = 0;   // cross-GPU exchange epoch
    uint32_t pipe_it = 0;  // running chunk counter, identical in producer and consumers

    double acc[KPL];      // per-lane partial column sums (consumer warps)
    double alpha_r[KPL];  // alpha (iterate/final passes); raw column sums during the prep pass

    // Sum the G per-CTA partial vectors (fixed order) into dst[Ks] (smem). All 16 warps take part.
    auto reduce_partials = [&](const double* part /*[G][Ks]*/, double* dst) {
        double v[KPL + 1];
#pragma unroll
        for (int i = 0; i <= KPL; ++i) v[i] = 0.0;
#pragma unroll 2
        for (int j = warp; j < G; j += SK_WARPS + 1) {
            double t[KPL + 1];
#pragma unroll
            for (int i = 0; i <= KPL; ++i) t[i] = ld_cg_f64(part + (size_t)j * Ks + lane + 32 * i);
#pragma unroll
            for (int i = 0; i <= KPL; ++i) v[i] += t[i];
        }
#pragma unroll
        for (int i = 0; i <= KPL; ++i) s_red[warp * Ks + lane + 32 * i] = v[i];
        __syncthreads();
        for (int k = tid; k < Ks; k += SK_THREADS) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w <= SK_WARPS; ++w) s += s_red[w * Ks + k];
            dst[k] = s;
        }
        __syncthreads();
    };

    // Cross-GPU all-reduce of vec[Ks] (smem), summed in rank order on every CTA of every rank (identical alpha
    // everywhere).  PUSH model: CTA 0 of rank r stores its vector into slot r of EVERY rank's receive buffer over
    // NVSwitch (posted stores, one-way latency) and then raises flag r there; consumers poll and read LOCAL memory only.
    auto cross_gpu_sum = [&](double* vec) {
        if (a.world <= 1) return;
        xepoch += 1;
        const int buf = xepoch & 1;
        if (cta == 0) {
            for (int idx = tid; idx < a.world * Ks; idx += SK_THREADS) {
                const int pr = idx / Ks, k = idx - pr * Ks;
                a.peer_sum[pr][(size_t)(buf * a.world + a.rank) * Ks + k] = vec[k];
            }
            __threadfence_system();
            __syncthreads();
            if (tid < a.world) st_release_sys_u32(a.peer_flag[tid] + a.rank, xepoch);
        }
        if (tid < a.world) {
            while (ld_acquire_sys_u32(a.peer_flag[a.rank] + tid) < xepoch) {
            }
        }
        __syncthreads();
        const double* mine = a.peer_sum[a.rank] + (size_t)buf * a.world * Ks;
        for (int k = tid; k < Ks; k += SK_THREADS) {
            double s = 0.0;
            for (int r = 0; r < a.world; ++r) s += ld_volatile_f64(mine + (size_t)r * Ks + k);
            vec[k] = s;
        }
        __syncthreads();
    };

    // One pass over this CTA's rows in the given mode; publishes the per-CTA partial vector.
    auto row_pass = [&](const int mode, const bool check, const bool backwards, double* part_dst,
                        double* part_raw_dst) {
#pragma unroll
        for (int i = 0; i < KPL; ++i) {
            acc[i] = 0.0;
            if (mode == MODE_PREP) alpha_r[i] = 0.0;
        }
        double err_acc = 0.0, cost_acc = 0.0;
        if (is_producer) {
            if (lane == 0) {
                for (int c = 0; c < n_chunks; ++c) {
                    const uint32_t g = pipe_it + c;
                    const int s = g % SK_STAGES;
                    const uint32_t ph = (g / SK_STAGES) & 1;
                    sv::mbar_wait(&empty_bar[s], ph ^ 1);
                    const int cc = backwards ? (n_chunks - 1 - c) : c;
                    const long long r0 = row_begin + (long long)cc * R;
                    const long long nr = (row_end - r0) < R ? (row_end - r0) : R;
                    const uint32_t bytes = (uint32_t)(nr * K * sizeof(double));
                    // an odd element count leaves bytes % 16 == 8: consumers fetch the last double by hand
                    const uint32_t bulk = bytes & ~15u;
                    sv::mbar_arrive_expect_tx(&full_bar[s], bulk);
                    if (bulk) sv::bulk_g2s(s_rows + s * stage_stride, a.PS + (size_t)r0 * K, bulk, &full_bar[s]);
                }
            }
        } else {
            for (int c = 0; c < n_chunks; ++c) {
                const uint32_t g = pipe_it + c;
                const int s = g % SK_STAGES;
                const uint32_t ph = (g / SK_STAGES) & 1;
                const int cc = backwards ? (n_chunks - 1 - c) : c;
                const long long r0 = row_begin + (long long)cc * R;
                const int nr = (int)((row_end - r0) < R ? (row_end - r0) : R);
                const bool tail8 = (((size_t)nr * K) & 1) != 0;  // last double not covered by the bulk copy
                sv::mbar_wait(&full_bar[s], ph);
                const double* rows = reinterpret_cast<const double*>(s_rows + s * stage_stride);
                // rows are dealt round-robin over the warps, continuing across chunks (balanced for any R)
                const int first = (warp + SK_WARPS - (int)(((long long)cc * R) % SK_WARPS)) % SK_WARPS;
                for (int rr = first; rr < nr; rr += SK_WARPS) {
                    const long long n = r0 + rr;
                    double x[KPL];
                    const double* rowp = rows + rr * K;
                    // (ncu: the kernel issues 278 instructions per row and half of all issue slots are busy — the odd-tail
                    // test used to sit inside the element loop as a branch around a global load; it concerns one row of
                    // one chunk of the last CTA)
                    if (!(tail8 && rr == nr - 1)) {
#pragma unroll
                        for (int i = 0; i < KPL; ++i) {
                            const int k = lane + 32 * i;
                            x[i] = (k < K) ? rowp[k] : 0.0;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < KPL; ++i) {
                            const int k = lane + 32 * i;
                            double v = 0.0;
                            if (k < K) v = (k == K - 1) ? a.PS[(size_t)n * K + k] : rowp[k];
                            x[i] = v;
                        }
                    }
                    if (mode == MODE_PREP) {
#pragma unroll
                        for (int i = 0; i < KPL; ++i) {
                            const int k = lane + 32 * i;
                            if (k < K) {
                                alpha_r[i] += x[i];
                                const double p = pow(x[i], a.pow_exp);
                                a.PS[(size_t)n * K + k] = p;
                                acc[i] = fma(cN, p, acc[i]);
                            }
                        }
                        if (lane == 0) a.beta[n] = cN;
                    } else if (mode == MODE_ITER) {
                        double d0 = 0.0, d1 = 0.0;
#pragma unroll
                        for (int i = 0; i < KPL; i += 2) {
                            d0 = fma(x[i], alpha_r[i], d0);
                            if (i + 1 < KPL) d1 = fma(x[i + 1], alpha_r[i + 1], d1);
                        }
                        const double dot = warp_sum(d0 + d1);
                        const double bnew = cN / dot;
                        if (lane == 0) {
                            if (check) err_acc += fabs(a.beta[n] / bnew - 1.0);
                            a.beta[n] = bnew;
                        }
#pragma unroll
                        for (int i = 0; i < KPL; ++i) acc[i] = fma(bnew, x[i], acc[i]);
                    } else {  // MODE_FINAL: labels + cost term, same operation order as the reference
                        const double b = a.beta[n];
                        double best = -1.0;
                        int bestk = 0x7fffffff;
#pragma unroll
                        for (int i = 0; i < KPL; ++i) {
                            const int k = lane + 32 * i;
                            if (k < K) {
                                const double v = (x[i] * b) * alpha_r[i];
                                if (v > best) {
                                    best = v;
                                    bestk = k;
                                }
                            }
                        }
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                            const int ok = __shfl_xor_sync(0xffffffffu, bestk, o);
                            if (ob > best || (ob == best && ok < bestk)) {
                                best = ob;
                                bestk = ok;
                            }
                        }
                        if (lane == 0) {
                            if (bestk == 0x7fffffff) bestk = 0;
                            a.labels[n] = bestk;
                            // PS[n,L] after the reference's "return back" multiplications (sk_utils.py:416-417)
                            const double al = s_alpha[bestk];
                            const double back = ((1.0 / al) * best) * (1.0 / b);
                            const double lg = log(back);
                            if (lg == lg) cost_acc += lg;  // nansum
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) sv::mbar_arrive(&empty_bar[s]);
            }
        }
        pipe_it += n_chunks;

        // publish this CTA's partial vector: [colsum(Kp) | misc(32): err, cost]
        if (!is_producer) {
#pragma unroll
            for (int i = 0; i < KPL; ++i) s_red[warp * Ks + lane + 32 * i] = acc[i];
            const double cost0 = __shfl_sync(0xffffffffu, cost_acc, 0);  // err/cost are accumulated by lane 0
            s_red[warp * Ks + Kp + lane] = (lane == 0) ? err_acc : (lane == 1 ? cost0 : 0.0);
        }
        __syncthreads();
        for (int k = tid; k < Ks; k += SK_THREADS) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < SK_WARPS; ++w) s += s_red[w * Ks + k];
            part_dst[(size_t)cta * Ks + k] = s;
        }
        if (mode == MODE_PREP) {
            __syncthreads();
            if (!is_producer) {
#pragma unroll
                for (int i = 0; i < KPL; ++i) s_red[warp * Ks + lane + 32 * i] = alpha_r[i];
            }
            __syncthreads();
            for (int k = tid; k < Kp; k += SK_THREADS) {
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < SK_WARPS; ++w) s += s_red[w * Ks + k];
                part_raw_dst[(size_t)cta * Ks + k] = s;
            }
            for (int k = Kp + tid; k < Ks; k += SK_THREADS) part_raw_dst[(size_t)cta * Ks + k] = 0.0;
        }
        grid_barrier(a.bar, epoch);
    };

    int pbase = a.do_prep ? 0 : a.state[0];
    if (a.do_prep) {
        row_pass(MODE_PREP, false, false, a.part + (size_t)pbase * G * Ks, a.part_raw);
        // marginals r (src/sk_utils.py:366-395); every CTA computes them redundantly (deterministic)
        if (a.use_dist) {
            double* s_raw = s_red;              // [Ks] column sums of raw PS
            double* s_kd = s_red + 4 * Ks;      // [Kp] kdist
            reduce_partials(a.part_raw, s_vec);
            cross_gpu_sum(s_vec);
            for (int k = tid; k < K; k += SK_THREADS) {
                s_raw[k] = s_vec[k];
                s_kd[k] = a.kdist[k];
            }
            grid_barrier(a.bar, epoch);  // everyone has read kdist before CTA 0 overwrites it below
            // sk_utils.py:388  _K_dist[argsort(colsum)] = torch.sort(_K_dist)[0]; _K_dist is [K,1] so the sort
            // runs over the size-1 last dim (identity): new[argsort[i]] = old[i]  <=>  new[k] = old[rank(k)].
            for (int k = tid; k < K; k += SK_THREADS) {
                const double v = s_raw[k];
                int rk = 0;
                for (int j = 0; j < K; ++j) rk += (s_raw[j] < v) || (s_raw[j] == v && j < k);
                s_r[k] = s_kd[rk];
            }
            __syncthreads();
            if (tid == 0) {
                double s = 0.0;
                for (int k = 0; k < K; ++k) s += 1.0 / s_r[k];
                s_misc[0] = s;
            }
            __syncthreads();
            const double rs = s_misc[0];
            for (int k = tid; k < K; k += SK_THREADS) {
                const double kd = s_r[k];
                if (cta == 0) a.kdist[k] = kd;
                s_r[k] = (1.0 / kd) / rs;
            }
        } else {
            // _K_dist = ones -> r = 1/K (r = 1./ones; r /= r.sum())
            for (int k = tid; k < K; k += SK_THREADS) s_r[k] = 1.0 / (double)K;
        }
        __syncthreads();
        if (cta == 0)
            for (int k = tid; k < K; k += SK_THREADS) a.r[k] = s_r[k];
    } else {
        for (int k = tid; k < K; k += SK_THREADS) s_r[k] = a.r[k];
        for (int k = tid; k < K; k += SK_THREADS) s_alpha[k] = a.alpha[k];
    }
    for (int k = K + tid; k < Ks; k += SK_THREADS) {
        s_r[k] = 0.0;
        s_alpha[k] = 0.0;
    }
    __syncthreads();

    int it = 0;
    double err = 1e6;
    while (true) {
        // column sums for iteration `it` (+ err partial of iteration it-1 in the misc group)
        reduce_partials(a.part + (size_t)((pbase + it) & 1) * G * Ks, s_vec);
        cross_gpu_sum(s_vec);
        if (it > 0 && ((it - 1) % a.check_every == 0)) err = s_vec[Kp];
        const bool cont = (it < a.max_iters) && (!a.stop_on_converge || err > a.tol);
        if (!cont) break;
        for (int k = tid; k < Kp; k += SK_THREADS) {
            const double al = (k < K) ? s_r[k] / s_vec[k] : 0.0;
            s_alpha[k] = al;
            if (cta == 0 && k < K) a.alpha[k] = al;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < KPL; ++i) alpha_r[i] = s_alpha[lane + 32 * i];
        row_pass(MODE_ITER, (it % a.check_every) == 0, (it & 1) != 0,
                 a.part + (size_t)((pbase + it + 1) & 1) * G * Ks, nullptr);
        it += 1;
    }
    if (a.do_final) {
#pragma unroll
        for (int i = 0; i < KPL; ++i) alpha_r[i] = s_alpha[lane + 32 * i];
        // the partial buffer NOT holding the live column sums is free scratch
        double* scratch = a.part + (size_t)((pbase + it + 1) & 1) * G * Ks;
        row_pass(MODE_FINAL, false, false, scratch, nullptr);
        if (cta == 0 && tid == 0) {
            double cst = 0.0;
            for (int j = 0; j < G; ++j) cst += ld_cg_f64(scratch + (size_t)j * Ks + Kp + 1);
            *a.cost_out = cst;
        }
    }
    if (cta == 0 && tid == 0) {
        *a.iters_out = it;
        *a.err_out = err;
        a.state[0] = (pbase + it) & 1;
    }
}

template <int KPL>
size_t sk_smem_bytes(int K, int R) {
    constexpr int Ks = KPL * 32 + 32;
    size_t stage_bytes = (size_t)R * K * sizeof(double);
    size_t stage_stride = (stage_bytes + 127) & ~(size_t)127;
    return (size_t)(3 * Ks + (SK_WARPS + 1) * Ks + 8) * sizeof(double) + 2 * SK_STAGES * sizeof(uint64_t) + 128 +
           SK_STAGES * stage_stride;
}

template <int KPL>
int launch_sk(SkArgs& a, int grid, cudaStream_t stream) {
    // rows per stage: as many as fit in 224 KB, even, at most 30
    int R = 30;
    while (R > 2 && sk_smem_bytes<KPL>(a.K, R) > (size_t)224 * 1024) R -= 2;
    a.rows_per_stage = R;
    const size_t smem = sk_smem_bytes<KPL>(a.K, R);
    if (smem > (size_t)227 * 1024) return selavi_fail(-4, "sk: K too large for the shared-memory pipeline");
    cudaError_t e = cudaFuncSetAttribute(sk_kernel<KPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return selavi_cuda_fail(e, "sk: cudaFuncSetAttribute");
    void* args[] = {(void*)&a};
    e = cudaLaunchCooperativeKernel((void*)sk_kernel<KPL>, dim3(grid), dim3(SK_THREADS), args, smem, stream);
    if (e != cudaSuccess) return selavi_cuda_fail(e, "sk: cudaLaunchCooperativeKernel");
    return 0;
}

const int kKPLs[] = {1, 2, 4, 8, 10, 13, 16};

int pick_kpl(int K) {
    for (int v : kKPLs)
        if (v * 32 >= K) return v;
    return -1;
}

}  // namespace

namespace {
// PS[n,k] = softmax64(v[n,:])[k] * softmax64(a[n,:])[k]   (src/sk_utils.py:206-211,309-315: f64 softmax of both
// modalities' head outputs, then PS_v *= PS_a) — one warp per row, fp32 logits in, float64 out, one pass.
__global__ void sk_softmax_product_kernel(const float* __restrict__ v, const float* __restrict__ a, long long n, int K,
                                          double* __restrict__ PS) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const float* vr = v + (size_t)row * K;
    const float* ar = a + (size_t)row * K;
    double mv = -INFINITY, ma = -INFINITY;
    for (int k = lane; k < K; k += 32) {
        mv = fmax(mv, (double)vr[k]);
        ma = fmax(ma, (double)ar[k]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mv = fmax(mv, __shfl_xor_sync(0xffffffffu, mv, o));
        ma = fmax(ma, __shfl_xor_sync(0xffffffffu, ma, o));
    }
    double sv_ = 0.0, sa_ = 0.0;
    for (int k = lane; k < K; k += 32) {
        sv_ += exp((double)vr[k] - mv);
        sa_ += exp((double)ar[k] - ma);
    }
    sv_ = warp_sum(sv_);
    sa_ = warp_sum(sa_);
    double* out = PS + (size_t)row * K;
    for (int k = lane; k < K; k += 32) out[k] = (exp((double)vr[k] - mv) / sv_) * (exp((double)ar[k] - ma) / sa_);
}
}  // namespace

extern "C" int selavi_sk_softmax_product(const float* logits_v, const float* logits_a, long long n, int K, double* PS,
                                         void* stream) {
    if (!logits_v || !logits_a || !PS || n <= 0 || K <= 0) return selavi_fail(-1, "sk_softmax_product: bad arguments");
    const int warps = 8;
    const long long blocks = (n + warps - 1) / warps;
    sk_softmax_product_kernel<<<(unsigned)blocks, warps * 32, 0, (cudaStream_t)stream>>>(logits_v, logits_a, n, K, PS);
    SV_CUDA_CHECK(cudaGetLastError(), "sk_softmax_product: launch");
    return 0;
}

extern "C" int selavi_sk_kp(int K) {
    const int kpl = pick_kpl(K);
    return kpl < 0 ? -1 : kpl * 32 + 32;
}

extern "C" size_t selavi_sk_workspace_bytes(int K) {
    const int Ks = selavi_sk_kp(K);
    if (Ks < 0) return 0;
    // part[2][G][Ks] + part_raw[G][Ks] + r[Ks] + bar/state (256 B)
    return (size_t)(3 * SK_GMAX * Ks + Ks) * sizeof(double) + 256;
}

extern "C" int selavi_sk_solve(double* PS, long long n_local, long long n_global, int K, double lamb, int use_dist,
                               double* kdist, double* alpha_out, double* beta_out, long long* labels_out,
                               void* workspace, int max_iters, int check_every, double tol, int stop_on_converge,
                               int do_prep, int do_final, int* iters_out, double* err_out, double* cost_sum_out,
                               int world, int rank, void* const* peer_sum, void* const* peer_flag,
                               void* stream_) {
    const int KPL = pick_kpl(K);
    if (K <= 0 || KPL < 0) return selavi_fail(-1, "sk: K must be in [1, 512]");
    if (n_local <= 0 || n_global < n_local) return selavi_fail(-1, "sk: bad row counts");
    if (world < 1 || world > SK_MAX_WORLD || rank < 0 || rank >= world) return selavi_fail(-2, "sk: bad world/rank");
    if (max_iters < 1 || check_every < 1) return selavi_fail(-1, "sk: max_iters and check_every must be >= 1");
    if (use_dist && !kdist) return selavi_fail(-1, "sk: use_dist needs kdist");
    if (world > 1 && (!peer_sum || !peer_flag)) return selavi_fail(-2, "sk: world > 1 needs peer buffers");
    cudaStream_t stream = (cudaStream_t)stream_;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int grid = sms < SK_GMAX ? sms : SK_GMAX;
    // few rows: do not launch CTAs that would own nothing (the grid must be the same on every rank only
    // in the sense that each rank reduces its own partials; ranks may differ)
    long long max_useful = (n_local + 1) / 2;
    if (grid > max_useful) grid = (int)max_useful;
    if (grid < 1) grid = 1;

    const int Ks = KPL * 32 + 32;
    SkArgs a;
    a.PS = PS;
    a.n_local = n_local;
    a.n_global = n_global;
    a.K = K;
    a.rows_per_stage = 2;
    a.pow_exp = 0.5 * lamb;
    a.use_dist = use_dist;
    a.kdist = kdist;
    a.alpha = alpha_out;
    a.beta = beta_out;
    a.labels = labels_out;
    double* w = reinterpret_cast<double*>(workspace);
    a.part = w;
    a.part_raw = w + (size_t)2 * SK_GMAX * Ks;
    a.r = a.part_raw + (size_t)SK_GMAX * Ks;
    a.bar = reinterpret_cast<unsigned*>(a.r + Ks);
    a.state = reinterpret_cast<int*>(a.bar) + 4;
    a.max_iters = max_iters;
    a.check_every = check_every;
    a.tol = tol;
    a.stop_on_converge = stop_on_converge;
    a.do_prep = do_prep;
    a.do_final = do_final;
    a.iters_out = iters_out;
    a.err_out = err_out;
    a.cost_out = cost_sum_out;
    a.world = world;
    a.rank = rank;
    for (int r = 0; r < SK_MAX_WORLD; ++r) {
        a.peer_sum[r] = (world > 1 && r < world) ? reinterpret_cast<double*>(peer_sum[r]) : nullptr;
        a.peer_flag[r] = (world > 1 && r < world) ? reinterpret_cast<unsigned*>(peer_flag[r]) : nullptr;
    }
    cudaError_t e = cudaMemsetAsync(a.bar, 0, 16, stream);
    if (e != cudaSuccess) return selavi_cuda_fail(e, "sk: memset");
    switch (KPL) {
        case 1: return launch_sk<1>(a, grid, stream);
        case 2: return launch_sk<2>(a, grid, stream);
        case 4: return launch_sk<4>(a, grid, stream);
        case 8: return launch_sk<8>(a, grid, stream);
        case 10: return launch_sk<10>(a, grid, stream);
        case 13: return launch_sk<13>(a, grid, stream);
        default: return launch_sk<16>(a, grid, stream);
    }
}
