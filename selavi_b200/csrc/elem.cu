// HBM-bound elementwise / reduction kernels around the convolutions: train-mode BatchNorm statistics,
// normalise(+residual)(+ReLU), BatchNorm backward, pooling, layout conversion, fused SGD.
// All activations are channels-last fp32 [M, Cs] (Cs multiple of 4), accessed as float4.
//
// Reference semantics: nn.BatchNorm3d/2d/1d in train mode (biased batch variance for normalisation, unbiased
// for running_var, momentum 0.1, eps 1e-5; SyncBN path torch:nn/modules/_functions.py:39-200), the
// BasicBlock residual add + ReLU (tv:video/resnet.py:107-119, tv:resnet.py:89-105), MaxPool2d(3,2,1)
// (tv:resnet.py:271), AdaptiveAvgPool (tv:video/resnet.py:230), torch.optim.SGD (main.py:132-137).
#include <cuda_bf16.h>
#include <stdint.h>

#include <mutex>
#include <vector>

#include "../../include/selavi_b200.h"
#include "common.cuh"

namespace {

constexpr int EW_THREADS = 256;

inline int ew_blocks(long long n) {
    long long b = (n + EW_THREADS - 1) / EW_THREADS;
    const long long cap = 148LL * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ---------------------------------------------------------------- stats: partial tiles -> fp64 sums
// partial [tiles][2][ctot] fp32 -> sums [2][cs] fp64 (fixed order; channels >= ctot are zero).
// grid (ceil(cs/32), 2, nchunk): every z-chunk reduces a contiguous range of tiles into scratch[chunk][2][cs]; the block
// that finishes last (atomic ticket) adds the chunks in chunk order => deterministic, and ~nchunk x more parallel than
// a single pass (12 544 tiles for the layer-1 convolutions).
constexpr int RP_MAX_CHUNKS = 16;
constexpr int RP_MAX_CS = 2048;

__global__ void bn_reduce_partials_kernel(const float* __restrict__ partial, int tiles, int ctot, int cs,
                                          double* __restrict__ sums, double* __restrict__ scratch, unsigned* __restrict__ tickets) {
    __shared__ double red[32][33];
    __shared__ bool is_last;
    const int c = blockIdx.x * 32 + threadIdx.x;
    const int which = blockIdx.y;
    const int nchunk = gridDim.z, chunk = blockIdx.z;
    const int per = (tiles + nchunk - 1) / nchunk;
    const int t0 = chunk * per, t1 = min(tiles, t0 + per);
    double s = 0.0;
    if (c < ctot && c < cs) {
        for (int t = t0 + threadIdx.y; t < t1; t += 32) s += (double)partial[((size_t)t * 2 + which) * ctot + c];
    }
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < cs) {
        double tot = 0.0;
        for (int i = 0; i < 32; ++i) tot += red[i][threadIdx.x];
        if (nchunk == 1) sums[(size_t)which * cs + c] = tot;
        else scratch[((size_t)chunk * 2 + which) * RP_MAX_CS + c] = tot;
    }
    if (nchunk == 1) return;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        const unsigned prev = atomicAdd(&tickets[blockIdx.y * gridDim.x + blockIdx.x], 1u);
        is_last = (prev == (unsigned)nchunk - 1);
        if (is_last) tickets[blockIdx.y * gridDim.x + blockIdx.x] = 0;   // re-arm for the next launch
    }
    __syncthreads();
    if (is_last && threadIdx.y == 0 && c < cs) {
        __threadfence();
        double tot = 0.0;
        for (int k = 0; k < nchunk; ++k) tot += scratch[((size_t)k * 2 + which) * RP_MAX_CS + c];
        sums[(size_t)which * cs + c] = tot;
    }
}

// sums [2][cs] (sum, sum of squares over `count` elements) -> scale/shift/mean/invstd, running stats update
__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* running_mean, float* running_var,
                                   float momentum, float eps, int c_real, int cs, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ mean_out,
                                   float* __restrict__ invstd_out, int update_running) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cs) return;
    if (c >= c_real) {
        scale[c] = 0.f;
        shift[c] = 0.f;
        mean_out[c] = 0.f;
        invstd_out[c] = 0.f;
        return;
    }
    const double mean = sums[c] / count;
    double var = sums[cs + c] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    const float sc = g * invstd;
    scale[c] = sc;
    shift[c] = b - (float)mean * sc;
    mean_out[c] = (float)mean;
    invstd_out[c] = invstd;
    if (update_running) {
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}

__global__ void bn_eval_affine_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                      float eps, int c_real, int cs, float* __restrict__ scale, float* __restrict__ shift) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cs) return;
    if (c >= c_real) {
        scale[c] = 0.f;
        shift[c] = 0.f;
        return;
    }
    const float invstd = 1.f / sqrtf(running_var[c] + eps);
    const float sc = (gamma ? gamma[c] : 1.f) * invstd;
    scale[c] = sc;
    shift[c] = (beta ? beta[c] : 0.f) - running_mean[c] * sc;
}

__device__ __forceinline__ float4 affine4(float4 x, float4 s, float4 b) {
    return make_float4(fmaf(x.x, s.x, b.x), fmaf(x.y, s.y, b.y), fmaf(x.z, s.z, b.z), fmaf(x.w, s.w, b.w));
}
__device__ __forceinline__ float4 relu4(float4 x) {
    return make_float4(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f), fmaxf(x.z, 0.f), fmaxf(x.w, 0.f));
}

// out = act( z*scale+shift  [+ res  |  + res*rscale+rshift] )
__global__ void bn_apply_kernel(const float4* __restrict__ z, const float4* __restrict__ scale,
                                const float4* __restrict__ shift, const float4* __restrict__ res,
                                const float4* __restrict__ rscale, const float4* __restrict__ rshift, int relu,
                                float4* __restrict__ out, long long total4, int c4n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % c4n);
        float4 v = affine4(z[i], __ldg(scale + c4), __ldg(shift + c4));
        if (res) {
            float4 r = res[i];
            if (rscale) r = affine4(r, __ldg(rscale + c4), __ldg(rshift + c4));
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        }
        if (relu) v = relu4(v);
        out[i] = v;
    }
}

// masked upstream gradient: mode 0: g; 1: g * (act > 0); 2: g * (z*scale+shift > 0)
__device__ __forceinline__ float4 masked_g(float4 g, int mode, float4 a) {
    if (mode == 0) return g;
    return make_float4(a.x > 0.f ? g.x : 0.f, a.y > 0.f ? g.y : 0.f, a.z > 0.f ? g.z : 0.f, a.w > 0.f ? g.w : 0.f);
}

// per-block partial sums of  g  and  g * zhat  (zhat = (z-mean)*invstd).  grid (row blocks, ceil(c4n/32)), 256 thr.
__global__ void bn_bwd_reduce_kernel(const float4* __restrict__ g, const float4* __restrict__ z,
                                     const float4* __restrict__ act, int mask_mode, const float4* __restrict__ scale,
                                     const float4* __restrict__ shift, const float4* __restrict__ mean,
                                     const float4* __restrict__ invstd, long long M, int c4n,
                                     float* __restrict__ partial /*[gridDim.x][2][c4n*4]*/) {
    __shared__ float4 red[2][8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c4 = blockIdx.y * 32 + lane;
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
    if (c4 < c4n) {
        const float4 mu = __ldg(mean + c4), is = __ldg(invstd + c4);
        float4 sc = make_float4(0.f, 0.f, 0.f, 0.f), sh = sc;
        if (mask_mode == 2) {
            sc = __ldg(scale + c4);
            sh = __ldg(shift + c4);
        }
        const long long rows_per_block = (M + gridDim.x - 1) / gridDim.x;
        const long long r0 = (long long)blockIdx.x * rows_per_block;
        long long r1 = r0 + rows_per_block;
        if (r1 > M) r1 = M;
        // 4 rows per iteration: 8-12 independent 16-byte loads in flight per thread (HBM-latency bound otherwise)
        for (long long r = r0 + warp; r < r1; r += 32) {
            float4 zz[4], gg[4], aa[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long rr = r + 8 * u;
                const bool ok = rr < r1;
                const size_t i = (size_t)(ok ? rr : r) * c4n + c4;
                zz[u] = z[i];
                gg[u] = g[i];
                aa[u] = (mask_mode == 1) ? act[i] : zz[u];
                if (!ok) gg[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float4 a = aa[u];
                if (mask_mode == 2) a = affine4(zz[u], sc, sh);
                const float4 gm = masked_g(gg[u], mask_mode, a);
                s1.x += gm.x; s1.y += gm.y; s1.z += gm.z; s1.w += gm.w;
                s2.x = fmaf(gm.x, (zz[u].x - mu.x) * is.x, s2.x);
                s2.y = fmaf(gm.y, (zz[u].y - mu.y) * is.y, s2.y);
                s2.z = fmaf(gm.z, (zz[u].z - mu.z) * is.z, s2.z);
                s2.w = fmaf(gm.w, (zz[u].w - mu.w) * is.w, s2.w);
            }
        }
    }
    red[0][warp][lane] = s1;
    red[1][warp][lane] = s2;
    __syncthreads();
    if (warp < 2 && c4 < c4n) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int w = 0; w < 8; ++w) {
            const float4 v = red[warp][w][lane];
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        float* dst = partial + ((size_t)blockIdx.x * 2 + warp) * (c4n * 4) + c4 * 4;
        *reinterpret_cast<float4*>(dst) = t;
    }
}

// Same sums for narrow tensors (c4n <= 128 float4 per row): the block is R complete rows wide (R * c4n threads, thread t
// owns channel group t % c4n of row slot t / c4n), so every lane is busy and consecutive threads read consecutive
// addresses across row boundaries.  The 32-lanes-per-row mapping above leaves 4 of 32 lanes active on the second
// column block of a 144-channel tensor (c4n = 36) and half of them on a 64-channel one.
__global__ void bn_bwd_reduce_rows_kernel(const float4* __restrict__ g, const float4* __restrict__ z,
                                          const float4* __restrict__ act, int mask_mode, const float4* __restrict__ scale,
                                          const float4* __restrict__ shift, const float4* __restrict__ mean,
                                          const float4* __restrict__ invstd, long long M, int c4n, int R,
                                          float* __restrict__ partial /*[gridDim.x][2][c4n*4]*/) {
    extern __shared__ float4 red_rows[];   // [2][R][c4n]
    const int t = threadIdx.x;
    const int rs = t / c4n, c4 = t - rs * c4n;
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
    const float4 mu = __ldg(mean + c4), is = __ldg(invstd + c4);
    float4 sc = make_float4(0.f, 0.f, 0.f, 0.f), sh = sc;
    if (mask_mode == 2) {
        sc = __ldg(scale + c4);
        sh = __ldg(shift + c4);
    }
    const long long rows_per_block = (M + gridDim.x - 1) / gridDim.x;
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    long long r1 = r0 + rows_per_block;
    if (r1 > M) r1 = M;
    for (long long r = r0 + rs; r < r1; r += 4 * R) {   // 4 row groups per iteration: 8-12 independent 16-byte loads in flight
        float4 zz[4], gg[4], aa[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long rr = r + (long long)R * u;
            const bool ok = rr < r1;
            const size_t i = (size_t)(ok ? rr : r) * c4n + c4;
            zz[u] = z[i];
            gg[u] = g[i];
            aa[u] = (mask_mode == 1) ? act[i] : zz[u];
            if (!ok) gg[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float4 a = aa[u];
            if (mask_mode == 2) a = affine4(zz[u], sc, sh);
            const float4 gm = masked_g(gg[u], mask_mode, a);
            s1.x += gm.x; s1.y += gm.y; s1.z += gm.z; s1.w += gm.w;
            s2.x = fmaf(gm.x, (zz[u].x - mu.x) * is.x, s2.x);
            s2.y = fmaf(gm.y, (zz[u].y - mu.y) * is.y, s2.y);
            s2.z = fmaf(gm.z, (zz[u].z - mu.z) * is.z, s2.z);
            s2.w = fmaf(gm.w, (zz[u].w - mu.w) * is.w, s2.w);
        }
    }
    red_rows[t] = s1;
    red_rows[R * c4n + t] = s2;
    __syncthreads();
    if (t < 2 * c4n) {
        const int which = t / c4n, c = t - which * c4n;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < R; ++k) {
            const float4 v = red_rows[(which * R + k) * c4n + c];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        float* dst = partial + ((size_t)blockIdx.x * 2 + which) * (c4n * 4) + c * 4;
        *reinterpret_cast<float4*>(dst) = acc;
    }
}

// dz = scale * (g - sum_g/count - zhat * sum_gz/count);   optional gres (+)= g (masked) for the residual branch
__global__ void bn_bwd_apply_kernel(const float4* __restrict__ g, const float4* __restrict__ z,
                                    const float4* __restrict__ act, int mask_mode, const float4* __restrict__ scale,
                                    const float4* __restrict__ shift, const float4* __restrict__ mean,
                                    const float4* __restrict__ invstd, const double* __restrict__ sums, double count,
                                    long long total4, int c4n, float4* __restrict__ dz, float4* __restrict__ gres,
                                    int gres_accumulate, uint2* __restrict__ dz_hi, uint2* __restrict__ dz_lo,
                                    uint2* __restrict__ a_hi, uint2* __restrict__ a_lo) {
    // dz = A*g + B*z + C per channel, A = scale, B = -scale*invstd*m2, C = -scale*m1 - B*mean (m1, m2 = sums / count);
    // the coefficient table lives in shared memory, the element loop is a flat grid-stride stream (HBM-bound)
    extern __shared__ float4 s_coef[];   // [3][c4n]: A | B | C   (+ [c4n] shift for mask_mode 2)
    const int cs = c4n * 4;
    const double inv = 1.0 / count;
    for (int c4 = threadIdx.x; c4 < c4n; c4 += blockDim.x) {
        const float4 sc = __ldg(scale + c4), mu = __ldg(mean + c4), is = __ldg(invstd + c4);
        const int c = c4 * 4;
        const float m1x = (float)(sums[c] * inv), m1y = (float)(sums[c + 1] * inv), m1z = (float)(sums[c + 2] * inv),
                    m1w = (float)(sums[c + 3] * inv);
        const float m2x = (float)(sums[cs + c] * inv), m2y = (float)(sums[cs + c + 1] * inv),
                    m2z = (float)(sums[cs + c + 2] * inv), m2w = (float)(sums[cs + c + 3] * inv);
        const float4 B = make_float4(-sc.x * is.x * m2x, -sc.y * is.y * m2y, -sc.z * is.z * m2z, -sc.w * is.w * m2w);
        s_coef[c4] = sc;
        s_coef[c4n + c4] = B;
        s_coef[2 * c4n + c4] = make_float4(-sc.x * m1x - B.x * mu.x, -sc.y * m1y - B.y * mu.y, -sc.z * m1z - B.z * mu.z,
                                           -sc.w * m1w - B.w * mu.w);
        s_coef[3 * c4n + c4] = (mask_mode == 2) ? __ldg(shift + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int cstep = (int)(stride % c4n);
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int c4 = (int)(i % c4n);
    for (; i < total4; i += stride) {
        const float4 zz = z[i];
        const float4 sc = s_coef[c4];
        float4 a = zz;
        if (mask_mode == 1) a = act[i];
        else if (mask_mode == 2) a = affine4(zz, sc, s_coef[3 * c4n + c4]);
        const float4 gg = masked_g(g[i], mask_mode, a);
        const float4 B = s_coef[c4n + c4], C = s_coef[2 * c4n + c4];
        float4 o;
        o.x = fmaf(sc.x, gg.x, fmaf(B.x, zz.x, C.x));
        o.y = fmaf(sc.y, gg.y, fmaf(B.y, zz.y, C.y));
        o.z = fmaf(sc.z, gg.z, fmaf(B.z, zz.z, C.z));
        o.w = fmaf(sc.w, gg.w, fmaf(B.w, zz.w, C.w));
        if (dz) dz[i] = o;
        if (dz_hi) {   // bf16 hi/lo planes for the bf16x3 data / weight gradient kernels
            const __nv_bfloat162 h0 = __floats2bfloat162_rn(o.x, o.y), h1 = __floats2bfloat162_rn(o.z, o.w);
            const __nv_bfloat162 l0 = __floats2bfloat162_rn(o.x - __low2float(h0), o.y - __high2float(h0));
            const __nv_bfloat162 l1 = __floats2bfloat162_rn(o.z - __low2float(h1), o.w - __high2float(h1));
            dz_hi[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
            dz_lo[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
        }
        if (a_hi) {   // bf16 hi/lo planes of this unit's OUTPUT ACTIVATION (mode 1: act; mode 2: relu(z*scale+shift)): the input
                      // operand of the weight gradient of the convolution that consumed it (saves that kernel's split pass)
            const float4 y = (mask_mode == 2) ? relu4(a) : a;
            const __nv_bfloat162 h0 = __floats2bfloat162_rn(y.x, y.y), h1 = __floats2bfloat162_rn(y.z, y.w);
            const __nv_bfloat162 l0 = __floats2bfloat162_rn(y.x - __low2float(h0), y.y - __high2float(h0));
            const __nv_bfloat162 l1 = __floats2bfloat162_rn(y.z - __low2float(h1), y.w - __high2float(h1));
            a_hi[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
            a_lo[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
        }
        if (gres) {
            float4 rr = gg;
            if (gres_accumulate) {
                const float4 old = gres[i];
                rr.x += old.x; rr.y += old.y; rr.z += old.z; rr.w += old.w;
            }
            gres[i] = rr;
        }
        c4 += cstep;
        if (c4 >= c4n) c4 -= c4n;
    }
}

// eval-mode / no-stat variant of the masked gradient (used for ReLU-only masks): out = g * (act > 0)
__global__ void relu_bwd_kernel(const float4* __restrict__ g, const float4* __restrict__ act, float4* __restrict__ out,
                                long long total4, int accumulate) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        float4 r = masked_g(g[i], 1, act[i]);
        if (accumulate) {
            const float4 old = out[i];
            r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w;
        }
        out[i] = r;
    }
}

// ---------------------------------------------------------------- pooling
// MaxPool2d(3, stride 2, pad 1) over relu(z*scale+shift); channels-last [nb, h, w, cs] -> [nb, ho, wo, cs]
__global__ void maxpool_fwd_kernel(const float4* __restrict__ z, const float4* __restrict__ scale,
                                   const float4* __restrict__ shift, float4* __restrict__ out, int nb, int h, int w,
                                   int c4n, int ho, int wo) {
    const long long total = (long long)nb * ho * wo * c4n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % c4n);
        long long r = i / c4n;
        const int ow = (int)(r % wo);
        r /= wo;
        const int oh = (int)(r % ho);
        const int n = (int)(r / ho);
        const float4 sc = __ldg(scale + c4), sh = __ldg(shift + c4);
        float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        for (int dh = 0; dh < 3; ++dh) {
            const int ih = oh * 2 - 1 + dh;
            if (ih < 0 || ih >= h) continue;
            for (int dw = 0; dw < 3; ++dw) {
                const int iw = ow * 2 - 1 + dw;
                if (iw < 0 || iw >= w) continue;
                const float4 a = relu4(affine4(z[((size_t)(n * h + ih) * w + iw) * c4n + c4], sc, sh));
                best.x = fmaxf(best.x, a.x); best.y = fmaxf(best.y, a.y);
                best.z = fmaxf(best.z, a.z); best.w = fmaxf(best.w, a.w);
            }
        }
        out[i] = best;
    }
}

// gradient wrt the pool INPUT activation a = relu(z*scale+shift): each input pixel gathers from the <= 4 windows
// that contain it and whose first maximum (scan order dh, dw — torch's tie rule) is this pixel.
__global__ void maxpool_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ z,
                                   const float* __restrict__ scale, const float* __restrict__ shift,
                                   float* __restrict__ da, int nb, int h, int w, int cs, int ho, int wo) {
    const long long total = (long long)nb * h * w * cs;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cs);
        long long r = i / cs;
        const int iw = (int)(r % w);
        r /= w;
        const int ih = (int)(r % h);
        const int n = (int)(r / h);
        const float sc = scale[c], sh = shift[c];
        float acc = 0.f;
        for (int oh = (ih + 1) / 2 - 1; oh <= (ih + 1) / 2; ++oh) {
            if (oh < 0 || oh >= ho || ih < oh * 2 - 1 || ih > oh * 2 + 1) continue;
            for (int ow = (iw + 1) / 2 - 1; ow <= (iw + 1) / 2; ++ow) {
                if (ow < 0 || ow >= wo || iw < ow * 2 - 1 || iw > ow * 2 + 1) continue;
                float best = -INFINITY;
                int bh = -1, bw = -1;
                for (int dh = 0; dh < 3; ++dh) {
                    const int yy = oh * 2 - 1 + dh;
                    if (yy < 0 || yy >= h) continue;
                    for (int dw = 0; dw < 3; ++dw) {
                        const int xx = ow * 2 - 1 + dw;
                        if (xx < 0 || xx >= w) continue;
                        const float a = fmaxf(fmaf(z[((size_t)(n * h + yy) * w + xx) * cs + c], sc, sh), 0.f);
                        if (a > best) {
                            best = a;
                            bh = yy;
                            bw = xx;
                        }
                    }
                }
                if (bh == ih && bw == iw) acc += dout[((size_t)(n * ho + oh) * wo + ow) * cs + c];
            }
        }
        da[i] = acc;
    }
}

// feat[n, c] = mean_p y[n, p, c]      (y channels-last [nb, P, cs]; feat dense [nb, c_real])
__global__ void avgpool_fwd_kernel(const float* __restrict__ y, float* __restrict__ feat, int nb, int P, int cs, int c_real) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = blockIdx.y;
    if (c >= c_real) return;
    float s = 0.f;
    for (int p = 0; p < P; ++p) s += y[((size_t)n * P + p) * cs + c];
    feat[(size_t)n * c_real + c] = s / (float)P;
}

__global__ void avgpool_bwd_kernel(const float* __restrict__ dfeat, float* __restrict__ dy, int nb, int P, int cs, int c_real) {
    const long long total = (long long)nb * P * cs;
    const float inv = 1.f / (float)P;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cs);
        const int n = (int)(i / ((long long)P * cs));
        dy[i] = c < c_real ? dfeat[(size_t)n * c_real + c] * inv : 0.f;
    }
}

// ---------------------------------------------------------------- layout
// x [nb, C, P] (NCDHW flattened) -> out [nb, P, cs] with zero pad channels; tiled through shared memory
__global__ void nchw_to_cl_kernel(const float* __restrict__ x, float* __restrict__ out, int C, long long P, int cs) {
    const int n = blockIdx.y;
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    float* o = out + ((size_t)n * P + p) * cs;
    for (int c = 0; c < cs; ++c) o[c] = c < C ? x[((size_t)n * C + c) * P + p] : 0.f;
}

// ---------------------------------------------------------------- optimizer
struct SgdEntry {
    float* p;
    const float* g;
    float* m;
    long long n;
};
// torch.optim.SGD(momentum, weight_decay), dampening 0, no nesterov: d = g + wd*p; buf = first ? d : mu*buf + d; p -= lr*buf
__global__ void sgd_kernel(const SgdEntry* __restrict__ table, int n_tensors, float lr, float mu, float wd, int first) {
    for (int t = blockIdx.y; t < n_tensors; t += gridDim.y) {
        const SgdEntry e = table[t];
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < e.n; i += (long long)gridDim.x * blockDim.x) {
            const float p = e.p[i];
            const float d = fmaf(wd, p, e.g[i]);
            const float b = first ? d : fmaf(mu, e.m[i], d);
            e.m[i] = b;
            e.p[i] = p - lr * b;
        }
    }
}

constexpr int SGD_CHUNK = 48;
struct SgdChunk {
    SgdEntry e[SGD_CHUNK];
};
__global__ void sgd_chunk_kernel(const SgdChunk c, int n_tensors, float lr, float mu, float wd, int first) {
    const int t = blockIdx.y;
    if (t >= n_tensors) return;
    const SgdEntry e = c.e[t];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < e.n; i += (long long)gridDim.x * blockDim.x) {
        const float p = e.p[i];
        const float d = fmaf(wd, p, e.g[i]);
        const float b = first ? d : fmaf(mu, e.m[i], d);
        e.m[i] = b;
        e.p[i] = p - lr * b;
    }
}

}  // namespace

#define LAUNCH_CHECK(what) SV_CUDA_CHECK(cudaGetLastError(), what)

namespace {
// Scratch of the chunked reduction, one per (device, stream): launches on ONE stream are ordered, so a reduction owns its
// stream's scratch until it completes; reductions on different streams (the audio tower and the weight-gradient side
// streams run next to the video tower) never share chunk slots or tickets.
struct RpScratch {
    int dev;
    cudaStream_t stream;
    double* scratch;
    unsigned* tickets;
};
std::mutex g_rp_mutex;
std::vector<RpScratch> g_rp_table;

int rp_scratch_for(int dev, cudaStream_t stream, double** scratch, unsigned** tickets) {
    std::lock_guard<std::mutex> lock(g_rp_mutex);
    for (const RpScratch& e : g_rp_table) {
        if (e.dev == dev && e.stream == stream) {
            *scratch = e.scratch;
            *tickets = e.tickets;
            return 0;
        }
    }
    RpScratch e{dev, stream, nullptr, nullptr};
    const size_t tbytes = sizeof(unsigned) * 2 * (RP_MAX_CS / 32);
    SV_CUDA_CHECK(cudaMalloc(&e.scratch, sizeof(double) * RP_MAX_CHUNKS * 2 * RP_MAX_CS), "bn_reduce_partials: scratch");
    SV_CUDA_CHECK(cudaMalloc(&e.tickets, tbytes), "bn_reduce_partials: tickets");
    // cudaMalloc/cudaMemset on the legacy default stream complete before this call returns to the (stream-ordered) caller
    SV_CUDA_CHECK(cudaMemset(e.tickets, 0, tbytes), "bn_reduce_partials: tickets");
    SV_CUDA_CHECK(cudaDeviceSynchronize(), "bn_reduce_partials: tickets");
    g_rp_table.push_back(e);
    *scratch = e.scratch;
    *tickets = e.tickets;
    return 0;
}
}  // namespace

extern "C" int selavi_bn_reduce_partials(const float* partial, int tiles, int ctot, int cs, double* sums, void* stream) {
    if (!partial || !sums || tiles <= 0 || cs <= 0 || cs > RP_MAX_CS) return selavi_fail(-1, "bn_reduce_partials: bad arguments");
    int nchunk = tiles / 256;
    if (nchunk < 1) nchunk = 1;
    if (nchunk > RP_MAX_CHUNKS) nchunk = RP_MAX_CHUNKS;
    double* scratch = nullptr;
    unsigned* tickets = nullptr;
    if (nchunk > 1) {
        int dev = 0;
        SV_CUDA_CHECK(cudaGetDevice(&dev), "bn_reduce_partials: cudaGetDevice");
        const int rc = rp_scratch_for(dev, (cudaStream_t)stream, &scratch, &tickets);
        if (rc) return rc;
    }
    dim3 grid((cs + 31) / 32, 2, nchunk), block(32, 32);
    bn_reduce_partials_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(partial, tiles, ctot, cs, sums, scratch, tickets);
    LAUNCH_CHECK("bn_reduce_partials");
    return 0;
}

extern "C" int selavi_bn_finalize(const double* sums, double count, const float* gamma, const float* beta,
                                  float* running_mean, float* running_var, float momentum, float eps, int c_real, int cs,
                                  float* scale, float* shift, float* mean, float* invstd, int update_running, void* stream) {
    if (!sums || !scale || !shift || !mean || !invstd || count <= 0 || (cs & 3)) return selavi_fail(-1, "bn_finalize: bad arguments");
    if (update_running && (!running_mean || !running_var)) return selavi_fail(-1, "bn_finalize: running stats missing");
    bn_finalize_kernel<<<(cs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, count, gamma, beta, running_mean, running_var,
                                                                          momentum, eps, c_real, cs, scale, shift, mean,
                                                                          invstd, update_running);
    LAUNCH_CHECK("bn_finalize");
    return 0;
}

extern "C" int selavi_bn_eval_affine(const float* gamma, const float* beta, const float* running_mean,
                                     const float* running_var, float eps, int c_real, int cs, float* scale, float* shift,
                                     void* stream) {
    if (!running_mean || !running_var || !scale || !shift || (cs & 3)) return selavi_fail(-1, "bn_eval_affine: bad arguments");
    bn_eval_affine_kernel<<<(cs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(gamma, beta, running_mean, running_var, eps,
                                                                             c_real, cs, scale, shift);
    LAUNCH_CHECK("bn_eval_affine");
    return 0;
}

extern "C" int selavi_bn_apply(const float* z, const float* scale, const float* shift, const float* res,
                               const float* rscale, const float* rshift, int relu, float* out, long long M, int cs,
                               void* stream) {
    if (!z || !scale || !shift || !out || (cs & 3) || M <= 0) return selavi_fail(-1, "bn_apply: bad arguments");
    const long long total4 = M * (cs / 4);
    bn_apply_kernel<<<ew_blocks(total4), EW_THREADS, 0, (cudaStream_t)stream>>>(
        (const float4*)z, (const float4*)scale, (const float4*)shift, (const float4*)res, (const float4*)rscale,
        (const float4*)rshift, relu, (float4*)out, total4, cs / 4);
    LAUNCH_CHECK("bn_apply");
    return 0;
}

extern "C" int selavi_bn_bwd_blocks(long long M) {
    long long b = (M + 63) / 64;
    if (b > 148 * 4) b = 148 * 4;
    return (int)(b < 1 ? 1 : b);
}

extern "C" int selavi_bn_bwd_reduce(const float* g, const float* z, const float* act, int mask_mode, const float* scale,
                                    const float* shift, const float* mean, const float* invstd, long long M, int cs,
                                    float* partial, double* sums, void* stream) {
    if (!g || !z || !mean || !invstd || !partial || !sums || (cs & 3) || M <= 0) return selavi_fail(-1, "bn_bwd_reduce: bad arguments");
    if (mask_mode == 1 && !act) return selavi_fail(-1, "bn_bwd_reduce: mask_mode 1 needs act");
    if (mask_mode == 2 && (!scale || !shift)) return selavi_fail(-1, "bn_bwd_reduce: mask_mode 2 needs scale/shift");
    const int c4n = cs / 4;
    const int nblk = selavi_bn_bwd_blocks(M);
    if (c4n <= 128 && c4n % 32 != 0) {
        const int R = 256 / c4n;   // >= 2 complete rows per block; 2 * c4n <= R * c4n threads for the final sum
        bn_bwd_reduce_rows_kernel<<<nblk, R * c4n, (size_t)2 * R * c4n * sizeof(float4), (cudaStream_t)stream>>>(
            (const float4*)g, (const float4*)z, (const float4*)act, mask_mode, (const float4*)scale, (const float4*)shift,
            (const float4*)mean, (const float4*)invstd, M, c4n, R, partial);
    } else {
        dim3 grid(nblk, (c4n + 31) / 32);
        bn_bwd_reduce_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)g, (const float4*)z, (const float4*)act,
                                                                     mask_mode, (const float4*)scale, (const float4*)shift,
                                                                     (const float4*)mean, (const float4*)invstd, M, c4n, partial);
    }
    LAUNCH_CHECK("bn_bwd_reduce");
    return selavi_bn_reduce_partials(partial, nblk, cs, cs, sums, stream);
}

extern "C" int selavi_bn_bwd_apply(const float* g, const float* z, const float* act, int mask_mode, const float* scale,
                                   const float* shift, const float* mean, const float* invstd, const double* sums,
                                   double count, long long M, int cs, float* dz, float* gres, int gres_accumulate,
                                   void* dz_hi, void* dz_lo, void* act_hi, void* act_lo, void* stream) {
    if (!g || !z || !scale || !mean || !invstd || !sums || (!dz && !dz_hi) || (cs & 3) || M <= 0 || count <= 0 ||
        ((dz_hi == nullptr) != (dz_lo == nullptr)) || ((act_hi == nullptr) != (act_lo == nullptr)))
        return selavi_fail(-1, "bn_bwd_apply: bad arguments");
    if (act_hi && mask_mode == 0) return selavi_fail(-1, "bn_bwd_apply: activation planes need mask_mode 1 or 2");
    if (mask_mode == 1 && !act) return selavi_fail(-1, "bn_bwd_apply: mask_mode 1 needs act");
    if (mask_mode == 2 && !shift) return selavi_fail(-1, "bn_bwd_apply: mask_mode 2 needs shift");
    const int c4n = cs / 4;
    const long long total4 = M * c4n;
    bn_bwd_apply_kernel<<<ew_blocks(total4), EW_THREADS, (size_t)4 * c4n * sizeof(float4), (cudaStream_t)stream>>>(
        (const float4*)g, (const float4*)z, (const float4*)act, mask_mode, (const float4*)scale, (const float4*)shift,
        (const float4*)mean, (const float4*)invstd, sums, count, total4, c4n, (float4*)dz, (float4*)gres, gres_accumulate,
        (uint2*)dz_hi, (uint2*)dz_lo, (uint2*)act_hi, (uint2*)act_lo);
    LAUNCH_CHECK("bn_bwd_apply");
    return 0;
}

extern "C" int selavi_relu_bwd(const float* g, const float* act, float* out, long long n, int accumulate, void* stream) {
    if (!g || !act || !out || (n & 3) || n <= 0) return selavi_fail(-1, "relu_bwd: bad arguments");
    relu_bwd_kernel<<<ew_blocks(n / 4), EW_THREADS, 0, (cudaStream_t)stream>>>((const float4*)g, (const float4*)act,
                                                                               (float4*)out, n / 4, accumulate);
    LAUNCH_CHECK("relu_bwd");
    return 0;
}

extern "C" int selavi_maxpool3x3s2_fwd(const float* z, const float* scale, const float* shift, float* out, int nb, int h,
                                       int w, int cs, void* stream) {
    if (!z || !scale || !shift || !out || (cs & 3)) return selavi_fail(-1, "maxpool_fwd: bad arguments");
    const int ho = (h + 2 - 3) / 2 + 1, wo = (w + 2 - 3) / 2 + 1;
    const long long total = (long long)nb * ho * wo * (cs / 4);
    maxpool_fwd_kernel<<<ew_blocks(total), EW_THREADS, 0, (cudaStream_t)stream>>>((const float4*)z, (const float4*)scale,
                                                                                  (const float4*)shift, (float4*)out, nb, h,
                                                                                  w, cs / 4, ho, wo);
    LAUNCH_CHECK("maxpool_fwd");
    return 0;
}

extern "C" int selavi_maxpool3x3s2_bwd(const float* dout, const float* z, const float* scale, const float* shift, float* da,
                                       int nb, int h, int w, int cs, void* stream) {
    if (!dout || !z || !scale || !shift || !da) return selavi_fail(-1, "maxpool_bwd: bad arguments");
    const int ho = (h + 2 - 3) / 2 + 1, wo = (w + 2 - 3) / 2 + 1;
    const long long total = (long long)nb * h * w * cs;
    maxpool_bwd_kernel<<<ew_blocks(total), EW_THREADS, 0, (cudaStream_t)stream>>>(dout, z, scale, shift, da, nb, h, w, cs, ho, wo);
    LAUNCH_CHECK("maxpool_bwd");
    return 0;
}

extern "C" int selavi_avgpool_fwd(const float* y, float* feat, int nb, int P, int cs, int c_real, void* stream) {
    if (!y || !feat || nb <= 0 || P <= 0) return selavi_fail(-1, "avgpool_fwd: bad arguments");
    dim3 grid((c_real + 127) / 128, nb);
    avgpool_fwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(y, feat, nb, P, cs, c_real);
    LAUNCH_CHECK("avgpool_fwd");
    return 0;
}

extern "C" int selavi_avgpool_bwd(const float* dfeat, float* dy, int nb, int P, int cs, int c_real, void* stream) {
    if (!dfeat || !dy || nb <= 0 || P <= 0) return selavi_fail(-1, "avgpool_bwd: bad arguments");
    avgpool_bwd_kernel<<<ew_blocks((long long)nb * P * cs), EW_THREADS, 0, (cudaStream_t)stream>>>(dfeat, dy, nb, P, cs, c_real);
    LAUNCH_CHECK("avgpool_bwd");
    return 0;
}

extern "C" int selavi_nchw_to_cl(const float* x, float* out, int nb, int C, long long P, int cs, void* stream) {
    if (!x || !out || nb <= 0 || C <= 0 || P <= 0 || cs < C) return selavi_fail(-1, "nchw_to_cl: bad arguments");
    dim3 grid((unsigned)((P + 255) / 256), nb);
    nchw_to_cl_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, out, C, P, cs);
    LAUNCH_CHECK("nchw_to_cl");
    return 0;
}

extern "C" int selavi_sgd_step(const void* table, int n_tensors, float lr, float momentum, float weight_decay, int first_step,
                               void* stream) {
    if (!table || n_tensors <= 0) return selavi_fail(-1, "sgd_step: bad arguments");
    dim3 grid(64, n_tensors < 512 ? n_tensors : 512);
    sgd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const SgdEntry*)table, n_tensors, lr, momentum, weight_decay, first_step);
    LAUNCH_CHECK("sgd_step");
    return 0;
}

extern "C" int selavi_sgd_step_host(void* const* params, void* const* grads, void* const* bufs, const long long* sizes,
                                    int n_tensors, float lr, float momentum, float weight_decay, int first_step, void* stream) {
    if (!params || !grads || !bufs || !sizes || n_tensors <= 0) return selavi_fail(-1, "sgd_step_host: bad arguments");
    for (int t0 = 0; t0 < n_tensors; t0 += SGD_CHUNK) {
        SgdChunk c;
        const int n = n_tensors - t0 < SGD_CHUNK ? n_tensors - t0 : SGD_CHUNK;
        for (int i = 0; i < n; ++i) {
            c.e[i].p = reinterpret_cast<float*>(params[t0 + i]);
            c.e[i].g = reinterpret_cast<const float*>(grads[t0 + i]);
            c.e[i].m = reinterpret_cast<float*>(bufs[t0 + i]);
            c.e[i].n = sizes[t0 + i];
        }
        sgd_chunk_kernel<<<dim3(64, n), 256, 0, (cudaStream_t)stream>>>(c, n, lr, momentum, weight_decay, first_step);
        LAUNCH_CHECK("sgd_step_host");
    }
    return 0;
}
