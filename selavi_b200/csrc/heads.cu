// Multi-head MLP projection (model.py:62-90 MLPv2, applied per head in AVModel.forward model.py:233-252) and the
// cross-entropy on pseudo-labels (utils.py:377-387), batched over all heads of a modality in single launches.
// The work is tiny (16.8 MFLOP per sample for 20 heads) and weight-read bound, so these are fp32 CUDA-core
// kernels: exact fp32 FMA arithmetic, ~30 launches per step instead of the reference's ~360.
#include <stdint.h>

#include "../../include/selavi_b200.h"
#include "common.cuh"

namespace {

struct BgemmParams {
    int H, M, N, K;
    const float* A; const void* const* A_tbl; long long a_bs, a_sm, a_sk;
    const float* Amask; long long am_bs;          // optional elementwise mask on A (same m/k strides)
    const float* B; const void* const* B_tbl; long long b_bs, b_sk, b_sn;
    const float* bias; const void* const* bias_tbl; long long bias_bs;
    float* C; long long c_bs, c_sm, c_sn;
    int accumulate;
};

// C[h](m,n) (+)= sum_k A[h](m,k) * Amask[h](m,k) * B[h](k,n) + bias[h](n);   16x16 tiles, generic strides
__global__ void bgemm_kernel(const BgemmParams p) {
    __shared__ float As[16][17];
    __shared__ float Bs[16][17];
    const int h = blockIdx.z;
    const float* A = p.A_tbl ? reinterpret_cast<const float*>(p.A_tbl[h]) : p.A + (size_t)h * p.a_bs;
    const float* B = p.B_tbl ? reinterpret_cast<const float*>(p.B_tbl[h]) : p.B + (size_t)h * p.b_bs;
    const float* Am = p.Amask ? p.Amask + (size_t)h * p.am_bs : nullptr;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int m = blockIdx.y * 16 + ty, n = blockIdx.x * 16 + tx;
    float acc = 0.f;
    for (int k0 = 0; k0 < p.K; k0 += 16) {
        // A tile: rows m (ty), cols k0+tx
        {
            const int k = k0 + tx;
            float v = 0.f;
            if (m < p.M && k < p.K) {
                v = A[(size_t)m * p.a_sm + (size_t)k * p.a_sk];
                if (Am) v *= Am[(size_t)m * p.a_sm + (size_t)k * p.a_sk];
            }
            As[ty][tx] = v;
        }
        {
            const int k = k0 + ty;
            float v = 0.f;
            if (k < p.K && n < p.N) v = B[(size_t)k * p.b_sk + (size_t)n * p.b_sn];
            Bs[ty][tx] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) acc = fmaf(As[ty][kk], Bs[kk][tx], acc);
        __syncthreads();
    }
    if (m < p.M && n < p.N) {
        if (p.bias || p.bias_tbl) {
            const float* b = p.bias_tbl ? reinterpret_cast<const float*>(p.bias_tbl[h]) : p.bias + (size_t)h * p.bias_bs;
            acc += b[n];
        }
        float* c = p.C + (size_t)h * p.c_bs + (size_t)m * p.c_sm + (size_t)n * p.c_sn;
        *c = p.accumulate ? (*c + acc) : acc;
    }
}

// z [H,B,F] -> sums [H][2][F] fp64 (sum, sum of squares over the B rows)
__global__ void heads_bn_stats_kernel(const float* __restrict__ z, int B, int F, double* __restrict__ sums) {
    const int h = blockIdx.y, f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    double s = 0.0, q = 0.0;
    for (int b = 0; b < B; ++b) {
        const double v = z[((size_t)h * B + b) * F + f];
        s += v;
        q += v * v;
    }
    sums[((size_t)h * 2) * F + f] = s;
    sums[((size_t)h * 2 + 1) * F + f] = q;
}

__global__ void heads_bn_finalize_kernel(const double* __restrict__ sums, double count, const void* const* gamma_tbl,
                                         const void* const* beta_tbl, const void* const* rmean_tbl,
                                         const void* const* rvar_tbl, float momentum, float eps, int F,
                                         float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                         float* __restrict__ invstd_out, int update_running) {
    const int h = blockIdx.y, f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const double mean = sums[((size_t)h * 2) * F + f] / count;
    double var = sums[((size_t)h * 2 + 1) * F + f] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float g = reinterpret_cast<const float*>(gamma_tbl[h])[f], b = reinterpret_cast<const float*>(beta_tbl[h])[f];
    const float sc = g * invstd;
    const size_t o = (size_t)h * F + f;
    scale[o] = sc;
    shift[o] = b - (float)mean * sc;
    mean_out[o] = (float)mean;
    invstd_out[o] = invstd;
    if (update_running) {
        float* rm = reinterpret_cast<float*>(const_cast<void*>(rmean_tbl[h]));
        float* rv = reinterpret_cast<float*>(const_cast<void*>(rvar_tbl[h]));
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        rm[f] = (1.f - momentum) * rm[f] + momentum * (float)mean;
        rv[f] = (1.f - momentum) * rv[f] + momentum * (float)unbiased;
    }
}

__global__ void heads_bn_eval_affine_kernel(const void* const* gamma_tbl, const void* const* beta_tbl,
                                            const void* const* rmean_tbl, const void* const* rvar_tbl, float eps, int F,
                                            float* __restrict__ scale, float* __restrict__ shift) {
    const int h = blockIdx.y, f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const float invstd = 1.f / sqrtf(reinterpret_cast<const float*>(rvar_tbl[h])[f] + eps);
    const float sc = reinterpret_cast<const float*>(gamma_tbl[h])[f] * invstd;
    scale[(size_t)h * F + f] = sc;
    shift[(size_t)h * F + f] = reinterpret_cast<const float*>(beta_tbl[h])[f] - reinterpret_cast<const float*>(rmean_tbl[h])[f] * sc;
}

// a = relu(z*scale+shift) * mask
__global__ void heads_act_kernel(const float* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ shift,
                                 const float* __restrict__ mask, float* __restrict__ a, int B, int F, long long total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(i % F);
        const int h = (int)(i / ((long long)B * F));
        float v = fmaxf(fmaf(z[i], scale[(size_t)h * F + f], shift[(size_t)h * F + f]), 0.f);
        if (mask) v *= mask[i];
        a[i] = v;
    }
}

// dy = da * mask * (z*scale+shift > 0);  sums [H][2][F] = (sum dy, sum dy*zhat)
__global__ void heads_bn_bwd_reduce_kernel(const float* __restrict__ da, const float* __restrict__ mask,
                                           const float* __restrict__ z, const float* __restrict__ scale,
                                           const float* __restrict__ shift, const float* __restrict__ mean,
                                           const float* __restrict__ invstd, int B, int F, double* __restrict__ sums) {
    const int h = blockIdx.y, f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const size_t o = (size_t)h * F + f;
    const float sc = scale[o], sh = shift[o], mu = mean[o], is = invstd[o];
    double s1 = 0.0, s2 = 0.0;
    for (int b = 0; b < B; ++b) {
        const size_t i = ((size_t)h * B + b) * F + f;
        const float zz = z[i];
        float g = da[i];
        if (mask) g *= mask[i];
        if (!(fmaf(zz, sc, sh) > 0.f)) g = 0.f;
        s1 += g;
        s2 += (double)g * (double)((zz - mu) * is);
    }
    sums[((size_t)h * 2) * F + f] = s1;
    sums[((size_t)h * 2 + 1) * F + f] = s2;
}

__global__ void heads_bn_bwd_apply_kernel(const float* __restrict__ da, const float* __restrict__ mask,
                                          const float* __restrict__ z, const float* __restrict__ scale,
                                          const float* __restrict__ shift, const float* __restrict__ mean,
                                          const float* __restrict__ invstd, const double* __restrict__ sums, double count,
                                          int B, int F, long long total, float* __restrict__ dz) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(i % F);
        const int h = (int)(i / ((long long)B * F));
        const size_t o = (size_t)h * F + f;
        const float zz = z[i], sc = scale[o];
        float g = da[i];
        if (mask) g *= mask[i];
        if (!(fmaf(zz, sc, shift[o]) > 0.f)) g = 0.f;
        const float m1 = (float)(sums[((size_t)h * 2) * F + f] / count);
        const float m2 = (float)(sums[((size_t)h * 2 + 1) * F + f] / count);
        dz[i] = sc * (g - m1 - (zz - mean[o]) * invstd[o] * m2);
    }
}

// out[b,f] (+)= sum_h x[h,b,f] * mask[h,b,f]
__global__ void heads_sum_masked_kernel(const float* __restrict__ x, const float* __restrict__ mask, float* __restrict__ out,
                                        int H, long long BF, int accumulate) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < BF; i += (long long)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int h = 0; h < H; ++h) {
            float v = x[(size_t)h * BF + i];
            if (mask) v *= mask[(size_t)h * BF + i];
            s += v;
        }
        out[i] = accumulate ? out[i] + s : s;
    }
}

// column sums: out[h][n] = sum_m x[h][m][n]   (bias gradients)
__global__ void heads_colsum_kernel(const float* __restrict__ x, float* __restrict__ out, int M, int N) {
    const int h = blockIdx.y, n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float s = 0.f;
    for (int m = 0; m < M; ++m) s += x[((size_t)h * M + m) * N + n];
    out[(size_t)h * N + n] = s;
}

// One block per (h, b): loss[h,b] = logsumexp(logits) - logits[label]; dlogits = (softmax - onehot) * gscale
__global__ void ce_kernel(const void* const* logit_tbl, const float* __restrict__ logits_base, long long head_stride,
                          const long long* __restrict__ labels, long long lab_sb, long long lab_sh, int B, int K, float gscale,
                          float* __restrict__ loss, float* __restrict__ dlogits) {
    __shared__ float red[32];
    const int h = blockIdx.y, b = blockIdx.x;
    const float* x = (logit_tbl ? reinterpret_cast<const float*>(logit_tbl[h]) : logits_base + (size_t)h * head_stride) + (size_t)b * K;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    float mx = -INFINITY;
    for (int k = tid; k < K; k += blockDim.x) mx = fmaxf(mx, x[k]);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < nw; ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float s = 0.f;
    for (int k = tid; k < K; k += blockDim.x) s += expf(x[k] - mx);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    s = 0.f;
    for (int w = 0; w < nw; ++w) s += red[w];
    const long long lab = labels[(size_t)b * lab_sb + (size_t)h * lab_sh];
    const float lse = logf(s) + mx;
    // A label outside [0, K) is a corrupt / uninitialised pseudo-label (torch.nn.functional.cross_entropy raises a device
    // assert for it): poison the loss row and its gradient row with NaN so that the step fails loudly instead of
    // training on softmax-only gradients.
    const bool lab_ok = lab >= 0 && lab < K;
    if (tid == 0) loss[(size_t)h * B + b] = lab_ok ? lse - x[lab] : __int_as_float(0x7fc00000);
    if (dlogits) {
        float* d = dlogits + ((size_t)h * B + b) * K;
        const float inv = 1.f / s;
        for (int k = tid; k < K; k += blockDim.x)
            d[k] = lab_ok ? (expf(x[k] - mx) * inv - (k == lab ? 1.f : 0.f)) * gscale : __int_as_float(0x7fc00000);
    }
}

// out[0] = mean(x[0..n))  (fixed order, one block)
__global__ void mean_kernel(const float* __restrict__ x, int n, float* __restrict__ out) {
    __shared__ double red[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += x[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = (float)(red[0] / n);
}

inline int hblocks(long long n) {
    long long b = (n + 255) / 256;
    return (int)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b));
}

}  // namespace

#define LAUNCH_CHECK(what) SV_CUDA_CHECK(cudaGetLastError(), what)

extern "C" int selavi_bgemm(int H, int M, int N, int K, const float* A, const void* const* A_tbl, long long a_bs,
                            long long a_sm, long long a_sk, const float* Amask, long long am_bs, const float* B,
                            const void* const* B_tbl, long long b_bs, long long b_sk, long long b_sn, const float* bias,
                            const void* const* bias_tbl, long long bias_bs, float* C, long long c_bs, long long c_sm,
                            long long c_sn, int accumulate, void* stream) {
    if (H <= 0 || M <= 0 || N <= 0 || K <= 0 || (!A && !A_tbl) || (!B && !B_tbl) || !C) return selavi_fail(-1, "bgemm: bad arguments");
    BgemmParams p{H, M, N, K, A, A_tbl, a_bs, a_sm, a_sk, Amask, am_bs, B, B_tbl, b_bs, b_sk, b_sn, bias, bias_tbl, bias_bs,
                  C, c_bs, c_sm, c_sn, accumulate};
    dim3 grid((N + 15) / 16, (M + 15) / 16, H), block(16, 16);
    bgemm_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(p);
    LAUNCH_CHECK("bgemm");
    return 0;
}

extern "C" int selavi_heads_bn_stats(const float* z, int H, int B, int F, double* sums, void* stream) {
    if (!z || !sums) return selavi_fail(-1, "heads_bn_stats: bad arguments");
    heads_bn_stats_kernel<<<dim3((F + 127) / 128, H), 128, 0, (cudaStream_t)stream>>>(z, B, F, sums);
    LAUNCH_CHECK("heads_bn_stats");
    return 0;
}

extern "C" int selavi_heads_bn_finalize(const double* sums, double count, const void* const* gamma_tbl,
                                        const void* const* beta_tbl, const void* const* rmean_tbl,
                                        const void* const* rvar_tbl, float momentum, float eps, int H, int F, float* scale,
                                        float* shift, float* mean, float* invstd, int update_running, void* stream) {
    if (!sums || !gamma_tbl || !beta_tbl || !scale || !shift || !mean || !invstd || count <= 0) return selavi_fail(-1, "heads_bn_finalize: bad arguments");
    heads_bn_finalize_kernel<<<dim3((F + 127) / 128, H), 128, 0, (cudaStream_t)stream>>>(
        sums, count, gamma_tbl, beta_tbl, rmean_tbl, rvar_tbl, momentum, eps, F, scale, shift, mean, invstd, update_running);
    LAUNCH_CHECK("heads_bn_finalize");
    return 0;
}

extern "C" int selavi_heads_bn_eval_affine(const void* const* gamma_tbl, const void* const* beta_tbl,
                                           const void* const* rmean_tbl, const void* const* rvar_tbl, float eps, int H,
                                           int F, float* scale, float* shift, void* stream) {
    if (!gamma_tbl || !beta_tbl || !rmean_tbl || !rvar_tbl || !scale || !shift) return selavi_fail(-1, "heads_bn_eval_affine: bad arguments");
    heads_bn_eval_affine_kernel<<<dim3((F + 127) / 128, H), 128, 0, (cudaStream_t)stream>>>(gamma_tbl, beta_tbl, rmean_tbl,
                                                                                           rvar_tbl, eps, F, scale, shift);
    LAUNCH_CHECK("heads_bn_eval_affine");
    return 0;
}

extern "C" int selavi_heads_act(const float* z, const float* scale, const float* shift, const float* mask, float* a, int H,
                                int B, int F, void* stream) {
    if (!z || !scale || !shift || !a) return selavi_fail(-1, "heads_act: bad arguments");
    const long long total = (long long)H * B * F;
    heads_act_kernel<<<hblocks(total), 256, 0, (cudaStream_t)stream>>>(z, scale, shift, mask, a, B, F, total);
    LAUNCH_CHECK("heads_act");
    return 0;
}

extern "C" int selavi_heads_bn_bwd_reduce(const float* da, const float* mask, const float* z, const float* scale,
                                          const float* shift, const float* mean, const float* invstd, int H, int B, int F,
                                          double* sums, void* stream) {
    if (!da || !z || !scale || !shift || !mean || !invstd || !sums) return selavi_fail(-1, "heads_bn_bwd_reduce: bad arguments");
    heads_bn_bwd_reduce_kernel<<<dim3((F + 127) / 128, H), 128, 0, (cudaStream_t)stream>>>(da, mask, z, scale, shift, mean,
                                                                                          invstd, B, F, sums);
    LAUNCH_CHECK("heads_bn_bwd_reduce");
    return 0;
}

extern "C" int selavi_heads_bn_bwd_apply(const float* da, const float* mask, const float* z, const float* scale,
                                         const float* shift, const float* mean, const float* invstd, const double* sums,
                                         double count, int H, int B, int F, float* dz, void* stream) {
    if (!da || !z || !scale || !shift || !mean || !invstd || !sums || !dz || count <= 0) return selavi_fail(-1, "heads_bn_bwd_apply: bad arguments");
    const long long total = (long long)H * B * F;
    heads_bn_bwd_apply_kernel<<<hblocks(total), 256, 0, (cudaStream_t)stream>>>(da, mask, z, scale, shift, mean, invstd, sums,
                                                                               count, B, F, total, dz);
    LAUNCH_CHECK("heads_bn_bwd_apply");
    return 0;
}

extern "C" int selavi_heads_sum_masked(const float* x, const float* mask, float* out, int H, long long BF, int accumulate,
                                       void* stream) {
    if (!x || !out || H <= 0 || BF <= 0) return selavi_fail(-1, "heads_sum_masked: bad arguments");
    heads_sum_masked_kernel<<<hblocks(BF), 256, 0, (cudaStream_t)stream>>>(x, mask, out, H, BF, accumulate);
    LAUNCH_CHECK("heads_sum_masked");
    return 0;
}

extern "C" int selavi_heads_colsum(const float* x, float* out, int H, int M, int N, void* stream) {
    if (!x || !out) return selavi_fail(-1, "heads_colsum: bad arguments");
    heads_colsum_kernel<<<dim3((N + 127) / 128, H), 128, 0, (cudaStream_t)stream>>>(x, out, M, N);
    LAUNCH_CHECK("heads_colsum");
    return 0;
}

extern "C" int selavi_ce_loss(const void* const* logit_tbl, const float* logits_base, long long head_stride,
                              const long long* labels, long long lab_stride_b, long long lab_stride_h, int H, int B, int K,
                              float grad_scale, float* loss_rows, float* loss_mean, float* dlogits, void* stream) {
    if ((!logit_tbl && !logits_base) || !labels || !loss_rows || !loss_mean || H <= 0 || B <= 0 || K <= 0) return selavi_fail(-1, "ce_loss: bad arguments");
    ce_kernel<<<dim3(B, H), 128, 0, (cudaStream_t)stream>>>(logit_tbl, logits_base, head_stride, labels, lab_stride_b, lab_stride_h, B, K, grad_scale,
                                                           loss_rows, dlogits);
    LAUNCH_CHECK("ce_loss");
    mean_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(loss_rows, H * B, loss_mean);
    LAUNCH_CHECK("ce_loss mean");
    return 0;
}
