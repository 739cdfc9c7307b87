// K x K L1 cost matrix of the head-alignment search `match_order` (src/sk_utils.py:424-467).
//
// The reference evaluates c(a, b) = sum_n |a_n - b_n| for four column pairs per hill-climb step (~10^5 steps x 15 tiny
// kernels with an .item() sync each).  Every value it ever needs is an entry of
//     C[i, j] = sum_n |P1[n, i] - P2[n, j]|        (P1, P2: float64 softmax outputs of the video / audio head, [N, K])
// so C is computed ONCE on the device (N*K^2 abs-diff-adds in float64, CUDA-core bound) and the hill-climb runs on it on
// the host with the identical np.random stream (selavi_b200/sk_utils.py:match_order).
//
// Kernel: a CTA owns a 64 x 64 tile of C and a slice of the rows; 256 threads hold 4 x 4 accumulators each; rows are
// staged through shared memory 16 at a time.  Row slices (split-N) fill the 148 SMs when K is small; their partial tiles
// are summed in slice order by a second kernel (deterministic, no atomics).
#include <stdint.h>

#include "../../include/selavi_b200.h"
#include "common.cuh"

namespace {

constexpr int L1_TILE = 64;
constexpr int L1_ROWS = 16;     // rows staged per step
constexpr int L1_THREADS = 256;

__global__ void __launch_bounds__(L1_THREADS) l1_cost_kernel(const double* __restrict__ P1, const double* __restrict__ P2,
                                                             long long n, int K, int kt, long long rows_per_slice,
                                                             double* __restrict__ partial /*[slices][kt*64][kt*64]*/) {
    __shared__ double sa[L1_ROWS][L1_TILE];
    __shared__ double sb[L1_ROWS][L1_TILE];
    const int ti = blockIdx.x / kt, tj = blockIdx.x % kt;
    const int slice = blockIdx.y;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;       // thread (ty, tx) owns rows i = ty + 16*u, columns j = tx + 16*v of the tile
    const long long r0 = (long long)slice * rows_per_slice;
    long long r1 = r0 + rows_per_slice;
    if (r1 > n) r1 = n;
    double acc[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = 0.0;
    const int i0 = ti * L1_TILE, j0 = tj * L1_TILE;
    for (long long r = r0; r < r1; r += L1_ROWS) {
        // stage 16 rows x 64 columns of both operands (out-of-range rows / columns load 0 for BOTH => |0 - 0| = 0)
        for (int e = tid; e < L1_ROWS * L1_TILE; e += L1_THREADS) {
            const int rr = e / L1_TILE, c = e % L1_TILE;
            const long long row = r + rr;
            const bool rok = row < r1;
            sa[rr][c] = (rok && i0 + c < K) ? P1[(size_t)row * K + i0 + c] : 0.0;
            sb[rr][c] = (rok && j0 + c < K) ? P2[(size_t)row * K + j0 + c] : 0.0;
        }
        __syncthreads();
#pragma unroll 4
        for (int rr = 0; rr < L1_ROWS; ++rr) {
            double a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) a[u] = sa[rr][ty + 16 * u];
#pragma unroll
            for (int v = 0; v < 4; ++v) b[v] = sb[rr][tx + 16 * v];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) acc[u][v] += fabs(a[u] - b[v]);
        }
        __syncthreads();
    }
    const int Kp = kt * L1_TILE;
    double* out = partial + (size_t)slice * Kp * Kp;
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) out[(size_t)(i0 + ty + 16 * u) * Kp + j0 + tx + 16 * v] = acc[u][v];
}

__global__ void l1_cost_reduce_kernel(const double* __restrict__ partial, int slices, int Kp, int K, double* __restrict__ C) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= K * K) return;
    const int i = idx / K, j = idx % K;
    double s = 0.0;
    for (int k = 0; k < slices; ++k) s += partial[((size_t)k * Kp + i) * Kp + j];
    C[idx] = s;
}

// out[n, k] = softmax_k(float64(logits[n, :]))  — torch.nn.functional.softmax(x, dim=1, dtype=torch.float64)
// (src/sk_utils.py:272-275): one warp per row
__global__ void softmax64_kernel(const float* __restrict__ x, long long n, int K, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const float* xr = x + (size_t)row * K;
    double mx = -INFINITY;
    for (int k = lane; k < K; k += 32) mx = fmax(mx, (double)xr[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    double s = 0.0;
    for (int k = lane; k < K; k += 32) s += exp((double)xr[k] - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    double* o_ = out + (size_t)row * K;
    for (int k = lane; k < K; k += 32) o_[k] = exp((double)xr[k] - mx) / s;
}

void l1_plan(long long n, int K, int* kt, int* slices, long long* rows_per_slice) {
    *kt = (K + L1_TILE - 1) / L1_TILE;
    const int tiles = (*kt) * (*kt);
    long long s = (148 * 3 + tiles - 1) / tiles;       // about three waves of CTAs
    const long long max_s = (n + 4 * L1_ROWS - 1) / (4 * L1_ROWS);   // at least 64 rows per slice
    if (s > max_s) s = max_s;
    if (s < 1) s = 1;
    *rows_per_slice = (n + s - 1) / s;
    *slices = (int)((n + *rows_per_slice - 1) / *rows_per_slice);
}

}  // namespace

extern "C" size_t selavi_l1_cost_workspace_bytes(long long n, int K) {
    if (n <= 0 || K <= 0) return 0;
    int kt, slices;
    long long rps;
    l1_plan(n, K, &kt, &slices, &rps);
    return (size_t)slices * kt * L1_TILE * kt * L1_TILE * sizeof(double);
}

extern "C" int selavi_l1_cost_matrix(const double* P1, const double* P2, long long n, int K, double* C, void* workspace,
                                     void* stream) {
    if (!P1 || !P2 || !C || !workspace || n <= 0 || K <= 0 || K > 4096) return selavi_fail(-1, "l1_cost_matrix: bad arguments");
    int kt, slices;
    long long rps;
    l1_plan(n, K, &kt, &slices, &rps);
    dim3 grid(kt * kt, slices);
    l1_cost_kernel<<<grid, L1_THREADS, 0, (cudaStream_t)stream>>>(P1, P2, n, K, kt, rps, reinterpret_cast<double*>(workspace));
    SV_CUDA_CHECK(cudaGetLastError(), "l1_cost_matrix: launch");
    l1_cost_reduce_kernel<<<(K * K + 255) / 256, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const double*>(workspace), slices,
                                                                                kt * L1_TILE, K, C);
    SV_CUDA_CHECK(cudaGetLastError(), "l1_cost_matrix: reduce launch");
    return 0;
}

extern "C" int selavi_sk_softmax64(const float* logits, long long n, int K, double* out, void* stream) {
    if (!logits || !out || n <= 0 || K <= 0) return selavi_fail(-1, "sk_softmax64: bad arguments");
    const int warps = 8;
    const long long blocks = (n + warps - 1) / warps;
    softmax64_kernel<<<(unsigned)blocks, warps * 32, 0, (cudaStream_t)stream>>>(logits, n, K, out);
    SV_CUDA_CHECK(cudaGetLastError(), "sk_softmax64: launch");
    return 0;
}
