/* selavi_b200 — C ABI of the B200-native SeLaVi training hot path.
 *
 * The reference (facebookresearch/selavi) is pure Python: its "FFI" for this path is the set of torch
 * library calls made by model.py / src/sk_utils.py / utils.py / datasets/audio_utils.py.  Each entry point
 * below replaces one such call site; the Python mirror modules in selavi_b200/ (and the drop-in shims in
 * dropin/) bind them with ctypes — see INTEGRATION.md for the reference-side binding.
 *
 * Conventions: plain pointers + sizes, device pointers unless the name says `host`; `stream` is a
 * cudaStream_t passed as void*; return 0 on success, a negative cudaError / argument code otherwise;
 * no hidden allocation (workspaces are caller-provided), nothing is retained past the call.
 */
#ifndef SELAVI_B200_H
#define SELAVI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

int selavi_version(void);
/* last error string of the calling thread's most recent failing call (static storage) */
const char* selavi_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * Sinkhorn-Knopp (reference: src/sk_utils.py:359-422 optimize_L_sk_gpu)
 * PS          [n_local, K] float64 row-major, consumed: raised to lamb/2 in place (sk_utils.py:391)
 * n_global    rows over all ranks (c = 1/N, beta0 = 1/N; sk_utils.py:390,395)
 * use_dist    0 = 'default' marginals (r = 1/K); 1 = match sorted `kdist` to argsort(PS.sum(0))
 *             (sk_utils.py:369,388); kdist [K] is permuted in place like args.dist[hc]
 * outputs     alpha[K], beta[n_local], labels[n_local] (int64 argmax, sk_utils.py:413),
 *             iters (int32), err (f64), cost_sum = nansum(log PS[n, L_n]) over local rows
 * stop rule   stop_on_converge=1: while err > tol and it < max_iters, err refreshed every
 *             `check_every` iterations (sk_utils.py:400-406). 0: exactly max_iters iterations.
 * do_prep     1: pow + marginals + initial sums (a fresh solve).  0: PS already powered, continue from the
 *             state left in `workspace` by the previous call (used by the iteration micro-benchmark).
 * world/rank  rows sharded over `world` GPUs; peer_sum[r] / peer_flag[r] are P2P-mapped symmetric
 *             buffers (selavi_symm_*) of 2*selavi_sk_kp(K) doubles / 1 uint32 per rank; the flag word
 *             must be zero on every rank when the call starts.  world == 1: pass NULL.
 * Replaces the NCCL all-gather + rank-0 solve of src/sk_utils.py:214-242,287-327.
 */
size_t selavi_sk_workspace_bytes(int K);
int selavi_sk_kp(int K);
int selavi_sk_solve(double* PS, long long n_local, long long n_global, int K, double lamb, int use_dist,
                    double* kdist, double* alpha_out, double* beta_out, long long* labels_out, void* workspace,
                    int max_iters, int check_every, double tol, int stop_on_converge, int do_prep, int do_final,
                    int* iters_out, double* err_out, double* cost_sum_out, int world, int rank,
                    void* const* peer_sum, void* const* peer_flag, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Convolutions as implicit GEMM on tcgen05 tensor cores (reference: the cuDNN conv3d/conv2d fwd/dgrad/wgrad
 * calls behind torchvision Conv2Plus1D / BasicBlock / stems / downsamples, tv:video/resnet.py:45-61,184-195,
 * 276-281 and tv:resnet.py:59-105, built by model.py:93-121).
 *
 * Activations: channels-last fp32 [N,T,H,W,Cs], Cs = channels padded to a multiple of 4 (pad channels zero).
 * geom[20] = {mode, nb, ts,hs,ws,cs, td,hd,wd,cd, kt,kh,kw, st,sh,sw, pt,ph,pw, n_out}
 *   mode 0 (forward):  src = conv input, dst = conv output; dst pixel (t,h,w) reads src (t*st-pt+kt, ...).
 *   mode 1 (dgrad):    src = gradient wrt conv output, dst = gradient wrt conv input; st.. and pt.. are the FORWARD
 *                      stride/padding; dst pixel (t,h,w) reads src ((t+pt-kt)/st, ...) where divisible.
 *   n_out = real channel count of dst (<= cd).
 * wpack: weights pre-tiled / pre-swizzled / tf32 hi-lo split by selavi_conv_pack_weights from the torch
 *   layout W[co][ci][kt][kh][kw] (mode as above; cs = channel stride of the gathered tensor).
 * pro_scale/pro_shift [cs] (nullable): fused prologue x -> x*scale+shift (+ReLU if pro_relu) on every
 *   gathered element = the train-mode BatchNorm(+ReLU) of the previous layer; zero padding stays zero.
 * stats_partial (nullable, forward): [ceil(M/128)][2][ntiles*bnt] per-tile column sum / sum of squares of dst.
 * passes: 3 = tf32x3 split (fp32-class accuracy), 1 = single tf32 pass.
 */
int selavi_conv_tiles(int n_out, int* bnt, int* ntiles);
size_t selavi_conv_wpack_bytes(int n_out, int k_total);
int selavi_conv_pack_weights(const float* W, int mode, int co, int ci, int taps, int cs, void* wpack, void* stream);
int selavi_conv_gemm(const float* src, float* dst, const void* wpack, const int* geom, const float* pro_scale,
                     const float* pro_shift, int pro_relu, float* stats_partial, int accumulate, int passes,
                     void* stream);
/* weight gradient: geom is the FORWARD geometry (mode 0), dz = gradient wrt the conv output [M, cd];
 * dW in the torch layout [co][ci_real][taps]; workspace of selavi_wgrad_workspace_bytes(co, taps, cs, M). */
size_t selavi_wgrad_workspace_bytes(int co, int taps, int cs, long long M);
int selavi_conv_wgrad(const float* src, const float* dz, float* dW, const int* geom, int ci_real,
                      const float* pro_scale, const float* pro_shift, int pro_relu, void* workspace, int accumulate,
                      int passes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Symmetric peer-mapped buffers (CUDA IPC), the transport of the in-kernel NVSwitch exchange.
 * alloc: cudaMalloc + zero + export a 64-byte handle; open/close: map / unmap a peer's handle.
 */
int selavi_symm_alloc(size_t bytes, void** ptr_out, unsigned char* handle64);
int selavi_symm_open(const unsigned char* handle64, void** ptr_out);
int selavi_symm_close(void* ptr);
int selavi_symm_free(void* ptr);
int selavi_symm_memset(void* ptr, int value, size_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif
