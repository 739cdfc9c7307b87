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
 *             buffers (selavi_symm_*) of 2*world*selavi_sk_kp(K) doubles / world uint32 per rank (receive slots
 *             and flags, one per source rank); the flag words must be zero on every rank when the call
 *             starts.  world == 1: pass NULL.
 * Replaces the NCCL all-gather + rank-0 solve of src/sk_utils.py:214-242,287-327.
 */
/* PS[n,k] = softmax_f64(logits_v[n,:])[k] * softmax_f64(logits_a[n,:])[k]  (src/sk_utils.py:206-211,309-315) */
int selavi_sk_softmax_product(const float* logits_v, const float* logits_a, long long n, int K, double* PS, void* stream);
/* out[n,k] = softmax_f64(logits[n,:])[k]  (torch.nn.functional.softmax(x, dim=1, dtype=torch.float64), src/sk_utils.py:272-275) */
int selavi_sk_softmax64(const float* logits, long long n, int K, double* out, void* stream);
/* C[i,j] = sum_n |P1[n,i] - P2[n,j]|, float64 [K,K]: every value of the cost function c(a, b) that the head-alignment
 * search `match_order` evaluates (src/sk_utils.py:430-447), computed once; P1, P2 [n,K] float64 row-major.  `workspace`
 * holds selavi_l1_cost_workspace_bytes(n, K) bytes of split-N partial tiles (summed in a fixed order: bit-reproducible). */
size_t selavi_l1_cost_workspace_bytes(long long n, int K);
int selavi_l1_cost_matrix(const double* P1, const double* P2, long long n, int K, double* C, void* workspace, void* stream);
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
 * Activations: channels-last fp32 [N,T,H,W,Cs], Cs = channels padded to a multiple of 8 (pad channels zero).
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
 * passes: 3 = tf32x3 split (fp32-class accuracy), 1 = single tf32 pass, 6 = fp16x3 (mode 0 only: operands split into
 *   fp16 hi/lo with exact power-of-two scaling, 22 significant bits like tf32x3 at twice the MMA rate and half the
 *   shared-memory bytes; wpack from selavi_conv_pack_weights mode 2, sized by selavi_conv_wpack_bytes_f16).
 */
int selavi_conv_tiles(int n_out, int* bnt, int* ntiles);
size_t selavi_conv_wpack_bytes(int n_out, int k_total);
size_t selavi_conv_wpack_bytes_f16(int n_out, int k_total);
int selavi_conv_pack_weights(const float* W, int mode, int co, int ci, int taps, int cs, void* wpack, void* stream);
int selavi_conv_gemm(const float* src, float* dst, const void* wpack, const int* geom, const float* pro_scale,
                     const float* pro_shift, int pro_relu, float* stats_partial, int accumulate, int passes,
                     void* stream);
/* weight gradient: geom is the FORWARD geometry (mode 0), dz = gradient wrt the conv output [M, cd];
 * dW in the torch layout [co][ci_real][taps]; workspace of selavi_wgrad_workspace_bytes(geom).
 * passes: 3 = bf16x3 split / 1 = plain bf16 (operands stay MN-major, tcgen05.mma.kind::f16);
 *         13 = tf32x3 / 11 = tf32 (K-major tiles, register-transposing loaders). */
size_t selavi_wgrad_workspace_bytes(const int* geom);
int selavi_conv_wgrad(const float* src, const float* dz, float* dW, const int* geom, int ci_real,
                      const float* pro_scale, const float* pro_shift, int pro_relu, void* workspace, int accumulate,
                      int passes, void* stream);
/* bf16x3 backward on pre-split gradients: z_hi/z_lo = bf16 hi/lo planes [M, cd] of the gradient wrt the conv output
 * (from selavi_bn_bwd_apply or selavi_split_bf16).  wgrad_bf16: same contract as selavi_conv_wgrad (passes 3 or 1).
 * dgrad_bf16: geom in mode 1; wpack from selavi_dgrad_pack_weights(W[co][ci][taps], same geom).  Strided convs are
 * processed per stride-parity class of the input pixels, each with its own tap subset (no zero-filled MMA work). */
int selavi_split_bf16(const float* x, const float* scale, const float* shift, int relu, void* hi, void* lo, long long M,
                      int cs, void* stream);
int selavi_conv_wgrad_bf16(const float* src, const void* z_hi, const void* z_lo, float* dW, const int* geom, int ci_real,
                           const float* pro_scale, const float* pro_shift, int pro_relu, void* workspace, int accumulate,
                           int passes, void* stream);
/* same with the conv input already split into bf16 hi/lo planes [pixels_in, cs] (act_hi/act_lo of selavi_bn_bwd_apply) */
int selavi_conv_wgrad_bf16_planes(const void* a_hi, const void* a_lo, const void* z_hi, const void* z_lo, float* dW,
                                  const int* geom, int ci_real, void* workspace, int accumulate, int passes, void* stream);
/* host-only: tiling of the bf16x3 weight gradient selavi_conv_wgrad(_bf16) will use for this geometry: 128-row tiles of the
 * flattened (tap, channel) rows, column tile width / count, row tiles per CTA (they share every dz stage), the largest number
 * of split-K slices a group of row tiles gets (slices are dealt in proportion to a group's row tiles; one wave of CTAs in
 * total), whether the operands are exchanged (dz as the tap-shifted row operand; stride-1 "same" convolutions only), and
 * the number of CTAs launched. */
int selavi_conv_wgrad_plan(const int* geom, int ci_real, int* mtiles, int* bnt, int* ntiles, int* tiles_per_cta, int* slices,
                           int* exchanged, int* ctas);
size_t selavi_dgrad_wpack_bytes(const int* geom);
int selavi_dgrad_pack_weights(const float* W, const int* geom, int co, void* wpack, void* stream);
int selavi_conv_dgrad_bf16(const void* z_hi, const void* z_lo, float* dx, const void* wpack, const int* geom,
                           int accumulate, int passes, void* stream);

/* Tap-reuse ("halo") forward convolution in fp16x3 for the stride-1 1x3x3 / 3x1x1 (and 2-D 3x3) convolutions of
 * tv:video/resnet.py:45-61 and tv:resnet.py:59-105: the input window of a 128-pixel tile is normalised, split and
 * staged in shared memory ONCE per 64-channel chunk and every tap reads it through a row-shifted UMMA descriptor
 * (conv_gemm gathers and converts it once per tap).  Same contract as selavi_conv_gemm in mode 0 (fused BN+ReLU
 * prologue, raw fp32 output, per-tile BN partial sums) except that stats_partial is [m_tiles][2][ntiles*bnt] with the
 * tile counts returned by selavi_conv_halo_plan.  plan returns 0 when the geometry is supported, 1 when it is not
 * (strided, other kernel shapes: use selavi_conv_gemm).  fp16 hi/lo operands carry 22 significant bits (as tf32x3);
 * activations are pre-scaled by 2^4, weights by 2^8 (exact), |activation| must stay below 4094.
 * flags: bit 0 = set the descriptor base-offset field for row-shifted windows (diagnostic; hardware wants 0). */
int selavi_conv_halo_plan(const int* geom, int* m_tiles, int* bnt, int* ntiles, size_t* wpack_bytes);
int selavi_conv_halo_pack_weights(const float* W, const int* geom, int k_real, void* wpack, void* stream);
int selavi_conv_halo_fwd(const float* src, float* dst, const void* wpack, const int* geom, const float* pro_scale,
                         const float* pro_shift, int pro_relu, float* stats_partial, int flags, void* stream);
/* Data gradient of the same convolutions with the same kernel (geom in mode 1; for stride 1 the gradient is the
 * convolution of dz with flipped taps and the transposed weight matrix): z_hi/z_lo = bf16 hi/lo planes of dz as for
 * selavi_conv_dgrad_bf16, copied into the tile by cp.async, bf16x3 MMAs, dx written or accumulated.  wpack from
 * selavi_conv_halo_pack_weights with the mode-1 geometry (k_real = co; forward: k_real = ci). */
int selavi_conv_halo_dgrad(const void* z_hi, const void* z_lo, float* dx, const void* wpack, const int* geom,
                           int accumulate, int flags, void* stream);
/* same, and the epilogue also emits the per-tile partial sums of the BatchNorm-backward pass that consumes dx:
 * stats_partial [m_tiles][2][ntiles*bnt] (tile counts of selavi_conv_halo_plan in mode 1) = (sum g*m, sum g*m*zhat), g = dx,
 * m = (bn_z*bn_scale+bn_shift > 0), zhat = (bn_z-bn_mean)*bn_invstd, for the unit whose raw output is bn_z [pixels_in, cis].
 * selavi_bn_reduce_partials turns them into the sums selavi_bn_bwd_reduce (mask_mode 2) would produce, without reading
 * dx and bn_z again from HBM.  The gradient is written (not accumulated). */
int selavi_conv_halo_dgrad_bnstats(const void* z_hi, const void* z_lo, float* dx, const void* wpack, const int* geom,
                                   const float* bn_z, const float* bn_scale, const float* bn_shift, const float* bn_mean,
                                   const float* bn_invstd, float* stats_partial, int flags, void* stream);

/* ------------------------------------------------------------------------------------------------
 * BatchNorm (train mode, nn.BatchNorm{1,2,3}d / SyncBatchNorm semantics), residual add, ReLU, pooling, layout,
 * SGD.  Activations channels-last fp32 [M, cs]; per-channel vectors have cs (padded) entries.
 *   reduce_partials: conv-epilogue tile partials [tiles][2][ctot] -> fp64 sums [2][cs] (sum, sum of squares)
 *   finalize:        sums over `count` elements (all ranks) -> scale = gamma*invstd, shift = beta - mean*scale,
 *                    saved mean / invstd, running-stat update (momentum, unbiased variance)
 *   apply:           out = act(z*scale+shift [+ res | + res*rscale+rshift])  (tv:video/resnet.py:107-119)
 *   bwd_reduce:      sums [2][cs] = (sum g, sum g*zhat), g masked by the following ReLU:
 *                    mask_mode 0 none, 1 act>0 (materialised block output), 2 z*scale+shift>0
 *   bwd_apply:       dz = scale*(g - sum_g/count - zhat*sum_gz/count) as fp32 (dz, nullable) and/or as bf16 hi/lo
 *                    planes (dz_hi/dz_lo, nullable) for the bf16x3 gradient kernels; optional gres (+)= masked g;
 *                    optional act_hi/act_lo (nullable, mask_mode 1 or 2): bf16 hi/lo planes of this unit's OUTPUT
 *                    activation (mode 1: act, mode 2: relu(z*scale+shift)) = the input operand of the weight gradient
 *                    of the convolution that consumed it (selavi_conv_wgrad_bf16_planes): saves that split pass
 */
int selavi_bn_reduce_partials(const float* partial, int tiles, int ctot, int cs, double* sums, void* stream);
int selavi_bn_finalize(const double* sums, double count, const float* gamma, const float* beta, float* running_mean,
                       float* running_var, float momentum, float eps, int c_real, int cs, float* scale, float* shift,
                       float* mean, float* invstd, int update_running, void* stream);
int selavi_bn_eval_affine(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                          float eps, int c_real, int cs, float* scale, float* shift, void* stream);
int selavi_bn_apply(const float* z, const float* scale, const float* shift, const float* res, const float* rscale,
                    const float* rshift, int relu, float* out, long long M, int cs, void* stream);
int selavi_bn_bwd_blocks(long long M);
int selavi_bn_bwd_reduce(const float* g, const float* z, const float* act, int mask_mode, const float* scale,
                         const float* shift, const float* mean, const float* invstd, long long M, int cs, float* partial,
                         double* sums, void* stream);
int selavi_bn_bwd_apply(const float* g, const float* z, const float* act, int mask_mode, const float* scale,
                        const float* shift, const float* mean, const float* invstd, const double* sums, double count,
                        long long M, int cs, float* dz, float* gres, int gres_accumulate, void* dz_hi, void* dz_lo,
                        void* act_hi, void* act_lo, void* stream);
int selavi_relu_bwd(const float* g, const float* act, float* out, long long n, int accumulate, void* stream);
/* MaxPool2d(3,2,1) over relu(z*scale+shift) (tv:resnet.py:268-271) and its gradient wrt that activation */
int selavi_maxpool3x3s2_fwd(const float* z, const float* scale, const float* shift, float* out, int nb, int h, int w,
                            int cs, void* stream);
int selavi_maxpool3x3s2_bwd(const float* dout, const float* z, const float* scale, const float* shift, float* da, int nb,
                            int h, int w, int cs, void* stream);
/* AdaptiveAvgPool(1): y [nb, P, cs] -> feat [nb, c_real] and back */
int selavi_avgpool_fwd(const float* y, float* feat, int nb, int P, int cs, int c_real, void* stream);
int selavi_avgpool_bwd(const float* dfeat, float* dy, int nb, int P, int cs, int c_real, void* stream);
/* [nb, C, P] (NCDHW) -> channels-last [nb, P, cs] with zero pad channels */
int selavi_nchw_to_cl(const float* x, float* out, int nb, int C, long long P, int cs, void* stream);
/* fused multi-tensor torch.optim.SGD step (main.py:132-137,302); table = device array of {p, g, buf, n} */
int selavi_sgd_step(const void* table, int n_tensors, float lr, float momentum, float weight_decay, int first_step,
                    void* stream);
/* same update, tensors given by HOST arrays of device pointers (passed to the kernels by value, 48 per launch): no
 * device-side table, hence no host-to-device copy / synchronisation when gradient buffers move between steps */
int selavi_sgd_step_host(void* const* params, void* const* grads, void* const* bufs, const long long* sizes, int n_tensors,
                         float lr, float momentum, float weight_decay, int first_step, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Projection heads (model.py:62-90 MLPv2 / nn.Linear heads, model.py:201-252) and cross-entropy
 * (utils.py:377-387), batched over the H heads of a modality.  *_tbl = device arrays of H pointers to the
 * per-head nn.Parameter storage (parameters stay ordinary tensors).
 *   bgemm: C[h](m,n) (+)= sum_k A[h](m,k)*Amask[h](m,k)*B[h](k,n) + bias[h](n), arbitrary element strides.
 *   heads_bn_*: BatchNorm1d over the batch rows of z [H,B,F] (train: batch stats / eval: running stats),
 *               heads_act: a = relu(z*scale+shift)*mask (mask = dropout mask/(1-p) or NULL).
 *   ce_loss: logits of head h at logit_tbl[h] (device pointer table) or, when logit_tbl is NULL, at
 *            logits_base + h*head_stride; loss_rows[h,b] = logsumexp(x) - x[label[b,h]], loss_mean = mean over (h,b) = get_loss();
 *            dlogits [H,B,K] = (softmax - onehot) * grad_scale (nullable).
 */
int selavi_bgemm(int H, int M, int N, int K, const float* A, const void* const* A_tbl, long long a_bs, long long a_sm,
                 long long a_sk, const float* Amask, long long am_bs, const float* B, const void* const* B_tbl,
                 long long b_bs, long long b_sk, long long b_sn, const float* bias, const void* const* bias_tbl,
                 long long bias_bs, float* C, long long c_bs, long long c_sm, long long c_sn, int accumulate, void* stream);
int selavi_heads_bn_stats(const float* z, int H, int B, int F, double* sums, void* stream);
int selavi_heads_bn_finalize(const double* sums, double count, const void* const* gamma_tbl, const void* const* beta_tbl,
                             const void* const* rmean_tbl, const void* const* rvar_tbl, float momentum, float eps, int H,
                             int F, float* scale, float* shift, float* mean, float* invstd, int update_running, void* stream);
int selavi_heads_bn_eval_affine(const void* const* gamma_tbl, const void* const* beta_tbl, const void* const* rmean_tbl,
                                const void* const* rvar_tbl, float eps, int H, int F, float* scale, float* shift,
                                void* stream);
int selavi_heads_act(const float* z, const float* scale, const float* shift, const float* mask, float* a, int H, int B,
                     int F, void* stream);
int selavi_heads_bn_bwd_reduce(const float* da, const float* mask, const float* z, const float* scale, const float* shift,
                               const float* mean, const float* invstd, int H, int B, int F, double* sums, void* stream);
int selavi_heads_bn_bwd_apply(const float* da, const float* mask, const float* z, const float* scale, const float* shift,
                              const float* mean, const float* invstd, const double* sums, double count, int H, int B, int F,
                              float* dz, void* stream);
int selavi_heads_sum_masked(const float* x, const float* mask, float* out, int H, long long BF, int accumulate, void* stream);
int selavi_heads_colsum(const float* x, float* out, int H, int M, int N, void* stream);
int selavi_ce_loss(const void* const* logit_tbl, const float* logits_base, long long head_stride, const long long* labels,
                   long long lab_stride_b, long long lab_stride_h, int H, int B, int K, float grad_scale, float* loss_rows,
                   float* loss_mean, float* dlogits, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Mel-spectrogram front end (datasets/audio_utils.py:47-63 -> python_speech_features.logfbank): pre-emphasis,
 * framing (rectangular window, zero padded), 1024-point FFT, |X|^2/NFFT, triangular mel filterbank given by its
 * floored bin edges bins[nfilt+2], eps floor, log, optional (x-1.93)/17.89.  signal [batch, samples] float64,
 * out [batch, 1, nfilt, numframes] float32.
 */
int selavi_mel_logfbank(const double* signal, int batch, long long samples, int frame_len, int frame_step, int numframes,
                        const double* bins, int nfilt, int nfft, double preemph, int z_normalize, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Clip augmentation of the input pipeline (datasets/video_transforms.py:462-504 clip_augmentation with the default flags:
 * x/255, -MEAN, /STD :474-477; bilinear short-side scale jitter :35-79; random / uniform crop :101-134,167-210; horizontal
 * flip :137-164; THWC -> CTHW :480,503) in one pass.  frames uint8 [n][T][H][W][3] (device), out float32
 * [n][3][T][crop][crop] (device); params is a HOST array [n][5] = (new_h, new_w, y_off, x_off, flip) per clip, drawn by
 * the caller with the reference's np.random call order (selavi_b200/video_transforms.py).
 */
int selavi_clip_augment(const unsigned char* frames, float* out, int n, int T, int H, int W, int crop, const int* params,
                        void* stream);

/* ------------------------------------------------------------------------------------------------
 * Symmetric peer-mapped buffers (CUDA IPC), the transport of the in-kernel NVSwitch exchange.
 * alloc: cudaMalloc + zero + export a 64-byte handle; open/close: map / unmap a peer's handle.
 */
int selavi_symm_alloc(size_t bytes, void** ptr_out, unsigned char* handle64);
int selavi_symm_open(const unsigned char* handle64, void** ptr_out);
int selavi_symm_close(void* ptr);
int selavi_symm_free(void* ptr);
int selavi_symm_memset(void* ptr, int value, size_t bytes, void* stream);
/* In-place sum all-reduce of data[n] (float64, n small) over peer memory, one CTA, push model — replaces the NCCL
 * collectives of SyncBatchNorm (torch:nn/modules/_functions.py:49-117,144-200).  peer_recv[r] / peer_flag[r]: rank r's
 * receive ring (doubles) and flag array (uint64, zero-initialised); every rank issues the same sequence of calls with
 * the same slot_off (ring offset in doubles; the call uses world*n doubles from there), flag_idx and seqval
 * (strictly increasing, > 0). */
int selavi_p2p_allreduce_f64(double* data, int n, int world, int rank, void* const* peer_recv, void* const* peer_flag,
                             long long slot_off, int flag_idx, long long seqval, void* stream);

#ifdef __cplusplus
}
#endif
#endif
