// Log mel filterbank front end (datasets/audio_utils.py:14-74 -> python_speech_features.logfbank), batched on the GPU.
// One CTA per (frame, clip): pre-emphasis + framing (rectangular window, zero padding) straight from the PCM
// slice, 1024-point radix-2 FFT in shared memory in float64 (the reference computes in float64 and casts the LOG
// to float32), power spectrum /NFFT, sparse triangular filterbank (bin edges from the host, ~2 non-zero spans per
// filter instead of the reference's dense 513x257 dot), eps floor, log, optional z-normalisation, transposed
// store out[b, 0, filt, frame] (coalesced over filters within a frame is not possible in that layout; the tensor
// is 100 KB per clip, the kernel is latency-bound).
#include <math.h>
#include <stdint.h>

#include "../../include/selavi_b200.h"
#include "common.cuh"

namespace {

constexpr int MEL_NFFT = 1024;
constexpr int MEL_LOG2 = 10;
constexpr int MEL_THREADS = 256;

__global__ void __launch_bounds__(MEL_THREADS) mel_kernel(const double* __restrict__ sig, long long L, int frame_len,
                                                          int frame_step, int numframes, const double* __restrict__ bins,
                                                          int nfilt, double preemph, int z_normalize,
                                                          float* __restrict__ out) {
    __shared__ double re[MEL_NFFT];
    __shared__ double im[MEL_NFFT];
    const int f = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const double* x = sig + (size_t)b * L;
    for (int n = tid; n < MEL_NFFT; n += MEL_THREADS) {
        const long long i = (long long)f * frame_step + n;
        double v = 0.0;
        if (n < frame_len && i < L) v = (i == 0) ? x[0] : x[i] - preemph * x[i - 1];
        const int r = __brev((unsigned)n) >> (32 - MEL_LOG2);
        re[r] = v;
        im[r] = 0.0;
    }
    __syncthreads();
    for (int s = 1; s <= MEL_LOG2; ++s) {
        const int half = 1 << (s - 1);
        for (int t = tid; t < MEL_NFFT / 2; t += MEL_THREADS) {
            const int j = t & (half - 1);
            const int i0 = ((t >> (s - 1)) << s) + j;
            const int i1 = i0 + half;
            double sn, cs;
            sincospi(-(double)j / (double)half, &sn, &cs);
            const double xr = re[i1] * cs - im[i1] * sn;
            const double xi = re[i1] * sn + im[i1] * cs;
            const double ur = re[i0], ui = im[i0];
            re[i0] = ur + xr;
            im[i0] = ui + xi;
            re[i1] = ur - xr;
            im[i1] = ui - xi;
        }
        __syncthreads();
    }
    // power spectrum of bins 0..512 into re[]
    for (int k = tid; k <= MEL_NFFT / 2; k += MEL_THREADS) {
        const double a = re[k], c = im[k];
        const double mag = sqrt(a * a + c * c);      // numpy.absolute, then numpy.square
        re[k] = (1.0 / MEL_NFFT) * (mag * mag);
    }
    __syncthreads();
    for (int j = tid; j < nfilt; j += MEL_THREADS) {
        const double b0 = bins[j], b1 = bins[j + 1], b2 = bins[j + 2];
        double acc = 0.0;
        for (int i = (int)b0; i < (int)b1; ++i) acc += re[i] * (((double)i - b0) / (b1 - b0));
        for (int i = (int)b1; i < (int)b2; ++i) acc += re[i] * ((b2 - (double)i) / (b2 - b1));
        if (acc == 0.0) acc = 2.220446049250313e-16;
        float v = (float)log(acc);
        if (z_normalize) v = (v - 1.93f) / 17.89f;
        out[((size_t)b * nfilt + j) * numframes + f] = v;
    }
}

}  // namespace

extern "C" int selavi_mel_logfbank(const double* signal, int batch, long long samples, int frame_len, int frame_step,
                                   int numframes, const double* bins, int nfilt, int nfft, double preemph, int z_normalize,
                                   float* out, void* stream) {
    if (!signal || !bins || !out || batch <= 0 || samples <= 0 || numframes <= 0 || nfilt <= 0)
        return selavi_fail(-1, "mel_logfbank: bad arguments");
    if (nfft != MEL_NFFT) return selavi_fail(-1, "mel_logfbank: only nfft=1024 (the reference's setting) is built");
    if (frame_len > nfft || frame_len <= 0 || frame_step <= 0) return selavi_fail(-1, "mel_logfbank: frame length must be in (0, nfft]");
    dim3 grid(numframes, batch);
    mel_kernel<<<grid, MEL_THREADS, 0, (cudaStream_t)stream>>>(signal, samples, frame_len, frame_step, numframes, bins, nfilt,
                                                              preemph, z_normalize, out);
    SV_CUDA_CHECK(cudaGetLastError(), "mel_logfbank: launch");
    return 0;
}
