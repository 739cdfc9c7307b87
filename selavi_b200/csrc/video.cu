// GPU clip augmentation of the input pipeline: uint8 THWC frames -> float32 CTHW clip, one pass.
// Replaces datasets/video_transforms.py:462-504 (clip_augmentation with the default flags): x/255, -MEAN, /STD
// (:474-477), bilinear short-side scale jitter (F.interpolate, align_corners=False, :35-79), crop (:101-134 /
// :167-210), horizontal flip (:137-164), THWC -> TCHW -> CTHW (:480,503).  The random draws (size, offsets, flip) are
// made on the host by the mirror (selavi_b200/video_transforms.py) with the reference's np.random call order and
// arrive here as five ints per clip.  HBM bound: 3 B read per source pixel touched, 12 B written per output pixel.
#include <stdint.h>
#include <stdlib.h>

#include "../../include/selavi_b200.h"
#include "common.cuh"

namespace {

constexpr int VA_MAX_CLIPS = 64;

struct ClipAugParams {
    const unsigned char* frames;   // [n][T][H][W][3]
    float* out;                    // [n][3][T][crop][crop]
    int n, T, H, W, crop;
    int fma;                       // bit 0: fused multiply-add in the source index, bit 1: in the blends
    int prm[VA_MAX_CLIPS][5];      // new_h, new_w, y_off, x_off, flip
};

__device__ __forceinline__ float va_norm(unsigned char v) {
    // ((v / 255) - 0.45) / 0.225 in float32, the reference's order of operations
    return __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.0f), 0.45f), 0.225f);
}

// torch's area_pixel_compute_source_index (align_corners = False): scale * (dst + 0.5) - 0.5, clamped at 0
__device__ __forceinline__ void va_src(int dst, float scale, int in_size, int fma, int& i0, int& step, float& l1) {
    float s = (fma & 1) ? fmaf(scale, __fadd_rn((float)dst, 0.5f), -0.5f) : __fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f);
    if (s < 0.f) s = 0.f;
    i0 = (int)s;
    if (i0 > in_size - 1) i0 = in_size - 1;
    step = i0 < in_size - 1 ? 1 : 0;
    l1 = __fsub_rn(s, (float)i0);
}

__global__ void clip_augment_kernel(const ClipAugParams p) {
    const long long per_clip = (long long)p.T * p.crop * p.crop;
    const long long total = per_clip * p.n;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(idx % p.crop);
        long long r = idx / p.crop;
        const int y = (int)(r % p.crop);
        r /= p.crop;
        const int t = (int)(r % p.T);
        const int n = (int)(r / p.T);
        const int nh = p.prm[n][0], nw = p.prm[n][1], yo = p.prm[n][2], xo = p.prm[n][3], flip = p.prm[n][4];
        const int xs = flip ? p.crop - 1 - x : x;   // the flip acts on the cropped image
        const int oy = yo + y, ox = xo + xs;        // position in the scaled image
        const unsigned char* f = p.frames + ((size_t)n * p.T + t) * (size_t)p.H * p.W * 3;
        float v[3];
        if (nh == p.H && nw == p.W) {               // no interpolation (short side already equals the drawn size)
            const unsigned char* q = f + ((size_t)oy * p.W + ox) * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c] = va_norm(q[c]);
        } else {
            const float sh = __fdiv_rn((float)p.H, (float)nh), sw = __fdiv_rn((float)p.W, (float)nw);
            int h0, hs, w0, ws;
            float lh1, lw1;
            va_src(oy, sh, p.H, p.fma, h0, hs, lh1);
            va_src(ox, sw, p.W, p.fma, w0, ws, lw1);
            const float lh0 = __fsub_rn(1.f, lh1), lw0 = __fsub_rn(1.f, lw1);
            const unsigned char* q00 = f + ((size_t)h0 * p.W + w0) * 3;
            const unsigned char* q01 = q00 + ws * 3;
            const unsigned char* q10 = q00 + (size_t)hs * p.W * 3;
            const unsigned char* q11 = q10 + ws * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float top, bot;
                if (p.fma & 2) {
                    top = fmaf(lw1, va_norm(q01[c]), __fmul_rn(lw0, va_norm(q00[c])));
                    bot = fmaf(lw1, va_norm(q11[c]), __fmul_rn(lw0, va_norm(q10[c])));
                    v[c] = fmaf(lh1, bot, __fmul_rn(lh0, top));
                } else {
                    top = __fadd_rn(__fmul_rn(lw0, va_norm(q00[c])), __fmul_rn(lw1, va_norm(q01[c])));
                    bot = __fadd_rn(__fmul_rn(lw0, va_norm(q10[c])), __fmul_rn(lw1, va_norm(q11[c])));
                    v[c] = __fadd_rn(__fmul_rn(lh0, top), __fmul_rn(lh1, bot));
                }
            }
        }
        float* o = p.out + (size_t)n * 3 * per_clip + ((size_t)t * p.crop + y) * p.crop + x;
#pragma unroll
        for (int c = 0; c < 3; ++c) o[(size_t)c * per_clip] = v[c];
    }
}

}  // namespace

// frames: uint8 [n][T][H][W][3] (device); params: HOST array [n][5] = (new_h, new_w, y_off, x_off, flip) per clip;
// out: float32 [n][3][T][crop][crop] (device).  Launches ceil(n / 64) kernels on `stream`, no synchronisation.
extern "C" int selavi_clip_augment(const unsigned char* frames, float* out, int n, int T, int H, int W, int crop, const int* params,
                                   void* stream) {
    if (!frames || !out || !params || n <= 0 || T <= 0 || H <= 0 || W <= 0 || crop <= 0)
        return selavi_fail(-1, "clip_augment: bad arguments");
    for (int i = 0; i < n; ++i) {
        const int* q = params + 5 * i;
        if (q[0] < crop || q[1] < crop || q[2] < 0 || q[3] < 0 || q[2] + crop > q[0] || q[3] + crop > q[1])
            return selavi_fail(-1, "clip_augment: crop window outside the scaled frame");
    }
    const size_t in_clip = (size_t)T * H * W * 3, out_clip = (size_t)3 * T * crop * crop;
    for (int base = 0; base < n; base += VA_MAX_CLIPS) {
        ClipAugParams p;
        p.n = n - base < VA_MAX_CLIPS ? n - base : VA_MAX_CLIPS;
        p.frames = frames + (size_t)base * in_clip;
        p.out = out + (size_t)base * out_clip;
        p.T = T; p.H = H; p.W = W; p.crop = crop;
        // torch's CPU kernel evaluates the source index with a fused multiply-add (measured: 4.8e-7 max deviation with
        // the fused index vs 3.2e-5 without, tools/va_probe.py); SELAVI_CLIPAUG_FMA overrides for that probe only
        p.fma = getenv("SELAVI_CLIPAUG_FMA") ? atoi(getenv("SELAVI_CLIPAUG_FMA")) : 1;
        for (int i = 0; i < p.n; ++i)
            for (int k = 0; k < 5; ++k) p.prm[i][k] = params[5 * (base + i) + k];
        const long long total = (long long)p.n * T * crop * crop;
        long long blocks = (total + 255) / 256;
        if (blocks > 148 * 16) blocks = 148 * 16;
        clip_augment_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(p);
        SV_CUDA_CHECK(cudaGetLastError(), "clip_augment: launch");
    }
    return 0;
}
