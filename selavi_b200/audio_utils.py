"""B200-native mirror of the reference `datasets/audio_utils.py:get_spec` (datasets/audio_utils.py:14-74).

`get_spec` keeps the reference signature and returns a CPU FloatTensor [1, nfilt, T] by default (it is a drop-in
for code running in the main process; CUDA cannot be used inside forked DataLoader workers).  `logfbank_batch`
is the batched GPU entry point for raw PCM already on the device: the whole log mel filterbank front end
(python_speech_features.logfbank) is one CUDA kernel launch (csrc/mel.cu).
"""
import decimal
import math

import numpy as np
import torch

from . import _lib


def _round_half_up(number):
    return int(decimal.Decimal(number).quantize(decimal.Decimal('1'), rounding=decimal.ROUND_HALF_UP))


def _filterbank_bins(nfilt, nfft, samplerate):
    hz2mel = lambda hz: 2595 * np.log10(1 + hz / 700.)          # noqa: E731
    mel2hz = lambda mel: 700 * (10 ** (mel / 2595.0) - 1)        # noqa: E731
    melpoints = np.linspace(hz2mel(0), hz2mel(samplerate / 2), nfilt + 2)
    return np.floor((nfft + 1) * mel2hz(melpoints) / samplerate)


_bins_cache = {}


def logfbank_batch(signal, samplerate, winlen=0.02, winstep=0.01, nfilt=257, nfft=1024, preemph=0.97, z_normalize=False):
    """signal: CUDA tensor [B, L] (any real dtype; converted to float64 like numpy does) -> float32 [B, 1, nfilt, T]."""
    if not signal.is_cuda:
        raise ValueError("logfbank_batch needs a CUDA tensor (no CPU fallback)")
    if signal.dim() == 1:
        signal = signal[None]
    sig = signal.to(torch.float64).contiguous()
    B, L = sig.shape
    frame_len = _round_half_up(winlen * samplerate)
    frame_step = _round_half_up(winstep * samplerate)
    numframes = 1 if L <= frame_len else 1 + int(math.ceil((1.0 * L - frame_len) / frame_step))
    key = (nfilt, nfft, samplerate, sig.device)
    bins = _bins_cache.get(key)
    if bins is None:
        bins = torch.from_numpy(_filterbank_bins(nfilt, nfft, samplerate)).to(sig.device)
        _bins_cache[key] = bins
    out = torch.empty((B, 1, nfilt, numframes), dtype=torch.float32, device=sig.device)
    with torch.cuda.device(sig.device):
        _lib.check(_lib.lib().selavi_mel_logfbank(_lib.ptr(sig), B, L, frame_len, frame_step, numframes, _lib.ptr(bins), nfilt,
                                                  nfft, float(preemph), 1 if z_normalize else 0, _lib.ptr(out),
                                                  _lib.stream_ptr()), "selavi_mel_logfbank")
    return out


def get_spec(wav, fr_sec, num_sec=1, sample_rate=48000, aug_audio=[], aud_spec_type=1, use_volume_jittering=False,
             use_temporal_jittering=False, z_normalize=False, device="cuda", return_cpu=True):
    """Same contract (and the same numpy RNG draws) as datasets/audio_utils.py:14-74."""
    if use_temporal_jittering:
        fr_sec = fr_sec + np.random.uniform(-0.5, 0.5)
    fr_aud = int(np.round(fr_sec * sample_rate))
    to_aud = int(np.round(fr_sec * sample_rate) + sample_rate * num_sec)
    if fr_aud + (to_aud - fr_aud) > len(wav):
        fr_aud = len(wav) - sample_rate * num_sec
        to_aud = len(wav)
    wav = wav[fr_aud: to_aud]
    if use_volume_jittering:
        wav = wav * np.random.uniform(0.9, 1.1)
    sig = torch.from_numpy(np.ascontiguousarray(wav)).to(device)
    nfilt = 40 if aud_spec_type == 1 else 257
    spec = logfbank_batch(sig, sample_rate, winlen=0.02, winstep=0.01, nfilt=nfilt, nfft=1024, z_normalize=z_normalize)[0]
    return spec.cpu() if return_cpu else spec
