"""CPU oracle (numpy float64) of the mel-spectrogram front end.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference computes it with third-party `python_speech_features==0.6`
(environment.yml:145; call site datasets/audio_utils.py:10,47-63) whose source is NOT in /root/reference and not
installable offline, and the reference has no test or fixture for it.  This file restates the package's published
algorithm (logfbank -> fbank -> preemphasis / framesig / powspec / get_filterbanks) from its documentation; it
is the de-facto spec for the CUDA kernel.  What IS checked against the reference: the call-site contract of
datasets/audio_utils.py:14-74 (slicing, jitter order, dtype, transpose, shapes 257x99 / 40x99, z-normalisation).
"""
import decimal
import math

import numpy as np


def round_half_up(number):
    return int(decimal.Decimal(number).quantize(decimal.Decimal('1'), rounding=decimal.ROUND_HALF_UP))


def hz2mel(hz):
    return 2595 * np.log10(1 + hz / 700.)


def mel2hz(mel):
    return 700 * (10 ** (mel / 2595.0) - 1)


def filterbank_bins(nfilt, nfft, samplerate, lowfreq=0, highfreq=None):
    """bin edges (float64, already floored) of the nfilt triangular filters: nfilt+2 values."""
    highfreq = highfreq or samplerate / 2
    melpoints = np.linspace(hz2mel(lowfreq), hz2mel(highfreq), nfilt + 2)
    return np.floor((nfft + 1) * mel2hz(melpoints) / samplerate)


def get_filterbanks(nfilt, nfft, samplerate, lowfreq=0, highfreq=None):
    b = filterbank_bins(nfilt, nfft, samplerate, lowfreq, highfreq)
    fbank = np.zeros([nfilt, nfft // 2 + 1])
    for j in range(nfilt):
        for i in range(int(b[j]), int(b[j + 1])):
            fbank[j, i] = (i - b[j]) / (b[j + 1] - b[j])
        for i in range(int(b[j + 1]), int(b[j + 2])):
            fbank[j, i] = (b[j + 2] - i) / (b[j + 2] - b[j + 1])
    return fbank


def frame_counts(slen, samplerate, winlen=0.02, winstep=0.01):
    frame_len = round_half_up(winlen * samplerate)
    frame_step = round_half_up(winstep * samplerate)
    numframes = 1 if slen <= frame_len else 1 + int(math.ceil((1.0 * slen - frame_len) / frame_step))
    return frame_len, frame_step, numframes


def logfbank(signal, samplerate, winlen=0.02, winstep=0.01, nfilt=257, nfft=1024, preemph=0.97):
    """[numframes, nfilt] float64 log mel filterbank energies (rectangular window)."""
    signal = np.asarray(signal)
    sig = np.append(signal[0], signal[1:] - preemph * signal[:-1]).astype(np.float64)
    frame_len, frame_step, numframes = frame_counts(len(sig), samplerate, winlen, winstep)
    padlen = (numframes - 1) * frame_step + frame_len
    pad = np.concatenate((sig, np.zeros(padlen - len(sig))))
    idx = np.arange(frame_len)[None, :] + frame_step * np.arange(numframes)[:, None]
    frames = pad[idx]
    pspec = 1.0 / nfft * np.square(np.absolute(np.fft.rfft(frames, nfft)))
    feat = np.dot(pspec, get_filterbanks(nfilt, nfft, samplerate).T)
    feat = np.where(feat == 0, np.finfo(float).eps, feat)
    return np.log(feat)


def get_spec(wav, fr_sec, num_sec=1, sample_rate=48000, aug_audio=(), aud_spec_type=1, use_volume_jittering=False,
             use_temporal_jittering=False, z_normalize=False):
    """datasets/audio_utils.py:14-74 restated (returns numpy float32 [1, nfilt, T])."""
    if use_temporal_jittering:
        fr_sec = fr_sec + np.random.uniform(-0.5, 0.5)                       # :26-27
    fr_aud = int(np.round(fr_sec * sample_rate))                             # :30
    to_aud = int(np.round(fr_sec * sample_rate) + sample_rate * num_sec)     # :31
    if fr_aud + (to_aud - fr_aud) > len(wav):                                # :34-36
        fr_aud = len(wav) - sample_rate * num_sec
        to_aud = len(wav)
    wav = wav[fr_aud: to_aud]                                                # :39
    if use_volume_jittering:
        wav = wav * np.random.uniform(0.9, 1.1)                              # :42-43
    nfilt = 40 if aud_spec_type == 1 else 257                                # :46-63
    spec = logfbank(wav, sample_rate, winlen=0.02, winstep=0.01, nfilt=nfilt, nfft=1024)
    spec = np.expand_dims(spec.astype('float32').T, axis=0)                  # :66-68
    if z_normalize:
        spec = (spec - np.float32(1.93)) / np.float32(17.89)                 # :71-72 (float32 tensor arithmetic)
    return spec
