"""Mel front end: oracle facts derived from the reference call site on CPU, kernel-vs-oracle on GPU."""
import numpy as np
import pytest

from oracle import mel_oracle as mo


def _wav(sr, sec, seed=0):
    return (np.random.default_rng(seed).standard_normal(sr * sec) * 3000).astype(np.int16)


def test_oracle_shapes_and_zero_filters():
    eps_log = np.log(np.finfo(float).eps)
    s24 = mo.get_spec(_wav(24000, 3), 0.5, sample_rate=24000, aud_spec_type=2)
    assert s24.shape == (1, 257, 99) and s24.dtype == np.float32
    assert int((np.abs(s24 - eps_log) < 1e-3).all(axis=2).sum()) == 38          # identically-zero filters -> log(eps)
    s48 = mo.get_spec(_wav(48000, 3), 1.0, sample_rate=48000, aud_spec_type=2)
    assert s48.shape == (1, 257, 99)
    assert int((np.abs(s48 - eps_log) < 1e-3).all(axis=2).sum()) == 56
    assert mo.get_spec(_wav(24000, 3), 0.0, num_sec=2, sample_rate=24000, aud_spec_type=1).shape == (1, 40, 199)
    assert mo.frame_counts(24000, 24000) == (480, 240, 99) and mo.frame_counts(48000, 48000) == (960, 480, 99)
    fb = mo.get_filterbanks(257, 1024, 24000)
    assert fb.shape == (257, 513) and 700 < np.count_nonzero(fb) < 900


def test_oracle_end_of_track_and_znorm():
    w = _wav(24000, 2)
    a = mo.get_spec(w, 1.7, sample_rate=24000, aud_spec_type=2)                    # clip would run past the end (:34-36)
    b = mo.get_spec(w, 1.0, sample_rate=24000, aud_spec_type=2)
    assert np.array_equal(a, b)
    z = mo.get_spec(w, 0.0, sample_rate=24000, aud_spec_type=2, z_normalize=True)
    r = mo.get_spec(w, 0.0, sample_rate=24000, aud_spec_type=2)
    np.testing.assert_allclose(z, (r - np.float32(1.93)) / np.float32(17.89), rtol=0, atol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("sr,sec,spec_type,znorm", [(24000, 1, 2, False), (48000, 1, 2, True), (24000, 2, 1, False),
                                                   (16000, 1, 2, False)])
def test_kernel_matches_oracle(cuda_device, sr, sec, spec_type, znorm):
    import torch
    from selavi_b200 import audio_utils
    w = _wav(sr, 4, seed=sr)
    for fr in (0.0, 1.3):
        ref = mo.get_spec(w, fr, num_sec=sec, sample_rate=sr, aud_spec_type=spec_type, z_normalize=znorm)
        out = audio_utils.get_spec(w, fr, num_sec=sec, sample_rate=sr, aud_spec_type=spec_type, z_normalize=znorm)
        assert isinstance(out, torch.Tensor) and out.dtype == torch.float32 and tuple(out.shape) == ref.shape
        np.testing.assert_allclose(out.numpy(), ref, rtol=2e-6, atol=2e-6)
        assert float((out.numpy() != ref).mean()) < 0.02                             # float32-rounding-boundary cases only
    # batched entry point, volume jitter consumes the same numpy draws
    np.random.seed(3)
    ref = mo.get_spec(w, 0.5, sample_rate=sr, aud_spec_type=spec_type, use_volume_jittering=True, use_temporal_jittering=True)
    np.random.seed(3)
    out = audio_utils.get_spec(w, 0.5, sample_rate=sr, aud_spec_type=spec_type, use_volume_jittering=True,
                               use_temporal_jittering=True)
    np.testing.assert_allclose(out.numpy(), ref, rtol=2e-6, atol=2e-6)
    sig = torch.from_numpy(np.stack([w[:sr], w[sr:2 * sr]])).to(cuda_device)
    batch = audio_utils.logfbank_batch(sig, sr, nfilt=257)
    assert tuple(batch.shape) == (2, 1, 257, mo.frame_counts(sr, sr)[2])
    np.testing.assert_allclose(batch[1, 0].cpu().numpy(), mo.logfbank(w[sr:2 * sr], sr).T.astype(np.float32), rtol=2e-6, atol=2e-6)
