"""SpecAugment masks: host-side band sampling (CPU tier) and the in-place GPU application."""
import numpy as np
import pytest
import torch

from avsl_b200.audio import SPEC_AUGMENT_POLICIES, spec_augment_bands
from oracle import specaug as OS


def test_band_sampling_ranges_and_determinism():
    frames = [3000, 517, 40, 0]
    a = spec_augment_bands(frames, 80, "ls-double", np.random.default_rng(5))
    b = spec_augment_bands(frames, 80, "ls-double", np.random.default_rng(5))
    np.testing.assert_array_equal(a, b)
    assert a.shape == (4, 4, 4) and a.dtype == np.int32
    for i, tau in enumerate(frames):
        for f0, f1, t0, t1 in a[i, :2]:                   # frequency masks: full time extent
            assert 0 <= f0 <= f1 <= 80 and f1 - f0 <= 27 and t0 == 0 and t1 >= 3000
        for f0, f1, t0, t1 in a[i, 2:]:                   # time masks inside the unpadded part
            assert (f0, f1) == (0, 80) and 0 <= t0 <= t1 <= max(tau, 0) and t1 - t0 <= min(100, tau)
    assert spec_augment_bands([100], 128, "ls-basic", np.random.default_rng(0)).shape == (1, 2, 4)
    assert set(SPEC_AUGMENT_POLICIES) == {"ls-basic", "ls-double"}
    with pytest.raises(NotImplementedError):
        spec_augment_bands([100], 80, "ld")


def test_oracle_apply_bands_clips():
    mel = np.ones((1, 4, 6), np.float32)
    out = OS.apply_bands(mel, np.array([[[1, 3, 0, 2 ** 31 - 1], [0, 4, 4, 9], [2, 2, 0, 6]]], np.int32), -1.0)
    exp = np.ones((4, 6), np.float32)
    exp[1:3, :] = -1.0
    exp[:, 4:6] = -1.0
    np.testing.assert_array_equal(out[0], exp)


@pytest.mark.gpu
@pytest.mark.parametrize("n_mels", [80, 128])
def test_gpu_masks_bit_exact(n_mels):
    import torch
    import avsl_b200 as A
    rng = np.random.default_rng(3)
    mel = rng.standard_normal((5, n_mels, 3000)).astype(np.float32)
    frames = [3000, 1200, 77, 1, 2999]
    bands = A.spec_augment_bands(frames, n_mels, "ls-double", np.random.default_rng(11))
    got = A.spec_augment(torch.from_numpy(mel).cuda(), bands=bands).cpu().numpy()
    np.testing.assert_array_equal(got, OS.apply_bands(mel, bands))
    assert (got == 0).sum() > 0 and (got != mel).sum() == (got == 0).sum() - (mel == 0).sum()
    # 2-D input, custom fill, drawn inside the call
    one = torch.from_numpy(mel[0]).cuda()
    A.spec_augment(one, audio_frames=[3000], policy="ls-basic", rng=np.random.default_rng(2), fill=-0.5)
    ref = OS.apply_bands(mel[:1], A.spec_augment_bands([3000], n_mels, "ls-basic", np.random.default_rng(2)), -0.5)
    np.testing.assert_array_equal(one.cpu().numpy(), ref[0])


# ------------------------------------------------------------------ time warping + the dual-encoder mirrors
def test_warp_points_sampling_contract():
    import avsl_b200 as A
    rng = np.random.default_rng(5)
    pts = A.spec_augment_warp_points([3000, 500, 160, 100, 161], W=80, rng=rng)
    assert pts.shape == (5, 3) and pts.dtype == np.int32
    for tau, c, w in pts[[0, 1, 4]]:
        assert 80 <= c < tau - 80 and abs(int(w) - int(c)) <= 80 and 0 < w < tau
    assert pts[2, 1] == 0 and pts[3, 1] == 0            # tau <= 2 W: no warp


def test_oracle_time_warp_properties():
    from oracle import specaug as OS
    rng = np.random.default_rng(1)
    mel = rng.standard_normal((2, 6, 300)).astype(np.float32)
    ident = OS.time_warp(mel, np.array([[300, 120, 120], [300, 0, 0]], np.int32))
    np.testing.assert_array_equal(ident, mel)           # center == warped: s(t) = t exactly
    ramp = np.tile(np.arange(300, dtype=np.float32), (1, 4, 1))
    out = OS.time_warp(ramp, np.array([[200, 100, 150]], np.int32))
    np.testing.assert_array_equal(out[0, :, 200:], ramp[0, :, 200:])          # padding untouched
    assert out[0, 0, 0] == 0 and abs(out[0, 0, 150] - 100) < 1e-4 and abs(out[0, 0, 199] - 198.0) < 1.1
    assert (np.diff(out[0, 0, :200]) > 0).all()         # monotone: a warp, not a shuffle


def test_align_mirrors_reference_semantics():
    import avsl_b200 as A
    a, v = np.zeros((10, 104)), np.zeros((8, 88, 88, 1))
    a2, v2 = A.align_audio_video_features(a, v)
    assert a2.shape[0] == 8 and v2.shape[0] == 8 and a2.base is a
    a3, v3 = A.align_audio_video_features(a[:5], v)
    assert a3.shape[0] == 5 and v3.shape[0] == 5
    assert A.align_audio_video_features(None, v) == (None, v)
    np.testing.assert_array_equal(A.aligned_lengths([10, 5, 7], [8, 8, 7]), [8, 5, 7])


@pytest.mark.gpu
def test_gpu_time_warp_matches_oracle_bit_for_bit():
    import avsl_b200 as A
    from oracle import specaug as OS
    rng = np.random.default_rng(2)
    mel = rng.standard_normal((5, 80, 3000)).astype(np.float32)
    pts = A.spec_augment_warp_points([3000, 1234, 700, 150, 2999], W=80, rng=rng)
    pts[4] = (2999, 500, 420)
    got = A.spec_time_warp(torch.from_numpy(mel).cuda(), pts).cpu().numpy()
    np.testing.assert_array_equal(got, OS.time_warp(mel, pts))
    assert not np.array_equal(got[0], mel[0]) and np.array_equal(got[3], mel[3])
    with pytest.raises(ValueError):
        x = torch.from_numpy(mel).cuda()
        A.spec_time_warp(x, pts, out=x)


@pytest.mark.gpu
def test_gpu_process_audio_dual_encoder():
    import avsl_b200 as A
    from avsl_b200 import synth
    from oracle import logfbank as OL
    from oracle import logmel as OM
    a = synth.audio_clip(24000, 3) * 4.0                 # some samples beyond [-1, 1]: the waveform gets peak-normalised
    d = A.process_audio_dual_encoder(a, stack_order=4, normalize=True)
    assert d["sample_rate"] == 16000 and d["waveform"].dtype == np.float32
    np.testing.assert_array_equal(d["waveform"], OM.peak_normalize(a))
    ref = OL.audio_to_tensor(OL.extract_logfbank_features(a, stack_order=4), normalize=True)
    assert d["av_hubert_features"].shape == ref.shape
    assert np.abs(d["av_hubert_features"] - ref).max() <= 2e-3
    dev = A.process_audio_dual_encoder(torch.from_numpy(a).cuda(), stack_order=4)
    assert dev["waveform"].is_cuda and dev["av_hubert_features"].is_cuda
