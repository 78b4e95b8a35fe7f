"""SpecAugment masks: host-side band sampling (CPU tier) and the in-place GPU application."""
import numpy as np
import pytest

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
