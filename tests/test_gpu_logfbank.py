"""GPU parity: AV-HuBERT audio features through the C ABI vs the oracle.

Tolerance: the oracle evaluates the transform in float64 (numpy), the kernel in float32; log
energies agree to 2e-4 absolute (measured ~2e-5 on noise), normalised rows to 1e-3."""
import numpy as np
import pytest
import torch

import avsl_b200 as A
from avsl_b200 import synth
from oracle import logfbank as OF

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("L", [16000, 48000 + 77, 401, 400, 37, 1])
@pytest.mark.parametrize("stack", [1, 4])
def test_single_clip_vs_oracle(L, stack):
    a = synth.audio_clip(L, 11) * 2.0
    got = A.extract_logfbank_features(a, stack_order=stack)
    ref = OF.extract_logfbank_features(a, stack_order=stack)
    assert got.shape == ref.shape and got.dtype == np.float32
    assert np.abs(got - ref).max() < 2e-4
    got_n = A.extract_logfbank_features(a, stack_order=stack, normalize=True)
    ref_n = OF.audio_to_tensor(ref)
    assert np.abs(got_n - ref_n).max() < 1e-3


def test_ragged_batch_and_silence():
    lens = [16000 * 3, 5000, 401, 16000 * 10 + 5, 800]
    clips = [synth.audio_clip(n, 20 + i) for i, n in enumerate(lens)]
    clips[4][:] = 0.0                                          # digital silence: log(eps) everywhere
    off = np.concatenate([[0], np.cumsum(lens)])
    packed = torch.from_numpy(np.concatenate(clips)).cuda()
    feats, row_off = A.logfbank_batch(packed, off, stack_order=4, normalize=False)
    feats = feats.cpu().numpy()
    for i, c in enumerate(clips):
        ref = OF.extract_logfbank_features(c, stack_order=4)
        assert row_off[i + 1] - row_off[i] == len(ref)
        assert np.abs(feats[row_off[i]:row_off[i + 1]] - ref).max() < 2e-4, i
    sil = feats[row_off[4]:row_off[5]]
    frames = OF.num_frames(800)
    np.testing.assert_allclose(sil.reshape(-1, 26)[:frames], np.log(np.finfo(float).eps), rtol=1e-6)
    # clips are independent: the same clip gives the same bytes wherever it sits
    again, ro2 = A.logfbank_batch(torch.from_numpy(np.concatenate(clips[::-1])).cuda(),
                                  np.concatenate([[0], np.cumsum(lens[::-1])]), 4, False)
    np.testing.assert_array_equal(again.cpu().numpy()[ro2[4]:ro2[5]], feats[row_off[0]:row_off[1]])


def test_full_size_properties():
    """64 x 30 s: frame count, normalised rows have zero mean / unit std, linear scaling of the
    input shifts every log energy by 2 log(s)."""
    B, L = 64, 480000
    audio = synth.audio_batch(B, L, 3407, device="cuda").reshape(-1)
    off = np.arange(B + 1) * L
    feats, row_off = A.logfbank_batch(audio, off, 4, True)
    rows = -(-A.logfbank_num_frames(L) // 4)
    assert feats.shape == (B * rows, 104) and row_off[-1] == B * rows
    full = feats.view(B, rows, 104)[:, :-1]                     # last row of a clip may hold padded frames
    assert full.mean(dim=-1).abs().max().item() < 1e-4
    assert (full.std(dim=-1, unbiased=False) - 1.0).abs().max().item() < 1e-3
    raw, _ = A.logfbank_batch(audio, off, 1, False)
    raw2, _ = A.logfbank_batch(audio * 0.5, off, 1, False)
    assert (raw2 - raw - 2.0 * np.log(0.5)).abs().max().item() < 1e-3
