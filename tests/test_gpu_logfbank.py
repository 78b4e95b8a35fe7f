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


def _oracle_logfbank(a, fb):
    """oracle/logfbank.py's pipeline with an arbitrary filterbank (float64)."""
    sig = np.append(a[0], a[1:] - 0.97 * a[:-1]).astype(np.float64)
    nfr = OF.num_frames(len(a))
    sig = np.concatenate([sig, np.zeros((nfr - 1) * 160 + 400 - len(sig))])
    idx = np.arange(400)[None, :] + 160 * np.arange(nfr)[:, None]
    P = np.abs(np.fft.rfft(sig[idx], 512)) ** 2 / 512.0
    e = P @ fb.astype(np.float64).T
    return np.log(np.where(e == 0, np.finfo(float).eps, e))


@pytest.mark.parametrize("stack", [1, 2, 8, 16])
def test_forty_filters_and_other_stack_orders(stack):
    """nfilt = 40 (the filter table's capacity: five filters per warp) and every stack order a
    16-frame tile divides by, on a ragged batch whose clips start at odd offsets."""
    lens = [16000 + 3, 4801, 399, 33333]
    clips = [synth.audio_clip(n, 40 + i) * 1.5 for i, n in enumerate(lens)]
    off = np.concatenate([[0], np.cumsum(lens)])
    fb = OF.get_filterbanks(40).astype(np.float32)
    feats, row_off = A.logfbank_batch(torch.from_numpy(np.concatenate(clips)).cuda(), off, stack, False, nfilt=40)
    feats = feats.cpu().numpy()
    for i, c in enumerate(clips):
        raw = _oracle_logfbank(c, fb)
        pad = (-len(raw)) % stack
        ref = np.concatenate([raw, np.zeros((pad, 40))]).reshape(-1, 40 * stack)
        assert row_off[i + 1] - row_off[i] == len(ref)
        assert np.abs(feats[row_off[i]:row_off[i + 1]] - ref).max() < 2e-4, (i, stack)


def test_dense_filterbank_takes_the_unpacked_path():
    """A filterbank without zeros (26 x 257 weights do not fit the packed table): weights are read
    from global memory, supports are the whole spectrum."""
    from avsl_b200 import _lib
    rng = np.random.default_rng(9)
    fb = (rng.random((26, 257)) * 0.01 + 1e-4).astype(np.float32)
    a = synth.audio_clip(16000 + 77, 3) * 2.0
    plan = A.LogfbankPlan([0, len(a)], 1, 26, torch.device("cuda"))
    d_a, d_fb = torch.from_numpy(a).cuda(), torch.from_numpy(fb).cuda()
    _lib.call("avfe_logfbank_f32", _lib.ptr(d_a), _lib.ptr(plan.d_off), _lib.ptr(plan.d_row), 1, plan.max_len,
              _lib.ptr(d_fb), 26, 1, 0, _lib.ptr(plan.out), _lib.ptr(plan.ws), plan.ws.numel(), _lib.stream_ptr())
    got = plan.out.cpu().numpy()
    assert np.abs(got - _oracle_logfbank(a, fb)).max() < 2e-4
