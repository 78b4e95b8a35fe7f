"""Host-side logic that needs no GPU: filterbank construction, packing, sharding, modality
dropout draws, synthetic generators, byte accounting."""
import numpy as np
import pytest
import torch

from avsl_b200 import frontend, fusion, synth
from avsl_b200.audio import _mel_filters_np
from oracle import fusion as OF
from oracle import logmel as OM

from conftest import GOLDEN


@pytest.mark.parametrize("n_mels", [80, 128])
def test_product_filterbank_equals_reference_and_oracle(n_mels):
    gold = np.load(GOLDEN / "logmel_golden.npz")
    fb = _mel_filters_np(n_mels)
    assert fb.dtype == np.float32 and fb.flags.c_contiguous
    np.testing.assert_array_equal(fb, gold[f"filters_{n_mels}"])      # HF / librosa slaney bank
    np.testing.assert_array_equal(fb, OM.mel_filters(n_mels))         # independent scalar restatement


def test_pack_utterances_layout_and_trim():
    audios = [np.arange(10, dtype=np.float32), np.arange(5, dtype=np.float32)]
    vids = [np.zeros((3, 4, 4, 3), np.uint8), np.ones((2, 4, 4, 3), np.uint8)]
    lms = [np.zeros((3, 68, 2)), np.ones((2, 68, 2))]
    vals = [np.array([1, 0, 1], np.uint8), np.array([1, 1], np.uint8)]
    b = frontend.pack_utterances(audios, vids, lms, vals, audio_max_length=8)
    assert b.n_utts == 2
    assert b.audio_offsets.tolist() == [0, 8, 13]                     # first clip trimmed to 8 samples
    assert b.clip_offsets.tolist() == [0, 3, 5]   # video is NOT cut here: the reference trims the features, after the lip path
    b = frontend.pack_utterances(audios, vids, lms, vals, audio_max_length=480000)
    assert b.clip_offsets.tolist() == [0, 3, 5]
    assert b.frames.shape == (5, 4, 4, 3) and b.landmarks.shape == (5, 68, 2)
    assert b.lm_valid.tolist() == [1, 0, 1, 1, 1]
    assert b.nbytes() == 15 * 4 + 3 * 8 * 2 + 5 * 48 + 5 * 136 * 8 + 5
    long_vid = np.zeros((800, 2, 2, 3), np.uint8)
    b = frontend.pack_utterances([np.zeros(480000, np.float32)], [long_vid], [np.zeros((800, 68, 2))])
    assert b.clip_offsets.tolist() == [0, 800]                        # all frames go through extract_lip_frames ...
    lip = torch.zeros(800, 88, 88, 1)                                 # ... and the trim is applied to the features
    assert [tuple(x.shape) for x in frontend.AVFrontEnd.split_lip(lip, b.clip_offsets, 480000)] == [(750, 88, 88, 1)]
    assert [tuple(x.shape) for x in frontend.AVFrontEnd.split_lip(lip, b.clip_offsets)] == [(800, 88, 88, 1)]


def test_shard_is_a_disjoint_cover():
    for n, w in [(10, 1), (10, 3), (7, 8), (10000, 8)]:
        parts = [frontend.shard(n, r, w) for r in range(w)]
        allidx = np.sort(np.concatenate(parts))
        np.testing.assert_array_equal(allidx, np.arange(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        frontend.shard(10, 3, 3)


def test_shard_balanced_partitions_and_balances():
    durs = synth.ami_durations(10000, 3407)
    for w in (1, 2, 8):
        parts = [frontend.shard_balanced(durs, r, w) for r in range(w)]
        np.testing.assert_array_equal(np.sort(np.concatenate(parts)), np.arange(len(durs)))
        loads = np.array([durs[p].sum() for p in parts])
        assert loads.max() - loads.min() <= durs.max() + 1e-9
        assert all((np.diff(p) > 0).all() for p in parts)
        # strided shards of the same list are less even than the balanced ones
        strided = np.array([durs[frontend.shard(len(durs), r, w)].sum() for r in range(w)])
        assert loads.max() - loads.min() <= strided.max() - strided.min() + 1e-9
    np.testing.assert_array_equal(frontend.shard_balanced([], 0, 2), np.zeros(0, dtype=np.int64))
    np.testing.assert_array_equal(frontend.shard_balanced([5.0, 1.0, 1.0, 3.0], 0, 2), [0])       # 5 | 3 + 1 + 1
    np.testing.assert_array_equal(frontend.shard_balanced([5.0, 1.0, 1.0, 3.0], 1, 2), [1, 2, 3])
    with pytest.raises(ValueError):
        frontend.shard_balanced([1.0], 2, 2)


def test_modality_dropout_draws_match_reference_block():
    """Same RNG stream, same decisions as av_hubert_encoder.py:292-298 (two draws per forward,
    even in eval)."""
    for training in (True, False):
        np.random.seed(3407)
        got = [fusion.modality_dropout_flags(training, 0.5, 0.5) for _ in range(200)]
        np.random.seed(3407)
        ref = [OF.modality_dropout_flags(training, 0.5, 0.5) for _ in range(200)]
        assert got == ref
        if training:
            assert {(True, True), (False, True), (True, False)} == set(got)
        else:
            assert set(got) == {(True, True)}
    rng = np.random.default_rng(0)
    m = fusion.modality_dropout_mask(64, True, 0.5, 0.5, rng, per_sample=True)
    assert m.shape == (64, 2) and m.dtype == np.uint8 and (m.sum(axis=1) >= 1).all()
    assert 5 < (m.sum(axis=1) == 1).sum() < 59
    m = fusion.modality_dropout_mask(8, True, 1.0, 1.0, rng)            # whole batch drops audio
    assert (m == np.array([[0, 1]] * 8)).all()


def test_fuse_validation_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(ValueError, match="Unsupported fusion type"):
        fusion.fuse_modalities(torch.zeros(1, 1, 1), torch.zeros(1, 1, 1), None, "gated")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fusion.fuse_modalities(torch.zeros(1, 1, 1), torch.zeros(1, 1, 1), None, "add")


def test_synth_generators_are_seeded():
    np.testing.assert_array_equal(synth.audio_clip(1000, 5), synth.audio_clip(1000, 5))
    assert not np.array_equal(synth.audio_clip(1000, 5), synth.audio_clip(1000, 6))
    f1, l1, v1 = synth.video_clip(3, 32, 48, seed=9)
    f2, l2, v2 = synth.video_clip(3, 32, 48, seed=9)
    np.testing.assert_array_equal(f1, f2)
    np.testing.assert_array_equal(l1, l2)
    assert f1.shape == (3, 32, 48, 3) and f1.dtype == np.uint8 and l1.shape == (3, 68, 2)
    assert np.array_equal(l1, np.rint(l1))                             # dlib-style integer detections
    x = synth.chirp_silence_clip(32000)
    assert x[:16000].std() > 0.1 and not x[16000:].any()
    d = synth.ami_durations(10000)
    assert d.min() >= 0.28 and d.max() <= 30.0 and 3.5 < d.mean() < 5.5
    np.testing.assert_allclose(np.round(d * 25), d * 25, atol=1e-9)    # whole video frames
    fa, fv, mask = synth.fusion_inputs(16, 4, 5)
    assert (mask.sum(axis=1) >= 1).all() and mask.min() == 0


def test_algorithmic_bytes_match_survey():
    # 2,880,000 B per 30 s clip (n_mels 80), 232,768 B per 224x224 frame
    assert frontend.algorithmic_bytes(1, 0) == 2_880_000
    assert frontend.algorithmic_bytes(0, 1) == 232_768
    assert frontend.algorithmic_bytes(1, 0, n_mels=128) == 3_456_000
    assert frontend.algorithmic_bytes(64, 250) == 64 * 2_880_000 + 250 * 232_768
