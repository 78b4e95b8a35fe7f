"""Oracle for the per-sample trim and the batch collation (CPU tier)."""
import numpy as np

from oracle import collate as C


def test_trim_rule_is_python_round():
    # avsl/whisper_flamingo_ft_ami.py:299: round(len(audio) / 16000 * 25), half to even
    assert C.video_frames_for_audio(480000) == 750
    assert C.video_frames_for_audio(160000) == 250
    assert C.video_frames_for_audio(16000 * 3 + 320) == 76        # 75.5 -> 76 (even)
    assert C.video_frames_for_audio(16000 * 3 - 320) == 74        # 74.5 -> 74 (even)
    v = np.arange(10 * 2 * 2, dtype=np.float32).reshape(10, 2, 2, 1)
    assert len(C.trim_video(v, 16000 // 25 * 4)) == 4
    assert C.trim_video(v, 480000) is v


def test_align_truncates_the_longer():
    a, v = np.zeros((7, 3)), np.zeros((5, 2))
    a2, v2 = C.align_audio_video_features(a, v)
    assert len(a2) == 5 and len(v2) == 5
    a2, v2 = C.align_audio_video_features(v, a)
    assert len(a2) == 5 and len(v2) == 5
    assert C.align_audio_video_features(None, v) == (None, v)


def test_collate_pads_with_zeros_and_masks():
    rng = np.random.default_rng(0)
    vids = [rng.standard_normal((t, 88, 88, 1)).astype(np.float32) for t in (3, 5, 1)]
    out = C.collate_video(vids, mels=[np.zeros((80, 3000), np.float32)] * 3)
    assert out["video"].shape == (3, 1, 5, 88, 88) and out["video"].dtype == np.float32
    assert out["padding_mask"].shape == (3, 5) and out["padding_mask"].dtype == bool
    np.testing.assert_array_equal(out["padding_mask"], [[0, 0, 0, 1, 1], [0] * 5, [0, 1, 1, 1, 1]])
    np.testing.assert_array_equal(out["video"][0, 0, :3], vids[0][..., 0])
    assert not out["video"][0, 0, 3:].any() and not out["video"][2, 0, 1:].any()
    assert out["input_ids"].shape == (3, 80, 3000)
    assert C.collate_video(vids, T_pad=8)["video"].shape == (3, 1, 8, 88, 88)
