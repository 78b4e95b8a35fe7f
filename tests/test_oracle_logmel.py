"""Pin the log-mel oracle (oracle/logmel.py) against the reference's extractor
(transformers.WhisperFeatureExtractor, avsl/whisper_ft.py:347-350): live when transformers is
importable and through the committed golden vectors."""
import numpy as np
import pytest
import torch

from avsl_b200 import synth
from oracle import logmel as O

from conftest import GOLDEN


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN / "logmel_golden.npz")


def _clip(name, gold):
    a = gold[f"audio_{name}"]
    if a.size:
        return a
    return {"noise30": lambda: synth.audio_clip(480000, 3407),
            "chirp30": lambda: synth.chirp_silence_clip(480000),
            "short7s": lambda: synth.audio_clip(112000, 11) * 3.0}[name]()


@pytest.mark.parametrize("n_mels", [80, 128])
def test_filterbank_matches_reference(gold, n_mels):
    fb = O.mel_filters(n_mels)
    assert fb.shape == (n_mels, 201) and fb.dtype == np.float32
    np.testing.assert_array_equal(fb, gold[f"filters_{n_mels}"])
    # triangular sparsity quoted in SURVEY.md 7
    assert int((fb != 0).sum()) == {80: 391, 128: 394}[n_mels]


@pytest.mark.parametrize("n_mels", [80, 128])
@pytest.mark.parametrize("name", ["noise30", "chirp30", "short7s"])
def test_oracle_matches_golden_30s(gold, name, n_mels):
    a = O.pad_or_trim(_clip(name, gold), 480000)
    out = O.log_mel_spectrogram(a, n_mels).numpy()
    assert out.shape == (n_mels, 3000) and out.dtype == np.float32
    np.testing.assert_allclose(out[:, gold["frame_sel"]], gold[f"{name}_{n_mels}"], rtol=0, atol=1e-6)


@pytest.mark.parametrize("n_mels", [80, 128])
@pytest.mark.parametrize("name", ["s1", "s2", "s3"])
def test_oracle_matches_golden_short(gold, name, n_mels):
    a = gold[f"audio_{name}"]
    out = O.log_mel_spectrogram(a, n_mels).numpy()
    assert out.shape == (n_mels, len(a) // 160)
    np.testing.assert_allclose(out, gold[f"{name}_{n_mels}"], rtol=0, atol=1e-6)


def test_oracle_matches_live_extractor():
    tr = pytest.importorskip("transformers")
    fe = tr.WhisperFeatureExtractor(feature_size=80)
    a = synth.audio_clip(480000, 123)
    ref = fe(a, sampling_rate=16000, return_tensors="np").input_features[0]
    out = O.log_mel_spectrogram(a, 80).numpy()
    assert np.abs(out - ref).max() <= 1e-6


def test_padding_argument_equals_explicit_pad():
    a = synth.audio_clip(20000, 1)
    x = O.log_mel_spectrogram(a, 80, padding=12000)
    y = O.log_mel_spectrogram(np.pad(a, (0, 12000)), 80)
    assert x.shape == (80, 200)
    assert torch.equal(x, y)


def test_per_clip_max_in_batches():
    b = synth.audio_batch(3, 16000, 5)
    batched = O.log_mel_spectrogram(b, 80)
    for i in range(3):
        assert torch.equal(batched[i], O.log_mel_spectrogram(b[i], 80))


def test_f64_truth_close_to_f32_oracle():
    a = synth.audio_clip(32000, 9)
    f32 = O.log_mel_spectrogram(a, 80).numpy()
    f64 = O.log_mel_spectrogram_f64(a, 80)
    assert np.abs(f32 - f64).max() < 2e-5


def test_pad_or_trim():
    a = np.arange(10, dtype=np.float32)
    assert O.pad_or_trim(a, 4).tolist() == [0, 1, 2, 3]
    assert O.pad_or_trim(a, 12).tolist() == list(range(10)) + [0, 0]
    t = O.pad_or_trim(torch.arange(6.0).view(2, 3), 5)
    assert t.shape == (2, 5) and t[1].tolist() == [3, 4, 5, 0, 0]


def test_peak_normalize():
    a = np.array([0.5, -2.0, 1.0], dtype=np.float32)
    np.testing.assert_array_equal(O.peak_normalize(a), a / np.float32(2.0))
    b = np.array([0.5, -1.0, 1.0], dtype=np.float32)
    np.testing.assert_array_equal(O.peak_normalize(b), b)
