"""GPU parity: log-mel through the C ABI (avsl_b200.audio -> libavfe.so) vs the oracle and the
golden vectors.  Tolerance (north_star): max-abs <= 1e-4 on the normalised log-mel."""
import numpy as np
import pytest
import torch

import avsl_b200 as A
from avsl_b200 import synth
from oracle import logmel as O

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN / "logmel_golden.npz")


@pytest.mark.parametrize("n_mels", [80, 128])
def test_config1_single_30s_clip(n_mels):
    a = synth.audio_clip(480000, 3407)
    out = A.log_mel_spectrogram(a, n_mels=n_mels)
    assert not out.is_cuda and out.shape == (n_mels, 3000) and out.dtype == torch.float32
    ref = O.log_mel_spectrogram(a, n_mels)
    assert (out - ref).abs().max().item() <= TOL


@pytest.mark.parametrize("n_mels", [80, 128])
@pytest.mark.parametrize("name", ["noise30", "chirp30", "short7s"])
def test_golden_30s(gold, name, n_mels):
    a = {"noise30": lambda: synth.audio_clip(480000, 3407),
         "chirp30": lambda: synth.chirp_silence_clip(480000),
         "short7s": lambda: synth.audio_clip(112000, 11) * 3.0}[name]()
    a = A.pad_or_trim(a, 480000)
    assert isinstance(a, np.ndarray) and a.shape == (480000,)
    out = A.log_mel_spectrogram(a, n_mels=n_mels).numpy()
    assert np.abs(out[:, gold["frame_sel"]] - gold[f"{name}_{n_mels}"]).max() <= TOL


@pytest.mark.parametrize("n_mels", [80, 128])
@pytest.mark.parametrize("name", ["s1", "s2", "s3"])
def test_golden_short_ragged(gold, name, n_mels):
    a = gold[f"audio_{name}"]
    out = A.log_mel_spectrogram(a, n_mels=n_mels).numpy()
    assert out.shape == gold[f"{name}_{n_mels}"].shape
    assert np.abs(out - gold[f"{name}_{n_mels}"]).max() <= TOL


def test_chirp_silence_against_f64_truth():
    """On the high-dynamic-range clip the kernel must be no further from exact arithmetic than
    the float32 torch oracle is (both are float32 pipelines)."""
    a = synth.chirp_silence_clip(160000)
    out = A.log_mel_spectrogram(a, 80).numpy().astype(np.float64)
    f32 = O.log_mel_spectrogram(a, 80).numpy().astype(np.float64)
    f64 = O.log_mel_spectrogram_f64(a, 80)
    assert np.abs(out - f64).max() <= max(TOL, 2 * np.abs(f32 - f64).max())
    assert np.abs(out - f32).max() <= TOL


@pytest.mark.parametrize("n_mels", [80, 128])
def test_config3_batch64_every_clip(n_mels):
    """BASELINE configs[2] at full size: 64 x 30 s, every clip of the batch against the oracle
    (per-clip maxima differ by construction of the batch), both filterbanks."""
    b = synth.audio_batch(64, 480000, 3407)
    out = A.log_mel_spectrogram(b.cuda(), n_mels=n_mels)
    assert out.is_cuda and out.shape == (64, n_mels, 3000)
    out = out.cpu()
    worst = 0.0
    for i in range(0, 64, 8):                      # the oracle runs 8 clips at a time (memory)
        ref = torch.stack([O.log_mel_spectrogram(b[k], n_mels) for k in range(i, i + 8)])
        worst = max(worst, (out[i:i + 8] - ref).abs().max().item())
    assert worst <= TOL


def test_padding_argument_and_leading_dims():
    a = synth.audio_batch(6, 20000, 1).view(2, 3, 20000)
    out = A.log_mel_spectrogram(a, n_mels=80, padding=12000)
    assert out.shape == (2, 3, 80, 200)
    ref = O.log_mel_spectrogram(a.reshape(6, -1), 80, padding=12000).view(2, 3, 80, 200)
    assert (out - ref).abs().max().item() <= TOL


def test_edge_lengths():
    # torch.stft refuses reflect padding of 200 on <= 200 samples; so does the drop-in
    for n, pad in [(100, 0), (100, 60), (200, 0)]:
        with pytest.raises(RuntimeError):
            A.log_mel_spectrogram(np.ones(n, np.float32), 80, padding=pad)
        with pytest.raises(RuntimeError):
            O.log_mel_spectrogram(np.ones(n, np.float32), 80, padding=pad)
    assert A.log_mel_spectrogram(np.ones(201, np.float32), 80).shape == (80, 1)
    # all-zero clip: log10(1e-10) = -10 everywhere -> (-10 + 4) / 4
    z = A.log_mel_spectrogram(np.zeros(16000, np.float32), 80)
    assert torch.equal(z, torch.full((80, 100), -1.5))
    # 201..399 samples: reflect padding reaches almost the whole clip
    a = synth.audio_clip(333, 2)
    assert (A.log_mel_spectrogram(a, 80) - O.log_mel_spectrogram(a, 80)).abs().max().item() <= TOL
    with pytest.raises(ValueError):
        A.log_mel_spectrogram(a, n_mels=129)


def test_linearity_property_full_size():
    """Size-independent property at BASELINE size: scaling the waveform by c shifts the raw
    log-power by 2*log10(c), so the normalised output is unchanged (until the 1e-10 floor)."""
    a = synth.audio_batch(8, 480000, 7).cuda()
    x = A.log_mel_spectrogram(a, 80)
    y = A.log_mel_spectrogram(a * 4.0, 80)
    assert (y - x - (2 * np.log10(4.0) / 4.0)).abs().max().item() <= 5e-5


def test_pad_or_trim_and_peak_normalize():
    a = synth.audio_clip(1000, 3)
    np.testing.assert_array_equal(A.pad_or_trim(a, 400), O.pad_or_trim(a, 400))
    np.testing.assert_array_equal(A.pad_or_trim(a, 1500), O.pad_or_trim(a, 1500))
    t = torch.from_numpy(a).view(2, 500).cuda()
    assert torch.equal(A.pad_or_trim(t, 700).cpu(), O.pad_or_trim(t.cpu(), 700))
    loud = (a * 30.0).astype(np.float32)
    np.testing.assert_array_equal(A.peak_normalize(loud), O.peak_normalize(loud))
    np.testing.assert_array_equal(A.peak_normalize(a), O.peak_normalize(a))
    both = np.stack([loud, a])
    np.testing.assert_array_equal(A.peak_normalize(both), np.stack([O.peak_normalize(loud), O.peak_normalize(a)]))


def test_unprepared_entry_point_and_custom_filterbanks():
    """avfe_logmel_f32 (dense filters analysed on every call) and filterbanks that are not the
    sparse slaney triangles (dense rows -> generic projection path; wide triangles -> table with
    coarser tiers) give the oracle's numbers too."""
    from avsl_b200 import _lib
    a = synth.audio_batch(3, 48000, 11)
    rng = np.random.default_rng(0)
    banks = {
        "slaney80": O.mel_filters(80),
        "dense20": np.abs(rng.normal(size=(20, 201))).astype(np.float32) * 1e-2,
        "wide12": np.maximum(0, 1 - np.abs(np.arange(201)[None, :] - np.linspace(8, 190, 12)[:, None]) / 7.5).astype(np.float32),
        "with_empty_row": np.concatenate([O.mel_filters(80)[:5], np.zeros((1, 201), np.float32)]),
    }
    for name, fb in banks.items():
        n_mels = fb.shape[0]
        ref = O.log_mel_spectrogram(a, n_mels, filters=fb)
        got = A.log_mel_spectrogram(a, n_mels, filters=torch.from_numpy(fb).cuda()).cpu()
        assert (got - ref).abs().max().item() <= TOL, name
        # raw C-ABI entry without a prepared pack
        d_a, d_fb = a.cuda(), torch.from_numpy(fb).cuda()
        out = torch.empty((3, n_mels, 300), device="cuda")
        lib = _lib.load()
        nbytes = int(lib.avfe_logmel_workspace_bytes(3, 48000, 0, n_mels))
        ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        _lib.call("avfe_logmel_f32", _lib.ptr(d_a), 3, 48000, 0, n_mels, _lib.ptr(d_fb), _lib.ptr(out),
                  _lib.ptr(ws), nbytes, _lib.stream_ptr())
        assert torch.equal(out.cpu(), got), name


def test_ragged_fused_pad_or_trim():
    """avfe_logmel_ragged_f32 == pad_or_trim + log_mel per clip, for lengths that are shorter,
    equal and longer than the target, unaligned clip starts, a clip that is pure padding after a
    few samples, and a clip of digital silence."""
    L = 48000
    lens = [48000, 12345, 60000, 201, 7, 30001]
    clips = [synth.audio_clip(n, 40 + i) for i, n in enumerate(lens)]
    clips[4] = np.zeros(7, np.float32) + 0.25
    clips.append(np.zeros(20000, np.float32))
    lens.append(20000)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    audio = torch.from_numpy(np.concatenate(clips)).cuda()
    for n_mels in (80, 128):
        got = A.log_mel_spectrogram_ragged(audio, torch.from_numpy(off).cuda(), L, n_mels).cpu()
        assert got.shape == (len(lens), n_mels, 300)
        for i, c in enumerate(clips):
            ref = O.log_mel_spectrogram(O.pad_or_trim(c, L), n_mels)
            assert (got[i] - ref).abs().max().item() <= TOL, (n_mels, i)
    # the dense entry with a `padding` argument takes the same silent-by-length shortcut
    a = synth.audio_clip(8000, 3)
    x = A.log_mel_spectrogram(a, 80, padding=40000)
    assert (x - O.log_mel_spectrogram(a, 80, padding=40000)).abs().max().item() <= TOL
