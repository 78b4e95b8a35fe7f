"""Oracle: Whisper log-mel spectrogram (CPU).  TEST INFRASTRUCTURE ONLY.

Restates ``whisper.log_mel_spectrogram(audio, n_mels, padding)`` as it is called at
``avsl/whisper_flamingo_ft_ami.py:209-213`` (upstream ``whisper_flamingo/whisper/audio.py``,
a fork of openai/whisper, is not vendored in the reference snapshot) and its twin
``WhisperFeatureExtractor.__call__`` used at ``avsl/whisper_ft.py:347-350``.

Pinned: ``tests/test_oracle_logmel.py`` checks this file against
``transformers.WhisperFeatureExtractor`` (live when importable, and through the committed
golden vectors in ``tests/golden/logmel_*.npz``).
"""
from __future__ import annotations

import math

import numpy as np
import torch

SAMPLE_RATE = 16000          # avsl/whisper_flamingo_ft_ami.py:147
N_FFT = 400
HOP_LENGTH = 160             # avsl/whisper_flamingo_ft_ami.py:148
N_FREQ = N_FFT // 2 + 1
N_SAMPLES = 480000           # 30 s, whisper.model N_SAMPLES (whisper_flamingo_ft_ami.py:493)


# --------------------------------------------------------------------------------------
# Slaney mel filterbank == librosa.filters.mel(sr=16000, n_fft=400, n_mels=M), which is what
# openai-whisper's assets/mel_filters.npz was generated from, and what HF builds with
# mel_filter_bank(201, M, 0, 8000, 16000, norm="slaney", mel_scale="slaney").
# Written as explicit scalar loops on purpose (independent of the vectorised product code).
# --------------------------------------------------------------------------------------
def _hz_to_mel_slaney(f: float) -> float:
    f_sp = 200.0 / 3.0
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    if f >= min_log_hz:
        return min_log_mel + math.log(f / min_log_hz) / logstep
    return f / f_sp


def _mel_to_hz_slaney(m: float) -> float:
    f_sp = 200.0 / 3.0
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    if m >= min_log_mel:
        return min_log_hz * math.exp(logstep * (m - min_log_mel))
    return f_sp * m


def mel_filters(n_mels: int) -> np.ndarray:
    """[n_mels, 201] float32 slaney-normalised triangular filterbank, 0..8000 Hz."""
    fft_freqs = [k * (SAMPLE_RATE / 2.0) / (N_FREQ - 1) for k in range(N_FREQ)]
    m_lo, m_hi = _hz_to_mel_slaney(0.0), _hz_to_mel_slaney(SAMPLE_RATE / 2.0)
    mel_pts = [_mel_to_hz_slaney(m_lo + (m_hi - m_lo) * i / (n_mels + 1)) for i in range(n_mels + 2)]
    fb = np.zeros((n_mels, N_FREQ), dtype=np.float64)
    for m in range(n_mels):
        lo, ce, hi = mel_pts[m], mel_pts[m + 1], mel_pts[m + 2]
        enorm = 2.0 / (hi - lo)
        for k, f in enumerate(fft_freqs):
            down = (f - lo) / (ce - lo)
            up = (hi - f) / (hi - ce)
            w = max(0.0, min(down, up))
            fb[m, k] = w * enorm
    return fb.astype(np.float32)


def pad_or_trim(array, length: int = N_SAMPLES, axis: int = -1):
    """whisper.pad_or_trim (call: avsl/whisper_flamingo_ft_ami.py:209-210): zero-pad or cut."""
    is_t = torch.is_tensor(array)
    a = array if is_t else np.asarray(array)
    n = a.shape[axis]
    if n > length:
        idx = [slice(None)] * a.ndim
        idx[axis] = slice(0, length)
        a = a[tuple(idx)]
    elif n < length:
        if is_t:
            pad = [0, 0] * a.ndim
            pad[2 * (a.ndim - 1 - (axis % a.ndim)) + 1] = length - n
            a = torch.nn.functional.pad(a, pad)
        else:
            widths = [(0, 0)] * a.ndim
            widths[axis] = (0, length - n)
            a = np.pad(a, widths)
    return a


def log_mel_spectrogram(audio, n_mels: int = 80, padding: int = 0, filters=None) -> torch.Tensor:
    """float32 [..., n_mels, (L+padding)//160].

    Same op sequence as upstream: F.pad(0,padding) -> torch.stft(400,160,hann periodic,
    center/reflect) -> drop last frame -> |.|^2 -> filters @ -> clamp(1e-10).log10() ->
    max(x, x.max()-8) -> (x+4)/4.  The max is PER CLIP: the reference only ever calls this
    per sample (avsl/whisper_flamingo_ft_ami.py:213) and HF reduces per row.
    """
    if not torch.is_tensor(audio):
        audio = torch.from_numpy(np.ascontiguousarray(audio))
    audio = audio.to(torch.float32)
    if padding > 0:
        audio = torch.nn.functional.pad(audio, (0, padding))
    window = torch.hann_window(N_FFT)
    stft = torch.stft(audio, N_FFT, HOP_LENGTH, window=window, return_complex=True)
    magnitudes = stft[..., :-1].abs() ** 2
    if filters is None:
        filters = mel_filters(n_mels)
    filters = torch.as_tensor(filters, dtype=torch.float32)
    mel_spec = filters @ magnitudes
    log_spec = torch.clamp(mel_spec, min=1e-10).log10()
    if log_spec.dim() == 2:
        mx = log_spec.max()
    else:
        mx = log_spec.amax(dim=(-2, -1), keepdim=True)
    log_spec = torch.maximum(log_spec, mx - 8.0)
    log_spec = (log_spec + 4.0) / 4.0
    return log_spec


def log_mel_spectrogram_f64(audio, n_mels: int = 80, padding: int = 0, filters=None) -> np.ndarray:
    """Same pipeline evaluated in float64 (numpy rfft): the 'truth' used to measure how far
    the float32 oracle and the CUDA kernel each sit from exact arithmetic."""
    a = np.asarray(audio, dtype=np.float32).astype(np.float64)
    if padding > 0:
        a = np.pad(a, [(0, 0)] * (a.ndim - 1) + [(0, padding)])
    lead = a.shape[:-1]
    a = a.reshape(-1, a.shape[-1])
    n = np.arange(N_FFT)
    window = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / N_FFT)
    a = np.pad(a, [(0, 0), (N_FFT // 2, N_FFT // 2)], mode="reflect")
    n_frames = (a.shape[-1] - N_FFT) // HOP_LENGTH + 1 - 1   # drop the last frame
    idx = np.arange(n_frames)[:, None] * HOP_LENGTH + n[None, :]
    fb = (mel_filters(n_mels) if filters is None else np.asarray(filters)).astype(np.float64)
    out = np.empty((a.shape[0], n_mels, n_frames), dtype=np.float64)
    for b in range(a.shape[0]):
        spec = np.fft.rfft(a[b][idx] * window, axis=-1)          # [frames, 201]
        power = spec.real ** 2 + spec.imag ** 2
        mel = fb @ power.T
        ls = np.log10(np.maximum(mel, 1e-10))
        ls = np.maximum(ls, ls.max() - 8.0)
        out[b] = (ls + 4.0) / 4.0
    return out.reshape(*lead, n_mels, n_frames)


def peak_normalize(audio: np.ndarray) -> np.ndarray:
    """preprocess/audio_process.py:312-317 (and :291-293): float32 cast; if any sample lies
    outside [-1, 1] divide by max(|max|, |min|)."""
    w = np.asarray(audio).astype(np.float32)
    if w.size and (w.max() > 1.0 or w.min() < -1.0):
        w = w / max(abs(w.max()), abs(w.min()))
    return w
