"""Oracle: AV-HuBERT audio features (CPU, numpy).  TEST INFRASTRUCTURE ONLY.

Restates ``extract_logfbank_features`` and ``audio_to_tensor`` of the reference
(preprocess/audio_process.py:152-197; identical code in utils/data_loading.py:181-201).  The
arithmetic of ``logfbank`` lives in the third-party package ``python_speech_features`` (imported at
preprocess/audio_process.py:8, absent from ``requirements.txt`` and from this image): PARITY
UNPINNED.  Restated from its published source (version 0.6, ``base.py`` / ``sigproc.py``):

* ``preemphasis(signal, 0.97)``: ``append(signal[0], signal[1:] - 0.97 * signal[:-1])`` (float32 in,
  float32 out);
* ``framesig``: 400-sample frames every 160, ``1 + ceil((len - 400) / 160)`` of them (1 if
  ``len <= 400``), the tail zero-padded (the concatenation with float64 zeros makes everything after
  it float64), window of ones;
* ``powspec``: ``1/512 * |rfft(frames, 512)|^2``;
* ``get_filterbanks(26, 512, 16000, 0, 8000)``: HTK mel scale, triangular filters between
  ``floor((nfft + 1) * mel2hz(melpoints) / samplerate)`` bins;
* ``fbank``: ``dot(pspec, fb.T)``, zeros replaced by ``finfo(float).eps``; ``logfbank``: ``log``.
"""
from __future__ import annotations

import math

import numpy as np

NFFT, FRAME_LEN, FRAME_STEP, NFILT, PREEMPH = 512, 400, 160, 26, 0.97


def hz2mel(hz):
    return 2595 * np.log10(1 + hz / 700.0)


def mel2hz(mel):
    return 700 * (10 ** (mel / 2595.0) - 1)


def get_filterbanks(nfilt: int = NFILT, nfft: int = NFFT, samplerate: int = 16000, lowfreq: float = 0,
                    highfreq=None) -> np.ndarray:
    highfreq = highfreq or samplerate / 2
    lowmel, highmel = hz2mel(lowfreq), hz2mel(highfreq)
    melpoints = np.linspace(lowmel, highmel, nfilt + 2)
    bins = np.floor((nfft + 1) * mel2hz(melpoints) / samplerate)
    fbank = np.zeros([nfilt, nfft // 2 + 1])
    for j in range(0, nfilt):
        for i in range(int(bins[j]), int(bins[j + 1])):
            fbank[j, i] = (i - bins[j]) / (bins[j + 1] - bins[j])
        for i in range(int(bins[j + 1]), int(bins[j + 2])):
            fbank[j, i] = (bins[j + 2] - i) / (bins[j + 2] - bins[j + 1])
    return fbank


def num_frames(slen: int) -> int:
    if slen <= FRAME_LEN:
        return 1
    return 1 + int(math.ceil((1.0 * slen - FRAME_LEN) / FRAME_STEP))


def logfbank(signal: np.ndarray, samplerate: int = 16000) -> np.ndarray:
    """python_speech_features.logfbank with its defaults.  float64 [num_frames, 26]."""
    signal = np.asarray(signal)
    signal = np.append(signal[0], signal[1:] - PREEMPH * signal[:-1])
    slen = len(signal)
    nfr = num_frames(slen)
    padlen = (nfr - 1) * FRAME_STEP + FRAME_LEN
    padsignal = np.concatenate((signal, np.zeros((padlen - slen,))))
    idx = np.arange(FRAME_LEN)[None, :] + FRAME_STEP * np.arange(nfr)[:, None]
    frames = padsignal[idx] * np.ones((FRAME_LEN,))
    pspec = 1.0 / NFFT * np.square(np.absolute(np.fft.rfft(frames, NFFT)))
    feat = np.dot(pspec, get_filterbanks(NFILT, NFFT, samplerate).T)
    feat = np.where(feat == 0, np.finfo(float).eps, feat)
    return np.log(feat)


def extract_logfbank_features(audio_data, sample_rate: int = 16000, stack_order: int = 1) -> np.ndarray:
    """Semantics of preprocess/audio_process.py:152-179: float32 log filterbank frames; with
    ``stack_order`` > 1, zero frames are appended up to a multiple of it and every ``stack_order``
    consecutive frames become one row."""
    frames = logfbank(audio_data, samplerate=sample_rate).astype(np.float32)
    if stack_order <= 1:
        return frames
    n, dim = frames.shape
    short = (-n) % stack_order
    if short:
        frames = np.pad(frames, ((0, short), (0, 0)))
    return frames.reshape(-1, stack_order * dim)


def audio_to_tensor(audio_features: np.ndarray, normalize: bool = True) -> np.ndarray:
    """Semantics of preprocess/audio_process.py:181-197: per row, subtract the mean and divide by
    (population standard deviation + 1e-5)."""
    if not normalize:
        return audio_features
    mu = audio_features.mean(axis=1, keepdims=True)
    sd = audio_features.std(axis=1, keepdims=True)
    return (audio_features - mu) / (sd + 1e-5)
