"""Audio front-end: host-side mirror of the reference's Whisper audio call sites.

Same names and argument meaning as the functions the reference calls
(``avsl/whisper_flamingo_ft_ami.py:209-213``: ``whisper.pad_or_trim`` and
``whisper.log_mel_spectrogram(audio, n_mels, padding)``; ``preprocess/audio_process.py:301-319``
for the waveform normalisation).  All arithmetic runs in libavfe.so on the GPU; numpy / CPU
inputs are staged through pinned memory and results come back in the container type the
reference would have returned.
"""
from __future__ import annotations

import ctypes
import functools
import math
from typing import Optional, Union

import numpy as np
import torch

from . import _lib

SAMPLE_RATE = 16000
N_FFT = 400
HOP_LENGTH = 160
N_FREQ = 201
CHUNK_LENGTH = 30
N_SAMPLES = CHUNK_LENGTH * SAMPLE_RATE   # 480000
N_FRAMES = N_SAMPLES // HOP_LENGTH       # 3000


@functools.lru_cache(maxsize=None)
def _mel_filters_np(n_mels: int) -> np.ndarray:
    """Slaney-scale, slaney-normalised triangular filterbank [n_mels, 201] float32 — the matrix
    openai-whisper ships as assets/mel_filters.npz (librosa.filters.mel(sr=16000, n_fft=400))."""
    f_sp, min_log_hz = 200.0 / 3.0, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, math.log(6.4) / 27.0

    def hz_to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, f / f_sp)

    def mel_to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)

    fft_freqs = np.linspace(0.0, SAMPLE_RATE / 2.0, N_FREQ)
    mel_pts = mel_to_hz(np.linspace(hz_to_mel(0.0), hz_to_mel(SAMPLE_RATE / 2.0), n_mels + 2))
    fdiff = np.diff(mel_pts)
    ramps = mel_pts[:, None] - fft_freqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    fb = np.maximum(0.0, np.minimum(lower, upper))
    fb *= (2.0 / (mel_pts[2:] - mel_pts[:-2]))[:, None]
    return np.ascontiguousarray(fb.astype(np.float32))


_FILTER_CACHE: dict = {}
_PACK_CACHE: dict = {}


def _filter_pack(fb: torch.Tensor, n_mels: int) -> torch.Tensor:
    """Sparse form of a filterbank tensor (avfe_logmel_prepare), built once per tensor."""
    key = (fb.data_ptr(), str(fb.device), n_mels, fb._version)
    hit = _PACK_CACHE.get(key)
    if hit is not None and hit[0] is fb:
        return hit[1]
    lib = _lib.load()
    pack = torch.empty(int(lib.avfe_logmel_pack_bytes()), dtype=torch.uint8, device=fb.device)
    with torch.cuda.device(fb.device):
        _lib.call("avfe_logmel_prepare", _lib.ptr(fb), n_mels, _lib.ptr(pack), _lib.stream_ptr())
    if len(_PACK_CACHE) > 16:
        _PACK_CACHE.clear()
    _PACK_CACHE[key] = (fb, pack)      # holding fb keeps data_ptr from being recycled
    return pack


def mel_filters(device, n_mels: int = 80) -> torch.Tensor:
    """``whisper.audio.mel_filters(device, n_mels)``: [n_mels, 201] float32 on ``device``."""
    if n_mels <= 0 or n_mels > 128:
        raise ValueError(f"Unsupported n_mels: {n_mels}")
    key = (str(torch.device(device)), n_mels)
    if key not in _FILTER_CACHE:
        _FILTER_CACHE[key] = torch.from_numpy(_mel_filters_np(n_mels)).to(device)
    return _FILTER_CACHE[key]


def _to_cuda_f32(x, device=None) -> tuple:
    """-> (contiguous float32 CUDA tensor, kind) with kind in {'numpy','cpu','cuda'}."""
    _lib.require_cuda()
    if isinstance(x, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        kind = "numpy"
    elif torch.is_tensor(x):
        t = x
        kind = "cuda" if x.is_cuda else "cpu"
    else:
        t = torch.as_tensor(np.asarray(x, dtype=np.float32))
        kind = "numpy"
    if not t.is_cuda:
        t = t.to(torch.float32).contiguous()
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        t = t.pin_memory().to(dev, non_blocking=True)
    else:
        t = t.to(torch.float32).contiguous()
    return t, kind


def pad_or_trim(array, length: int = N_SAMPLES, *, axis: int = -1):
    """``whisper.pad_or_trim``: zero-pad or cut ``axis`` to ``length``; numpy in -> numpy out,
    tensor in -> tensor out (same device)."""
    if axis not in (-1, (array.ndim - 1)):
        raise NotImplementedError("pad_or_trim is implemented for the last axis only")
    t, kind = _to_cuda_f32(array)
    lead = t.shape[:-1]
    L_in = t.shape[-1]
    B = int(np.prod(lead)) if lead else 1
    out = torch.empty((*lead, length), dtype=torch.float32, device=t.device)
    with torch.cuda.device(t.device):
        _lib.call("avfe_pad_or_trim_f32", _lib.ptr(t), B, L_in, length, _lib.ptr(out), _lib.stream_ptr())
    if kind == "numpy":
        return out.cpu().numpy()
    return out.cpu() if kind == "cpu" else out


def log_mel_spectrogram(audio: Union[np.ndarray, torch.Tensor], n_mels: int = 80,
                        padding: int = 0, device: Optional[Union[str, torch.device]] = None,
                        filters: Optional[torch.Tensor] = None,
                        out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Drop-in for ``whisper.log_mel_spectrogram(audio, n_mels, padding, device)``.

    audio: [..., L] float waveform at 16 kHz (numpy array, CPU tensor or CUDA tensor).
    Returns float32 ``[..., n_mels, (L + padding) // 160]``.  CUDA input (or an explicit
    ``device``) keeps the result on the GPU; numpy / CPU input gets a CPU tensor back, as the
    reference does.  The max-8 dB floor is applied per clip (the reference calls this once per
    sample).  ``out`` (CUDA, contiguous) lets a caller write straight into a batch buffer.
    """
    if n_mels <= 0 or n_mels > 128:
        raise ValueError(f"Unsupported n_mels: {n_mels}")
    t, kind = _to_cuda_f32(audio, device)
    if t.dim() == 0:
        raise ValueError("audio must have at least one dimension")
    lead = tuple(t.shape[:-1])
    L = int(t.shape[-1])
    B = int(np.prod(lead)) if lead else 1
    n_frames = (L + padding) // HOP_LENGTH
    if L + padding <= N_FFT // 2:
        # torch.stft: "Padding size should be less than the corresponding input dimension"
        raise RuntimeError("log_mel_spectrogram: reflect padding (200) needs more than 200 samples")
    fb = filters if filters is not None else mel_filters(t.device, n_mels)
    if not (fb.is_cuda and fb.dtype == torch.float32 and fb.is_contiguous() and tuple(fb.shape) == (n_mels, N_FREQ)):
        fb = fb.to(t.device, torch.float32).contiguous()
        if tuple(fb.shape) != (n_mels, N_FREQ):
            raise ValueError("filters must be [n_mels, 201]")
    shape = (*lead, n_mels, n_frames)
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=t.device)
    elif not (out.is_cuda and out.is_contiguous() and out.dtype == torch.float32 and tuple(out.shape) == shape):
        raise ValueError(f"out must be a contiguous float32 CUDA tensor of shape {shape}")
    with torch.cuda.device(t.device):
        lib = _lib.load()
        ws_bytes = int(lib.avfe_logmel_workspace_bytes(B, L, padding, n_mels))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=t.device)
        pack = _filter_pack(fb, n_mels)
        _lib.call("avfe_logmel_prepared_f32", _lib.ptr(t), B, L, padding, n_mels, _lib.ptr(fb), _lib.ptr(pack),
                  _lib.ptr(out), _lib.ptr(ws), ws_bytes, _lib.stream_ptr())
    if kind in ("numpy", "cpu") and device is None:
        return out.cpu()
    return out


def log_mel_spectrogram_ragged(audio: torch.Tensor, offsets: torch.Tensor, length: int = N_SAMPLES,
                               n_mels: int = 80, filters: Optional[torch.Tensor] = None,
                               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``pad_or_trim(clip, length)`` + ``log_mel_spectrogram(clip, n_mels)`` for a batch of clips
    stored back to back on the GPU (``audio`` float32 [sum L_i], ``offsets`` int64 [B+1]), in one
    launch sequence: the zero padding is never written or read.  Returns [B, n_mels, length//160]."""
    _lib.require_cuda()
    if not (audio.is_cuda and offsets.is_cuda and audio.dtype == torch.float32 and offsets.dtype == torch.int64):
        raise ValueError("audio (float32) and offsets (int64) must be CUDA tensors")
    if n_mels <= 0 or n_mels > 128:
        raise ValueError(f"Unsupported n_mels: {n_mels}")
    if length <= N_FFT // 2:
        raise RuntimeError("log_mel_spectrogram: reflect padding (200) needs more than 200 samples")
    audio = audio.contiguous()
    B = int(offsets.numel()) - 1
    fb = filters if filters is not None else mel_filters(audio.device, n_mels)
    shape = (B, n_mels, length // HOP_LENGTH)
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=audio.device)
    elif not (out.is_cuda and out.is_contiguous() and out.dtype == torch.float32 and tuple(out.shape) == shape):
        raise ValueError(f"out must be a contiguous float32 CUDA tensor of shape {shape}")
    with torch.cuda.device(audio.device):
        lib = _lib.load()
        ws_bytes = int(lib.avfe_logmel_workspace_bytes(B, length, 0, n_mels))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=audio.device)
        pack = _filter_pack(fb, n_mels)
        _lib.call("avfe_logmel_ragged_f32", _lib.ptr(audio), _lib.ptr(offsets), B, length, n_mels, _lib.ptr(fb),
                  _lib.ptr(pack), _lib.ptr(out), _lib.ptr(ws), ws_bytes, _lib.stream_ptr())
    return out


def peak_normalize(audio: Union[np.ndarray, torch.Tensor]):
    """The waveform conditioning of ``preprocess_audio_for_whisper`` /
    ``process_audio_dual_encoder['waveform']`` (preprocess/audio_process.py:291-293,312-317)
    without the file load: float32 cast; clips with a sample outside [-1, 1] are divided by
    max(|max|, |min|).  [..., L] -> same shape, same container type."""
    t, kind = _to_cuda_f32(audio)
    lead = tuple(t.shape[:-1])
    L = int(t.shape[-1])
    B = int(np.prod(lead)) if lead else 1
    out = torch.empty_like(t)
    scratch = torch.empty(max(2 * B, 2), dtype=torch.float32, device=t.device)
    with torch.cuda.device(t.device):
        _lib.call("avfe_peak_normalize_f32", _lib.ptr(t), B, L, _lib.ptr(out), _lib.ptr(scratch),
                  _lib.stream_ptr())
    if kind == "numpy":
        return out.cpu().numpy()
    return out.cpu() if kind == "cpu" else out


# ----------------------------------------------------------------------------- AV-HuBERT audio features
@functools.lru_cache(maxsize=None)
def _logfbank_filters_np(nfilt: int = 26, nfft: int = 512, samplerate: int = 16000) -> np.ndarray:
    """python_speech_features.get_filterbanks(nfilt, nfft, samplerate, 0, samplerate/2): HTK mel
    scale, triangular filters between floor((nfft + 1) * hz / samplerate) bins.  float32 [nfilt, 257]."""
    def hz2mel(hz):
        return 2595 * np.log10(1 + hz / 700.0)

    def mel2hz(mel):
        return 700 * (10 ** (mel / 2595.0) - 1)

    melpoints = np.linspace(hz2mel(0.0), hz2mel(samplerate / 2), nfilt + 2)
    bins = np.floor((nfft + 1) * mel2hz(melpoints) / samplerate)
    fb = np.zeros([nfilt, nfft // 2 + 1])
    for j in range(nfilt):
        for i in range(int(bins[j]), int(bins[j + 1])):
            fb[j, i] = (i - bins[j]) / (bins[j + 1] - bins[j])
        for i in range(int(bins[j + 1]), int(bins[j + 2])):
            fb[j, i] = (bins[j + 2] - i) / (bins[j + 2] - bins[j + 1])
    return np.ascontiguousarray(fb.astype(np.float32))


def logfbank_num_frames(n_samples: int) -> int:
    """Frames python_speech_features makes of ``n_samples`` (25 ms / 10 ms at 16 kHz)."""
    return int(_lib.load().avfe_logfbank_num_frames(int(n_samples)))


class LogfbankPlan:
    """Device-side bookkeeping of one packed batch layout (clip boundaries, output row boundaries,
    workspace, output buffer), reusable for every batch with the same clip lengths."""

    def __init__(self, offsets, stack_order: int, nfilt: int, device):
        off = np.asarray(offsets.cpu() if torch.is_tensor(offsets) else offsets, dtype=np.int64)
        lens = np.diff(off)
        if len(off) < 2 or (lens < 1).any():
            raise ValueError("every clip needs at least one sample")
        frames = np.where(lens <= 400, 1, 1 + -(-(lens - 400) // 160))
        rows = -(-frames // stack_order)
        self.offsets, self.row_offsets = off, np.concatenate([[0], np.cumsum(rows)]).astype(np.int64)
        self.B, self.max_len = len(lens), int(lens.max())
        self.stack_order, self.nfilt, self.device = int(stack_order), int(nfilt), device
        self.d_off = torch.from_numpy(self.offsets).to(device)
        self.d_row = torch.from_numpy(self.row_offsets).to(device)
        self.ws = torch.empty(int(_lib.load().avfe_logfbank_workspace_bytes()), dtype=torch.uint8, device=device)
        self.out = torch.empty((int(self.row_offsets[-1]), nfilt * stack_order), dtype=torch.float32, device=device)


def logfbank_batch(audio: torch.Tensor, offsets=None, stack_order: int = 4, normalize: bool = True,
                   nfilt: int = 26, plan: Optional[LogfbankPlan] = None):
    """AV-HuBERT audio features of a packed batch on the GPU.  ``audio`` float32 CUDA [sum L_i],
    ``offsets`` the [B+1] clip boundaries (host sequence or tensor), or a ``plan`` built once for
    that layout (steady-state loops: no host-to-device traffic, output buffer reused).  Returns
    ``(feats, row_offsets)``: float32 CUDA ``[sum rows_i, nfilt * stack_order]`` and the int64 numpy
    row boundaries (``rows_i = ceil(frames_i / stack_order)``)."""
    _lib.require_cuda()
    if not (audio.is_cuda and audio.dtype == torch.float32 and audio.is_contiguous() and audio.dim() == 1):
        raise ValueError("audio must be a contiguous 1-D float32 CUDA tensor")
    dev = audio.device
    if plan is None:
        plan = LogfbankPlan(offsets, stack_order, nfilt, dev)
    if int(plan.offsets[-1]) != audio.numel():
        raise ValueError("offsets do not cover the audio tensor")
    key = ("fbank", str(dev), plan.nfilt)
    hit = _FILTER_CACHE.get(key)
    with torch.cuda.device(dev):
        if hit is None:
            # the filterbank and its sparse form (avfe_logfbank_prepare), once per device and nfilt
            fb = torch.from_numpy(_logfbank_filters_np(plan.nfilt)).to(dev)
            pack = torch.empty(int(_lib.load().avfe_logfbank_workspace_bytes()), dtype=torch.uint8, device=dev)
            _lib.call("avfe_logfbank_prepare", _lib.ptr(fb), plan.nfilt, _lib.ptr(pack), pack.numel(), _lib.stream_ptr())
            hit = _FILTER_CACHE[key] = (fb, pack)
        fb, pack = hit
        _lib.call("avfe_logfbank_prepared_f32", _lib.ptr(audio), _lib.ptr(plan.d_off), _lib.ptr(plan.d_row), plan.B,
                  plan.max_len, _lib.ptr(fb), plan.nfilt, plan.stack_order, 1 if normalize else 0,
                  _lib.ptr(plan.out), _lib.ptr(pack), pack.numel(), _lib.stream_ptr())
    return plan.out, plan.row_offsets


def extract_logfbank_features(audio_data, sample_rate: int = 16000, stack_order: int = 1,
                              normalize: bool = False) -> np.ndarray:
    """``extract_logfbank_features`` (preprocess/audio_process.py:152-179) and, with
    ``normalize=True``, ``audio_to_tensor`` (:181-197) on top of it, for one clip given as a host
    array; float32 numpy ``[ceil(frames / stack_order), 26 * stack_order]`` like the reference."""
    if sample_rate != SAMPLE_RATE:
        raise ValueError("extract_logfbank_features is built for 16 kHz audio")
    a = torch.from_numpy(np.ascontiguousarray(np.asarray(audio_data, dtype=np.float32).reshape(-1)))
    _lib.require_cuda()
    feats, _ = logfbank_batch(a.cuda(), [0, a.numel()], stack_order, normalize)
    return feats.cpu().numpy()


# ----------------------------------------------------------------------------- SNR noise mixing
def snr_ratio(snr) -> np.float32:
    """``float32(10 ** (snr / 20))``: the divisor of preprocess/audio_process.py:135 as numpy 2
    evaluates it next to a float32 scalar (the Python float takes the scalar's type).  ``snr`` is
    taken as a Python number, which is what the reference's call site passes (``noise_snr=0``,
    :200); a numpy float64 scalar would make numpy carry the whole mix in float64 -- a different
    (unpinned) result that this library does not reproduce."""
    return np.float32(10 ** (float(snr) / 20))


class NoisePlan:
    """Device-side bookkeeping of one packed clean / noise layout (clip boundaries, per-clip SNR
    divisors, workspace, output buffer), reusable for every batch with the same clip lengths."""

    def __init__(self, clean_offsets, noise_offsets, snr, device, out_dtype=torch.int16):
        if out_dtype not in (torch.int16, torch.float32):
            raise ValueError("out_dtype must be torch.int16 or torch.float32")
        co = np.asarray(clean_offsets, dtype=np.int64)
        no = np.asarray(noise_offsets, dtype=np.int64)
        B = len(co) - 1
        if B < 0 or len(no) != B + 1 or (np.diff(co) < 0).any() or (np.diff(no) < 0).any():
            raise ValueError("clean_offsets / noise_offsets must be non-decreasing and of equal length")
        if ((np.diff(co) > 0) & (np.diff(no) == 0)).any():
            raise ZeroDivisionError("a clean clip has an empty noise clip")     # the reference's ceil(Lc / 0), :128
        ratios = np.broadcast_to(np.asarray([snr_ratio(v) for v in np.atleast_1d(snr)], dtype=np.float32), (B,))
        self.co, self.no, self.B, self.out_dtype, self.device = co, no, B, out_dtype, device
        self.max_len = int(np.diff(co).max()) if B else 0
        self.d_co, self.d_no = torch.from_numpy(co).to(device), torch.from_numpy(no).to(device)
        self.d_ratio = torch.from_numpy(np.array(ratios, dtype=np.float32)).to(device)     # broadcast_to is read-only: copy
        nbytes = int(_lib.load().avfe_add_noise_workspace_bytes(B, self.max_len))
        if nbytes == 0:
            raise ValueError("clip too long for avfe_add_noise")
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.out = None


def add_noise_batch(clean: torch.Tensor, clean_offsets=None, noise: torch.Tensor = None, noise_offsets=None,
                    snr=0, out_dtype=torch.int16, plan: Optional[NoisePlan] = None) -> torch.Tensor:
    """``add_noise`` (preprocess/audio_process.py:110-150) for a packed batch on the GPU, bit-exact
    with the reference's float32 numpy arithmetic.  ``clean`` / ``noise``: contiguous 1-D float32 CUDA
    tensors holding the clips back to back, ``*_offsets`` their [B+1] boundaries (host sequences),
    ``snr`` a number or one per clip (dB) -- or a ``plan`` built once for that layout (steady-state
    loops: no host-to-device traffic, output buffer reused).  A noise clip is repeated or cut to its
    clean clip's length like the reference does.  Returns the mixed clips packed like ``clean``:
    int16 (the reference's return type) or, with ``out_dtype=torch.float32``, the same integers as
    float32 -- what ``logfbank_batch`` takes next (:224-227)."""
    _lib.require_cuda()
    for name, t in (("clean", clean), ("noise", noise)):
        if not (torch.is_tensor(t) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.dim() == 1):
            raise ValueError(f"{name} must be a contiguous 1-D float32 CUDA tensor")
    if clean.data_ptr() % 16:
        raise ValueError("clean must start on a 16-byte boundary (pass the packed buffer, not a slice of it)")
    dev = clean.device
    if plan is None:
        plan = NoisePlan(clean_offsets, noise_offsets, snr, dev, out_dtype)
    co, no = plan.co, plan.no
    if plan.B and (int(co[-1]) > clean.numel() or int(no[-1]) > noise.numel()):
        raise ValueError("offsets run past the end of the tensors")
    if plan.out is None or plan.out.numel() != clean.numel():
        plan.out = torch.empty(clean.numel(), dtype=plan.out_dtype, device=dev)
    out = plan.out
    if plan.B == 0 or clean.numel() == 0:
        return out
    with torch.cuda.device(dev):
        _lib.call("avfe_add_noise", _lib.ptr(clean), _lib.ptr(plan.d_co), _lib.ptr(noise), _lib.ptr(plan.d_no),
                  _lib.ptr(plan.d_ratio), plan.B, plan.max_len, _lib.ptr(out) if plan.out_dtype == torch.int16 else None,
                  _lib.ptr(out) if plan.out_dtype == torch.float32 else None, _lib.ptr(plan.ws), plan.ws.numel(),
                  _lib.stream_ptr())
    if int(co[0]) > 0 or int(co[-1]) < clean.numel():      # samples outside every clip pass through unmixed
        out[:int(co[0])] = clean[:int(co[0])].to(plan.out_dtype)
        out[int(co[-1]):] = clean[int(co[-1]):].to(plan.out_dtype)
    return out


def add_noise(clean_wav, noise_wav, snr) -> np.ndarray:
    """``add_noise(clean_wav, noise_wav, snr)`` of preprocess/audio_process.py:110-150 for one clip
    given as host arrays: int16 numpy like the reference."""
    c = torch.from_numpy(np.ascontiguousarray(np.asarray(clean_wav).astype(np.float32).reshape(-1)))
    z = torch.from_numpy(np.ascontiguousarray(np.asarray(noise_wav).astype(np.float32).reshape(-1)))
    _lib.require_cuda()
    return add_noise_batch(c.cuda(), [0, c.numel()], z.cuda(), [0, z.numel()], snr).cpu().numpy()


def process_audio_for_av_hubert(audio, stack_order: int = 1, normalize: bool = True, add_noise_prob: float = 0.0,
                                noise=None, noise_snr=0, rng=None):
    """``process_audio_for_av_hubert`` (preprocess/audio_process.py:199-236) from the point where the
    waveforms are in memory (the reference's ``librosa.load`` of ``audio_path`` / ``noise_file`` is the
    caller's; pass the 16 kHz arrays): optional SNR noise mixing with probability ``add_noise_prob``
    (:222-224, drawn like the reference from ``np.random.rand()`` unless ``rng`` is given), logfbank
    features stacked ``stack_order`` at a time (:227), per-row normalisation (:230).  Everything after
    the draw runs on the GPU without a host round trip.  float32 numpy
    ``[ceil(frames / stack_order), 26 * stack_order]``; like the reference, any failure is reported
    and answered with ``None`` (:234-236)."""
    _lib.require_cuda()
    try:
        a = torch.from_numpy(np.ascontiguousarray(np.asarray(audio).astype(np.float32).reshape(-1))).cuda()
        draw = (rng.random() if rng is not None else np.random.rand()) if add_noise_prob > 0 and noise is not None else 1.0
        if add_noise_prob > 0 and noise is not None and draw < add_noise_prob:
            z = torch.from_numpy(np.ascontiguousarray(np.asarray(noise).astype(np.float32).reshape(-1))).cuda()
            a = add_noise_batch(a, [0, a.numel()], z, [0, z.numel()], noise_snr, out_dtype=torch.float32)
        feats, _ = logfbank_batch(a, [0, a.numel()], stack_order, normalize)
        return feats.cpu().numpy()
    except Exception as e:                                  # the reference's catch-all, :234-236
        print(f"Error processing audio for AV-HuBERT: {str(e)}")
        return None


# ----------------------------------------------------------------------------- SpecAugment masks
# LibriSpeech policies of the SpecAugment paper (Park et al. 2019, table 1), the names the
# reference passes as ``spec_augment_config`` (avsl/whisper_flamingo_ft_ami.py:165,217-220):
# frequency mask parameter F, number of frequency masks, time mask parameter T, upper bound p on
# the masked fraction of the utterance, number of time masks.  Time warping (W = 80) is not applied.
SPEC_AUGMENT_POLICIES = {"ls-basic": dict(F=27, n_freq_mask=1, T=100, p=1.0, n_time_mask=1),
                         "ls-double": dict(F=27, n_freq_mask=2, T=100, p=1.0, n_time_mask=2)}


def spec_augment_bands(audio_frames, n_mels: int = 80, policy: str = "ls-double", rng=None,
                       n_freq_mask: Optional[int] = None, n_time_mask: Optional[int] = None) -> np.ndarray:
    """Draw the mask rectangles for a batch on the host: int32 ``[B, n_bands, 4]`` rows
    ``(f0, f1, t0, t1)``.  ``audio_frames[b]`` is the clip's frame count before padding
    (``audio_frames_before_pad``, whisper_flamingo_ft_ami.py:206): time masks stay inside it.
    Per clip, in this order: for every frequency mask ``f ~ U{0..F}``, ``f0 ~ U{0..n_mels-f}``; then for
    every time mask ``t ~ U{0..min(T, floor(p*tau))}``, ``t0 ~ U{0..tau-t}`` (the paper's sampling)."""
    if policy not in SPEC_AUGMENT_POLICIES:
        raise NotImplementedError(policy)                      # as the reference (:221-222)
    cfg = dict(SPEC_AUGMENT_POLICIES[policy])
    if n_freq_mask is not None:
        cfg["n_freq_mask"] = n_freq_mask
    if n_time_mask is not None:
        cfg["n_time_mask"] = n_time_mask
    rng = rng if rng is not None else np.random.default_rng()
    frames = np.atleast_1d(np.asarray(audio_frames, dtype=np.int64))
    nb = cfg["n_freq_mask"] + cfg["n_time_mask"]
    bands = np.zeros((len(frames), nb, 4), dtype=np.int32)
    for b, tau in enumerate(frames):
        k = 0
        for _ in range(cfg["n_freq_mask"]):
            f = int(rng.integers(0, min(cfg["F"], n_mels) + 1))
            f0 = int(rng.integers(0, n_mels - f + 1))
            bands[b, k] = (f0, f0 + f, 0, np.iinfo(np.int32).max)
            k += 1
        for _ in range(cfg["n_time_mask"]):
            t_max = int(min(cfg["T"], np.floor(cfg["p"] * tau)))
            t = int(rng.integers(0, max(t_max, 0) + 1))
            t0 = int(rng.integers(0, max(int(tau) - t, 0) + 1))
            bands[b, k] = (0, n_mels, t0, t0 + t)
            k += 1
    return bands


def spec_augment_warp_points(audio_frames, W: int = 80, rng=None) -> np.ndarray:
    """Draw SpecAugment's time-warp points for a batch on the host: int32 ``[B, 3]`` rows
    ``(tau, center, warped)`` with ``center ~ U{W..tau-W-1}`` and ``warped = center + w``,
    ``w ~ U{-W..W}`` (the paper's W = 80 for the LibriSpeech policies); clips with ``tau <= 2 W`` are
    left alone (``center = 0``).  Drawn per clip BEFORE the clip's masks when combined with
    :func:`spec_augment_bands` on one generator (the paper's order: warp, frequency masks, time masks)."""
    rng = rng if rng is not None else np.random.default_rng()
    frames = np.atleast_1d(np.asarray(audio_frames, dtype=np.int64))
    pts = np.zeros((len(frames), 3), dtype=np.int32)
    for b, tau in enumerate(frames):
        pts[b, 0] = tau
        if tau - W > W:
            c = int(rng.integers(W, tau - W))
            pts[b, 1] = c
            pts[b, 2] = c + int(rng.integers(-W, W + 1))
    return pts


def spec_time_warp(mel: torch.Tensor, warp_points: np.ndarray, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """SpecAugment time warping of ``mel`` float32 CUDA ``[B, n_mels, n_frames]`` with given warp points
    (``spec_augment_warp_points``); returns a new tensor (the warp cannot run in place)."""
    _lib.require_cuda()
    if not (mel.is_cuda and mel.dtype == torch.float32 and mel.is_contiguous() and mel.dim() == 3):
        raise ValueError("mel must be a contiguous float32 CUDA tensor [B, n_mels, n_frames]")
    B, n_mels, n_frames = (int(s) for s in mel.shape)
    pts = np.ascontiguousarray(warp_points, dtype=np.int32)
    if pts.shape != (B, 3):
        raise ValueError("warp_points must be [B, 3]")
    if out is None:
        out = torch.empty_like(mel)
    elif out.data_ptr() == mel.data_ptr() or out.shape != mel.shape or out.dtype != mel.dtype or not out.is_contiguous():
        raise ValueError("out must be a separate contiguous tensor of mel's shape")
    d_pts = torch.from_numpy(pts).to(mel.device)
    with torch.cuda.device(mel.device):
        _lib.call("avfe_spec_time_warp_f32", _lib.ptr(mel), B, n_mels, n_frames, _lib.ptr(d_pts), _lib.ptr(out),
                  _lib.stream_ptr())
    return out


def align_audio_video_features(audio_features, video_features):
    """``align_audio_video_features`` (preprocess/audio_process.py:238-264): truncate the longer of the
    two feature sequences to the length of the shorter (first axis); ``None`` passes through.  Works on
    numpy arrays and (CUDA) tensors alike and returns views -- nothing is copied."""
    if audio_features is None or video_features is None:
        return audio_features, video_features
    audio_len, video_len = len(audio_features), len(video_features)
    if audio_len > video_len:
        audio_features = audio_features[:video_len]
    elif audio_len < video_len:
        video_features = video_features[:audio_len]
    return audio_features, video_features


def aligned_lengths(audio_rows, video_frames) -> np.ndarray:
    """Per-clip common length ``min(audio_rows[i], video_frames[i])`` for a packed batch: the
    ``keep_frames`` of :func:`avsl_b200.lips.lip_roi_collate` and the row count to keep of each clip's
    logfbank features, i.e. :func:`align_audio_video_features` for every clip at once."""
    return np.minimum(np.asarray(audio_rows, dtype=np.int64), np.asarray(video_frames, dtype=np.int64))


def process_audio_dual_encoder(audio, stack_order: int = 1, normalize: bool = True, sample_rate: int = SAMPLE_RATE):
    """``process_audio_dual_encoder`` (preprocess/audio_process.py:267-299) from the loaded 16 kHz
    waveform on (the reference's ``librosa.load`` is the caller's): one upload, two kernels --
    logfbank features for AV-HuBERT (stacked, normalised; :284-285) and the peak-normalised float32
    waveform for Whisper (:289-293).  numpy in -> numpy out; a CUDA tensor in -> CUDA tensors out."""
    _lib.require_cuda()
    is_dev = torch.is_tensor(audio) and audio.is_cuda
    a = audio if is_dev else torch.from_numpy(np.ascontiguousarray(np.asarray(audio).astype(np.float32).reshape(-1))).cuda()
    a = a.to(torch.float32).reshape(-1).contiguous()
    feats, _ = logfbank_batch(a, [0, a.numel()], stack_order, normalize)
    wave = peak_normalize(a)
    if not is_dev:
        feats, wave = feats.cpu().numpy(), wave.cpu().numpy()
    return {"waveform": wave, "sample_rate": sample_rate, "av_hubert_features": feats}


def spec_augment(mel: torch.Tensor, audio_frames=None, policy: str = "ls-double", rng=None,
                 bands: Optional[np.ndarray] = None, fill: float = 0.0) -> torch.Tensor:
    """Apply SpecAugment masks in place to ``mel`` float32 CUDA ``[B, n_mels, n_frames]`` (or
    ``[n_mels, n_frames]``).  Either pass ``bands`` (from :func:`spec_augment_bands`) or
    ``audio_frames`` + ``policy`` to draw them here."""
    _lib.require_cuda()
    if not (mel.is_cuda and mel.dtype == torch.float32 and mel.is_contiguous() and mel.dim() in (2, 3)):
        raise ValueError("mel must be a contiguous float32 CUDA tensor [B, n_mels, n_frames]")
    m3 = mel if mel.dim() == 3 else mel.unsqueeze(0)
    B, n_mels, n_frames = (int(s) for s in m3.shape)
    if bands is None:
        if audio_frames is None:
            audio_frames = [n_frames] * B
        bands = spec_augment_bands(audio_frames, n_mels, policy, rng)
    bands = np.ascontiguousarray(bands, dtype=np.int32)
    if bands.shape[0] != B or bands.shape[-1] != 4:
        raise ValueError("bands must be [B, n_bands, 4]")
    d_bands = torch.from_numpy(bands).to(mel.device)
    with torch.cuda.device(mel.device):
        _lib.call("avfe_spec_mask_f32", _lib.ptr(m3), B, n_mels, n_frames, _lib.ptr(d_bands), int(bands.shape[1]),
                  float(fill), _lib.stream_ptr())
    return mel
