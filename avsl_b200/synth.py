"""Seeded synthetic inputs for the BASELINE.json configurations.

There is no dataset access, so tests and bench.py run on these generators: 16 kHz waveforms,
224x224 BGR "closeup" clips with 68-point landmarks derived from the mean face, AV-HuBERT-sized
feature maps, and an AMI-shaped duration distribution.  SEED = 3407 is the reference's own
seed (avsl/whisper_flamingo_ft_ami.py:149).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch

from .lips import mean_face_landmarks

SEED = 3407
SAMPLE_RATE = 16000
FPS = 25


# ----------------------------------------------------------------------------- audio
def audio_clip(n_samples: int = 480000, seed: int = SEED) -> np.ndarray:
    """Config 1: clamp(0.1 * randn, -1, 1), float32."""
    g = torch.Generator().manual_seed(seed)
    return torch.clamp(0.1 * torch.randn(n_samples, generator=g), -1, 1).numpy()


def chirp_silence_clip(n_samples: int = 480000) -> np.ndarray:
    """Deterministic chirp (first half) + digital silence (second half): exercises the 1e-10
    floor and the max-8 clamp of the log-mel."""
    t = np.arange(n_samples, dtype=np.float64) / SAMPLE_RATE
    half = n_samples // 2
    f0, f1 = 100.0, 7000.0
    dur = max(half / SAMPLE_RATE, 1e-9)
    phase = 2 * np.pi * (f0 * t + 0.5 * (f1 - f0) / dur * t * t)
    x = 0.5 * np.sin(phase)
    x[half:] = 0.0
    return x.astype(np.float32)


def audio_batch(batch: int = 64, n_samples: int = 480000, seed: int = SEED, device="cpu") -> torch.Tensor:
    """Config 3: randn * 0.1 with a per-row scale U(0.01, 1) so per-clip maxima differ."""
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.randn(batch, n_samples, generator=g, device=device) * 0.1
    scale = torch.rand(batch, 1, generator=g, device=device) * 0.99 + 0.01
    return (x * scale).contiguous()


# ----------------------------------------------------------------------------- video
def landmarks_for_clip(T: int, H: int = 224, W: int = 224, seed: int = SEED, invalid_frac: float = 0.05,
                       integer: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """float64 [T,68,2] (x, y) = mean_face * s + t + jitter with s ~ 0.75 * min(H,W)/224, a slow
    random-walk translation (<= +-6 px) and N(0, 0.5^2) jitter; ``integer`` rounds detections to
    whole pixels as dlib returns them.  valid uint8 [T]: ``invalid_frac`` of the frames are
    flagged as failed detections (never all of them)."""
    rng = np.random.default_rng(seed)
    mf = mean_face_landmarks()
    s = 0.75 * min(H, W) / 224.0
    centre = mf.mean(axis=0) * s
    base = np.array([W / 2.0, H / 2.0]) - centre
    walk = np.cumsum(rng.normal(0.0, 0.4, size=(T, 2)), axis=0)
    walk = np.clip(walk, -6.0, 6.0)
    lm = mf[None] * s + base[None, None] + walk[:, None, :] + rng.normal(0.0, 0.5, size=(T, 68, 2))
    if integer:
        lm = np.rint(lm)
    valid = (rng.random(T) >= invalid_frac).astype(np.uint8)
    if not valid.any():
        valid[T // 2] = 1
    return np.ascontiguousarray(lm, dtype=np.float64), valid


def video_clip(T: int = 250, H: int = 224, W: int = 224, seed: int = SEED, invalid_frac: float = 0.05):
    """Config 2 (numpy, for tests): frames uint8 [T,H,W,3] BGR = smooth low-frequency field +
    uniform noise (so the bilinear warp is non-trivial), landmarks, valid."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    frames = np.empty((T, H, W, 3), dtype=np.uint8)
    for t in range(T):
        for c in range(3):
            field = (127.0 + 60.0 * np.sin(2 * np.pi * (xx / W * (1.5 + c) + 0.010 * t))
                     + 40.0 * np.cos(2 * np.pi * (yy / H * (2.0 + 0.5 * c) - 0.013 * t)))
            noise = rng.integers(-20, 21, size=(H, W)).astype(np.float32)
            frames[t, :, :, c] = np.clip(field + noise, 0, 255).astype(np.uint8)
    lm, valid = landmarks_for_clip(T, H, W, seed + 1, invalid_frac)
    return frames, lm, valid


def video_frames_cuda(n_frames: int, H: int = 224, W: int = 224, seed: int = SEED,
                      device="cuda", chunk: int = 512) -> torch.Tensor:
    """Same kind of content generated on the GPU for bench.py: uint8 [n_frames,H,W,3]."""
    g = torch.Generator(device=device).manual_seed(seed)
    out = torch.empty((n_frames, H, W, 3), dtype=torch.uint8, device=device)
    yy = torch.arange(H, device=device, dtype=torch.float32).view(1, H, 1, 1)
    xx = torch.arange(W, device=device, dtype=torch.float32).view(1, 1, W, 1)
    cc = torch.arange(3, device=device, dtype=torch.float32).view(1, 1, 1, 3)
    for s in range(0, n_frames, chunk):
        e = min(n_frames, s + chunk)
        tt = torch.arange(s, e, device=device, dtype=torch.float32).view(-1, 1, 1, 1)
        field = (127.0 + 60.0 * torch.sin(2 * torch.pi * (xx / W * (1.5 + cc) + 0.010 * tt))
                 + 40.0 * torch.cos(2 * torch.pi * (yy / H * (2.0 + 0.5 * cc) - 0.013 * tt)))
        noise = torch.randint(-20, 21, field.shape, generator=g, device=device).to(torch.float32)
        out[s:e] = (field + noise).clamp_(0, 255).to(torch.uint8)
    return out


# ----------------------------------------------------------------------------- fusion
def fusion_inputs(batch: int = 64, channels: int = 1024, frames: int = 750, seed: int = SEED,
                  dtype=torch.float32, device="cpu", p_drop: float = 0.5, p_audio: float = 0.5):
    """Config 4: fa, fv = randn(B, C, T); mask [B,2] uint8 with P(drop)=0.5, P(audio|drop)=0.5
    (config/avhubert_large.yaml:15-16), never both dropped."""
    g = torch.Generator(device=device).manual_seed(seed)
    fa = torch.randn(batch, channels, frames, generator=g, device=device).to(dtype)
    fv = torch.randn(batch, channels, frames, generator=g, device=device).to(dtype)
    rng = np.random.default_rng(seed)
    mask = np.ones((batch, 2), dtype=np.uint8)
    for b in range(batch):
        if rng.random() < p_drop:
            mask[b, 0 if rng.random() < p_audio else 1] = 0
    return fa, fv, mask


# ----------------------------------------------------------------------------- AMI-shaped sweep
def ami_durations(n_utts: int = 10000, seed: int = SEED) -> np.ndarray:
    """Config 5: d_i = clip(lognormal(ln 3.0, 0.9), 0.3, 30.0) s snapped to whole video frames
    (0.04 s); AMI segments were observed at 0.27-23.81 s."""
    rng = np.random.default_rng(seed)
    d = np.clip(rng.lognormal(np.log(3.0), 0.9, size=n_utts), 0.3, 30.0)
    return np.round(d / 0.04) * 0.04


def shard_indices(n_items: int, rank: int, world_size: int) -> np.ndarray:
    """Utterance shard of one rank: i % world == rank (no collective on the hot path)."""
    return np.arange(rank, n_items, world_size)
