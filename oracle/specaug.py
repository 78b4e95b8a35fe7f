"""Oracle: SpecAugment masks (CPU, numpy).  TEST INFRASTRUCTURE ONLY.

The reference calls the upstream ``whisper_flamingo.spec_augment.spec_augment(mel.T, audio_frames=...)``
(avsl/whisper_flamingo_ft_ami.py:216-224) with policies "ls-double" / "ls-basic"; that package is
un-vendored and absent: PARITY UNPINNED (mask sampling follows the SpecAugment paper's LibriSpeech
policies, documented in avsl_b200/audio.py).  What this file pins is the application of a given set
of rectangles, which is plain assignment."""
import numpy as np


def apply_bands(mel: np.ndarray, bands: np.ndarray, fill: float = 0.0) -> np.ndarray:
    """mel [B, n_mels, n_frames]; bands [B, n_bands, 4] = (f0, f1, t0, t1)."""
    out = mel.copy()
    B, n_mels, n_frames = out.shape
    for b in range(B):
        for f0, f1, t0, t1 in bands[b]:
            f0, f1 = max(int(f0), 0), min(int(f1), n_mels)
            t0, t1 = max(int(t0), 0), min(int(t1), n_frames)
            if f1 > f0 and t1 > t0:
                out[b, f0:f1, t0:t1] = fill
    return out


def time_warp(mel: np.ndarray, warp: np.ndarray) -> np.ndarray:
    """SpecAugment time warping as avfe_spec_time_warp_f32 defines it (PARITY UNPINNED: build-defined,
    see the kernel's header): mel [B, n_mels, n_frames] float32, warp [B, 3] = (tau, center, warped)."""
    out = mel.copy()
    B, n_mels, n_frames = mel.shape
    f32 = np.float32
    for b in range(B):
        tau, center, warped = (int(v) for v in warp[b])
        tau = min(max(tau, 0), n_frames)
        if not (tau >= 2 and 0 < center < tau and 0 < warped < tau):
            continue
        a0 = f32(center) / f32(warped)
        a1 = f32(tau - center) / f32(tau - warped)
        t = np.arange(tau)
        s = np.where(t < warped, t.astype(f32) * a0, f32(center) + (t - warped).astype(f32) * a1).astype(f32)
        s0 = np.clip(np.floor(s).astype(np.int64), 0, tau - 1)
        s1 = np.minimum(s0 + 1, tau - 1)
        w = np.clip((s - s0.astype(f32)).astype(f32), f32(0), f32(1))
        out[b, :, :tau] = ((f32(1) - w) * mel[b][:, s0]).astype(f32) + (w * mel[b][:, s1]).astype(f32)
    return out
