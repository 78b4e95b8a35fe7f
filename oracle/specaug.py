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
