"""Oracle: per-sample trim + batch collation of the video features (CPU, numpy).  TEST
INFRASTRUCTURE ONLY.

* trim: ``AmiVideoHFDataset.__getitem__``, avsl/whisper_flamingo_ft_ami.py:299-302 --
  ``max_len = round(len(audio) / 16000 * 25); video_feats = video_feats[:max_len]`` with
  ``audio`` already padded/trimmed to ``audio_max_length`` (:209-210).
* align: ``align_audio_video_features``, preprocess/audio_process.py:238-264 (truncate the longer
  of the two feature sequences).
* collate: ``WhisperVideoCollatorWithPadding`` (imported at avsl/whisper_flamingo_ft_ami.py:126,
  used at :686, its ``padding_mask`` read at :504).  The class lives in the un-vendored, unpinned
  upstream package ``whisper_flamingo`` (roudimit/whisper-flamingo, ``utils.py``) and is absent
  from the snapshot: PARITY UNPINNED.  Restated from the published upstream: every clip's
  ``[T,88,88,1]`` features are zero-padded along T to the longest clip of the batch
  (``np.pad(..., constant_values=0)``), stacked, permuted ``(0,4,1,2,3)`` to ``[B,1,T,88,88]``;
  the padding mask is ``[B,T]`` bool, True on padded frames; mels (all ``[n_mels,3000]`` after
  pad_or_trim) are stacked.  Labels / decoder ids are text and out of scope.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np


def video_frames_for_audio(n_audio_samples: int, sample_rate: int = 16000, fps: int = 25) -> int:
    """avsl/whisper_flamingo_ft_ami.py:299 (Python round: half to even)."""
    return round(n_audio_samples / sample_rate * fps)


def trim_video(video_feats: np.ndarray, n_audio_samples: int) -> np.ndarray:
    """avsl/whisper_flamingo_ft_ami.py:299-302."""
    m = video_frames_for_audio(n_audio_samples)
    return video_feats[:m] if len(video_feats) > m else video_feats


def align_audio_video_features(audio_features, video_features):
    """Semantics of preprocess/audio_process.py:238-264: both sequences are cut to the shorter
    length; a missing stream leaves the other untouched."""
    if audio_features is None or video_features is None:
        return audio_features, video_features
    n = min(len(audio_features), len(video_features))
    return audio_features[:n], video_features[:n]


def collate_video(videos: Sequence[np.ndarray], mels: Optional[Sequence[np.ndarray]] = None,
                  T_pad: Optional[int] = None) -> Dict[str, np.ndarray]:
    """videos: list of float32 [T_i,88,88,1] (already trimmed).  ``T_pad`` defaults to the
    longest clip (the collator's choice); a larger value pads further."""
    lengths = [len(v) for v in videos]
    T = max(lengths) if T_pad is None else T_pad
    padded = [np.pad(v, ((0, T - n), (0, 0), (0, 0), (0, 0)), "constant", constant_values=0)
              for v, n in zip(videos, lengths)]
    mask = [[False] * n + [True] * (T - n) for n in lengths]
    out = {"video": np.ascontiguousarray(np.transpose(np.array(padded), (0, 4, 1, 2, 3))),
           "padding_mask": np.array(mask, dtype=bool)}
    if mels is not None:
        out["input_ids"] = np.array(list(mels))
    return out
