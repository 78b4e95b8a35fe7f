"""CPU baseline runner: the oracle restatement of the reference's CPU path, timed by
``bench.py`` (``cpu_baseline`` leg and ``--impl reference`` arm).  TEST/BENCH INFRASTRUCTURE ONLY.

Per utterance it does what the reference does on the host
(``avsl/whisper_flamingo_ft_ami.py:209-213,279-302`` and ``preprocess/video_process.py:305-490``
without decoding and dlib): pad_or_trim -> log-mel (torch.stft, all torch threads); BGR->gray ->
landmark interpolation -> window-smoothed similarity fit -> bilinear warp -> cut_patch ->
crop 88 -> normalise.  The lip work is spread over a ``multiprocessing.Pool`` like the
reference's ``batch_process_lip_videos`` (``preprocess/video_process.py:777-799``), but in
chunks of frames rather than whole videos so that every core stays busy on a small sample.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time
from typing import Optional, Sequence

import numpy as np
import torch

from . import lips as OL
from . import logmel as OM

_DATA = None      # (videos, filled landmarks, mean_face) inherited by the forked workers
CHUNK = 48


def _worker_init():
    torch.set_num_threads(1)


def _lip_chunk(job):
    clip, start, stop = job
    videos, landmarks, mean_face = _DATA
    lm = landmarks[clip]
    T = len(lm)
    margin = min(T, OL.WINDOW_MARGIN)
    acc = 0.0
    for i in range(start, stop):
        gray = OL.bgr2gray(videos[clip][i])
        j = min(i, T - margin)                                     # preprocess/video_process.py:417-464
        smoothed = np.mean(lm[j:j + margin], axis=0)
        tf = OL.SimilarityTransform(OL.umeyama(smoothed[OL.STABLE_IDS], mean_face[OL.STABLE_IDS], True))
        t_lm = tf(lm[i])
        r0, c0 = OL.cut_patch_origin(t_lm[48:68], 48, 48, OL.STD_SIZE)
        roi = OL.to_u8(OL.warp_float(gray, tf.inverse.params, OL.STD_SIZE, (r0, r0 + 96), (c0, c0 + 96)))
        feats = OL.video_feats_from_u8(roi[None])
        acc += float(feats[0, 0, 0, 0]) + float(gray[0, 0])
    return acc


class CpuFrontend:
    """Holds one sample of utterances and runs the CPU path over it repeatedly."""

    def __init__(self, audios: Sequence[np.ndarray], videos: Sequence[np.ndarray],
                 landmarks: Sequence[np.ndarray], valids: Sequence[np.ndarray], mean_face: np.ndarray,
                 n_mels: int = 80, audio_max_length: int = 480000, workers: Optional[int] = None):
        global _DATA
        self.audios = [np.asarray(a, dtype=np.float32) for a in audios]
        self.n_mels, self.audio_max_length = n_mels, audio_max_length
        filled = []
        for lm, v in zip(landmarks, valids):
            lst = OL.landmarks_interpolate([lm[i] if v[i] else None for i in range(len(lm))])
            filled.append(np.asarray(lst, dtype=np.float64))
        _DATA = (list(videos), filled, mean_face)
        self.jobs = [(c, s, min(s + CHUNK, len(v))) for c, v in enumerate(videos) for s in range(0, len(v), CHUNK)]
        self.workers = workers or max(1, (os.cpu_count() or 2) - 1)
        # fork: workers inherit _DATA without pickling; they only ever run numpy code
        self.pool = mp.get_context("fork").Pool(processes=self.workers, initializer=_worker_init)
        self.n_frames = sum(len(v) for v in videos)

    def run(self) -> float:
        """One pass over the sample; returns elapsed seconds."""
        t0 = time.perf_counter()
        res = self.pool.map_async(_lip_chunk, self.jobs, chunksize=1)
        padded = np.stack([OM.pad_or_trim(a, self.audio_max_length) for a in self.audios])
        mel = OM.log_mel_spectrogram(torch.from_numpy(padded), self.n_mels)
        _ = float(mel[0, 0, 0])
        res.get()
        return time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()
