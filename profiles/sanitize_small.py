"""Tiny pass over every kernel family (a few frames, three short clips): the smallest program that
launches each libavfe kernel once, for ncu / debugging sessions.  (compute-sanitizer is closed on
this GPU pool; out-of-bounds accesses are caught instead by the parity tests on ragged and
edge-case shapes and by the two-implementation cross-checks in tests/test_gpu_lips.py.)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import avsl_b200 as A  # noqa: E402
from avsl_b200 import lips as L  # noqa: E402
from avsl_b200 import synth  # noqa: E402

clips = [synth.video_clip(t, 96, 128, seed=3 + t, invalid_frac=0.2) for t in (5, 14, 3)]
F = torch.from_numpy(np.concatenate([c[0] for c in clips])).cuda()
LM = torch.from_numpy(np.concatenate([c[1] for c in clips])).cuda()
V = torch.from_numpy(np.concatenate([c[2] for c in clips])).cuda()
off = torch.tensor([0, 5, 19, 22], dtype=torch.int64).cuda()
L.lip_roi_batch(F, off, LM, V, want_gray=True, want_u8=False)          # frame-owner kernel, 88
L.lip_roi_batch(F, off, LM, V, want_gray=True, want_u8=True)           # frame-owner kernel, 96
L.lip_roi_batch(F, off, LM, V, want_gray=False, want_u8=True)          # generic path
L.lip_roi_collate(F, off, LM, V, T_pad=16)
a = torch.from_numpy(np.concatenate([synth.audio_clip(n, i) for i, n in enumerate((20000, 32000, 700))])).cuda()
aoff = torch.tensor([0, 20000, 52000, 52700], dtype=torch.int64).cuda()
mel = A.log_mel_spectrogram_ragged(a, aoff, 32000, 80)
A.spec_augment(mel, audio_frames=[125, 200, 4])
A.logfbank_batch(a, [0, 20000, 52000, 52700], 4, True)
fa, fv, mask = synth.fusion_inputs(3, 40, 50, seed=1)
for mode in ("concat", "add", "weighted_sum"):
    A.fuse_modalities(fa.cuda(), fv.cuda(), mask, mode)
    A.fuse_transpose_layernorm(fa.cuda(), fv.cuda(), mask, mode)
    A.fuse_transpose_layernorm(fa.half().cuda(), fv.half().cuda(), mask, mode)
torch.cuda.synchronize()
print("sanitize_small done")
