"""Oracle: AV-HuBERT modality dropout + fusion (CPU torch).  TEST INFRASTRUCTURE ONLY.

Restates ``AVHuBERTEncoderWrapper.forward``'s inner block,
``avsl/modules/av_hubert_encoder.py:292-298`` (modality dropout) and ``:315-326`` (fusion).

* ``concat`` -> ``torch.cat([fa, fv], dim=1)``; ``add`` -> ``fa + fv`` — the reference's two
  expressions, bit-exact by construction.
* a missing modality is ZERO-FILLED, which is what upstream av_hubert does and what the
  reference's own patch enforces (``avsl/scripts/preparation/setup_whisper_flamingo_env.sh:51-58``:
  ``features_audio = torch.zeros_like(features_video)``); the per-sample ``[B,2]`` mask is the
  batched generalisation of the reference's per-forward flags.
* ``weighted_sum`` is advertised (``config/avhubert_large.yaml:14``) but raises ``ValueError`` in
  the reference (``av_hubert_encoder.py:321-322``) — PARITY UNPINNED, build-defined as
  ``w_a * fa_masked + w_v * fv_masked`` evaluated as two float multiplies and one add.
"""
from __future__ import annotations

import numpy as np
import torch

MODES = {"concat": 0, "add": 1, "sum": 1, "weighted_sum": 2}


def modality_dropout_flags(training: bool, modality_dropout: float, audio_dropout: float,
                           rng=np.random):
    """av_hubert_encoder.py:292-298 — returns (use_audio, use_visual) for one forward."""
    modality_drop_prob, audio_drop_prob = rng.random(), rng.random()
    use_audio, use_visual = True, True
    if training:
        if modality_drop_prob < modality_dropout:
            if audio_drop_prob < audio_dropout:
                use_audio = False
            else:
                use_visual = False
    return use_audio, use_visual


def fuse(fa: torch.Tensor, fv: torch.Tensor, mask=None, mode: str = "concat",
         w_a: float = 0.5, w_v: float = 0.5) -> torch.Tensor:
    """fa, fv: [B, C, T] (same dtype).  mask: [B, 2] (col 0 = audio present, col 1 = video
    present) or None.  Returns [B, 2C, T] for concat, [B, C, T] otherwise."""
    if mode not in MODES:
        raise ValueError(f"Unsupported fusion type: {mode}")
    if mask is not None:
        m = torch.as_tensor(np.asarray(mask)).to(torch.bool)
        zero = torch.zeros((), dtype=fa.dtype)
        fa = torch.where(m[:, 0].view(-1, 1, 1), fa, zero)
        fv = torch.where(m[:, 1].view(-1, 1, 1), fv, zero)
    code = MODES[mode]
    if code == 0:
        return torch.cat([fa, fv], dim=1)
    if code == 1:
        return fa + fv
    wa = torch.tensor(w_a, dtype=torch.float32)
    wv = torch.tensor(w_v, dtype=torch.float32)
    # products rounded separately in fp32, then one add, result cast to the input dtype
    return (wa * fa.to(torch.float32) + wv * fv.to(torch.float32)).to(fa.dtype)


def fuse_transpose_layernorm(fa: torch.Tensor, fv: torch.Tensor, mask=None, mode: str = "concat",
                             weight=None, bias=None, eps: float = 1e-5, w_a: float = 0.5,
                             w_v: float = 0.5) -> torch.Tensor:
    """av_hubert_encoder.py:315-330: fusion, ``features.transpose(1, 2)``, ``self.layer_norm``.
    ``LayerNorm`` is ``nn.LayerNorm`` run on ``x.float()`` and cast back to ``x.dtype``
    (avsl/modules/av_hubert_layers.py:438-440); these are the reference's own torch ops, so the
    oracle IS the reference arithmetic (CPU).  Returns [B, T, C']."""
    x = fuse(fa, fv, mask, mode, w_a, w_v).transpose(1, 2)
    y = torch.nn.functional.layer_norm(x.float(), (x.shape[-1],), weight, bias, eps)
    return y.type(x.dtype).contiguous()
