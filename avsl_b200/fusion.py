"""Modality dropout + fusion: host-side mirror of the fusion block of
``AVHuBERTEncoderWrapper.forward`` (``avsl/modules/av_hubert_encoder.py:273-326``).

``fusion_type`` keeps the reference's vocabulary ("concat", "add"; "weighted_sum" is the mode
the reference advertises in ``config/avhubert_large.yaml:14`` but raises on).  A missing
modality is zero-filled, as upstream av_hubert does (and as the reference's own patch
``avsl/scripts/preparation/setup_whisper_flamingo_env.sh:51-58`` enforces).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib

_MODES = {"concat": _lib.FUSE_CONCAT, "add": _lib.FUSE_SUM, "sum": _lib.FUSE_SUM,
          "weighted_sum": _lib.FUSE_WSUM}
_DTYPES = {torch.float32: _lib.AVFE_F32, torch.float16: _lib.AVFE_F16, torch.bfloat16: _lib.AVFE_BF16}


def modality_dropout_flags(training: bool, modality_dropout: float, audio_dropout: float,
                           rng=np.random) -> Tuple[bool, bool]:
    """``av_hubert_encoder.py:292-298``: two uniform draws per forward (drawn even in eval, as
    the reference does); in training, with probability ``modality_dropout`` drop audio (with
    probability ``audio_dropout``) or else video.  Returns (use_audio, use_visual)."""
    modality_drop_prob, audio_drop_prob = rng.random(), rng.random()
    use_audio, use_visual = True, True
    if training and modality_drop_prob < modality_dropout:
        if audio_drop_prob < audio_dropout:
            use_audio = False
        else:
            use_visual = False
    return use_audio, use_visual


def modality_dropout_mask(batch_size: int, training: bool, modality_dropout: float,
                          audio_dropout: float, rng=np.random, per_sample: bool = False) -> np.ndarray:
    """[B, 2] uint8 keep-mask (col 0 audio, col 1 video).  ``per_sample=False`` reproduces the
    reference (one decision for the whole batch); ``True`` draws one decision per sample."""
    mask = np.ones((batch_size, 2), dtype=np.uint8)
    if per_sample:
        for b in range(batch_size):
            ua, uv = modality_dropout_flags(training, modality_dropout, audio_dropout, rng)
            mask[b] = (ua, uv)
    else:
        ua, uv = modality_dropout_flags(training, modality_dropout, audio_dropout, rng)
        mask[:] = (ua, uv)
    return mask


def _mask_tensor(mask, B: int, device) -> Optional[torch.Tensor]:
    """[B,2] keep-mask as a uint8 device tensor (None stays None); a sample with both modalities
    dropped raises like the reference (``av_hubert_encoder.py:301-302``)."""
    if mask is None:
        return None
    m = torch.as_tensor(np.asarray(mask.cpu() if torch.is_tensor(mask) else mask) != 0).to(torch.uint8)
    if tuple(m.shape) != (B, 2):
        raise ValueError("mask must be [B, 2]")
    if not bool((m.sum(dim=1) > 0).all()):
        raise ValueError("At least one input modality must be provided and enabled")
    return m.contiguous().to(device, non_blocking=True)


def fuse_modalities(features_audio: torch.Tensor, features_video: torch.Tensor,
                    mask=None, fusion_type: str = "concat", weights: Tuple[float, float] = (0.5, 0.5),
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """features_audio, features_video: [B, C, T] CUDA tensors of one dtype (fp32/fp16/bf16).
    mask: [B,2] (numpy / tensor; nonzero = present) or None.  Returns [B,2C,T] for "concat",
    [B,C,T] for "add" / "weighted_sum"."""
    if fusion_type not in _MODES:
        raise ValueError(f"Unsupported fusion type: {fusion_type}")
    _lib.require_cuda()
    fa, fv = features_audio, features_video
    if not (fa.is_cuda and fv.is_cuda):
        raise ValueError("fuse_modalities takes CUDA tensors")
    if fa.shape != fv.shape or fa.dim() != 3 or fa.dtype != fv.dtype:
        raise ValueError("features_audio and features_video must both be [B, C, T] of one dtype")
    if fa.dtype not in _DTYPES:
        raise ValueError(f"unsupported dtype {fa.dtype}")
    fa, fv = fa.contiguous(), fv.contiguous()
    B, C, T = (int(s) for s in fa.shape)
    mode = _MODES[fusion_type]
    shape = (B, 2 * C, T) if mode == _lib.FUSE_CONCAT else (B, C, T)
    if out is None:
        out = torch.empty(shape, dtype=fa.dtype, device=fa.device)
    elif not (out.is_cuda and out.is_contiguous() and out.dtype == fa.dtype and tuple(out.shape) == shape):
        raise ValueError(f"out must be a contiguous {fa.dtype} CUDA tensor of shape {shape}")
    m = _mask_tensor(mask, B, fa.device)
    with torch.cuda.device(fa.device):
        _lib.call("avfe_fuse", _lib.ptr(fa), _lib.ptr(fv), _lib.ptr(m), mode, float(weights[0]),
                  float(weights[1]), _DTYPES[fa.dtype], B, C, T, _lib.ptr(out), _lib.stream_ptr())
    return out


def fuse_transpose_layernorm(features_audio: torch.Tensor, features_video: torch.Tensor, mask=None,
                             fusion_type: str = "concat", weight: Optional[torch.Tensor] = None,
                             bias: Optional[torch.Tensor] = None, eps: float = 1e-5,
                             weights: Tuple[float, float] = (0.5, 0.5),
                             out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Fusion plus the two lines that follow it in ``AVHuBERTEncoderWrapper.forward``
    (``av_hubert_encoder.py:329-330``): ``features.transpose(1, 2)`` and ``self.layer_norm``
    (``LayerNorm`` of ``av_hubert_layers.py:438-440``: float32 arithmetic, result cast back), in
    one kernel.  [B,C,T] x 2 -> [B,T,C'] (C' = 2C for "concat").  ``weight`` / ``bias`` are the
    LayerNorm parameters (float32 [C'])."""
    if fusion_type not in _MODES:
        raise ValueError(f"Unsupported fusion type: {fusion_type}")
    _lib.require_cuda()
    fa, fv = features_audio, features_video
    if not (fa.is_cuda and fv.is_cuda):
        raise ValueError("fuse_transpose_layernorm takes CUDA tensors")
    if fa.shape != fv.shape or fa.dim() != 3 or fa.dtype != fv.dtype:
        raise ValueError("features_audio and features_video must both be [B, C, T] of one dtype")
    if fa.dtype not in _DTYPES:
        raise ValueError(f"unsupported dtype {fa.dtype}")
    fa, fv = fa.contiguous(), fv.contiguous()
    B, C, T = (int(s) for s in fa.shape)
    mode = _MODES[fusion_type]
    Cout = 2 * C if mode == _lib.FUSE_CONCAT else C
    for name, prm in (("weight", weight), ("bias", bias)):
        if prm is not None and not (prm.is_cuda and prm.dtype == torch.float32 and prm.is_contiguous()
                                    and tuple(prm.shape) == (Cout,)):
            raise ValueError(f"{name} must be a contiguous float32 CUDA tensor of shape ({Cout},)")
    if out is None:
        out = torch.empty((B, T, Cout), dtype=fa.dtype, device=fa.device)
    elif not (out.is_cuda and out.is_contiguous() and out.dtype == fa.dtype and tuple(out.shape) == (B, T, Cout)):
        raise ValueError(f"out must be a contiguous {fa.dtype} CUDA tensor of shape {(B, T, Cout)}")
    m = _mask_tensor(mask, B, fa.device)
    with torch.cuda.device(fa.device):
        _lib.call("avfe_fuse_layernorm", _lib.ptr(fa), _lib.ptr(fv), _lib.ptr(m), mode, float(weights[0]),
                  float(weights[1]), _DTYPES[fa.dtype], B, C, T, _lib.ptr(weight), _lib.ptr(bias),
                  float(eps), _lib.ptr(out), _lib.stream_ptr())
    return out


class ModalityFusion(torch.nn.Module):
    """The modality-select / dropout / fuse block of ``AVHuBERTEncoderWrapper.forward`` as a
    module: same flags (``use_audio``, ``use_visual``, ``modality_override``), same dropout draw,
    same errors; fusion itself is one libavfe kernel."""

    def __init__(self, fusion_type: str = "concat", modality_dropout: float = 0.0,
                 audio_dropout: float = 0.0, use_audio: bool = True, use_visual: bool = True,
                 weights: Tuple[float, float] = (0.5, 0.5)):
        super().__init__()
        self.fusion_type = fusion_type
        self.modality_dropout = modality_dropout
        self.audio_dropout = audio_dropout
        self.use_audio = use_audio
        self.use_visual = use_visual
        self.weights = weights

    def forward(self, features_audio: Optional[torch.Tensor], features_video: Optional[torch.Tensor],
                modality_override: Optional[str] = None) -> torch.Tensor:
        use_audio = self.use_audio and features_audio is not None
        use_visual = self.use_visual and features_video is not None
        if modality_override == "audio":
            use_visual = False
        elif modality_override == "visual":
            use_audio = False
        drop_a, drop_v = modality_dropout_flags(self.training, self.modality_dropout, self.audio_dropout)
        use_audio, use_visual = use_audio and drop_a, use_visual and drop_v
        if not (use_audio or use_visual):
            raise ValueError("At least one input modality must be provided and enabled")
        if self.fusion_type not in _MODES:
            raise ValueError(f"Unsupported fusion type: {self.fusion_type}")
        ref = features_audio if features_audio is not None else features_video
        fa = features_audio if features_audio is not None else ref   # never read when masked out
        fv = features_video if features_video is not None else ref
        mask = np.tile(np.array([[use_audio, use_visual]], dtype=np.uint8), (ref.shape[0], 1))
        return fuse_modalities(fa, fv, mask, self.fusion_type, self.weights)
