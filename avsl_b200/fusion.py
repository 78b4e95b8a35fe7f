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
    """[B,2] keep-mask as a uint8 device tensor (None stays None).  A host mask (numpy / CPU
    tensor) with a sample that drops both modalities raises like the reference
    (``av_hubert_encoder.py:301-302``).  A CUDA mask is used as it is, without a device-to-host
    copy -- the call stays asynchronous -- so that check is the caller's; the kernels zero-fill a
    sample whose two flags are both 0."""
    if mask is None:
        return None
    if torch.is_tensor(mask) and mask.is_cuda:
        if tuple(mask.shape) != (B, 2):
            raise ValueError("mask must be [B, 2]")
        m = mask if mask.dtype == torch.uint8 else (mask != 0).to(torch.uint8)
        return m.contiguous().to(device)
    m = torch.as_tensor(np.asarray(mask) != 0).to(torch.uint8)
    if tuple(m.shape) != (B, 2):
        raise ValueError("mask must be [B, 2]")
    if not bool((m.sum(dim=1) > 0).all()):
        raise ValueError("At least one input modality must be provided and enabled")
    return m.contiguous().to(device, non_blocking=True)


def _needs_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


class _FuseFn(torch.autograd.Function):
    """avfe_fuse with its backward kernel (avfe_fuse_backward): the reference's fusion runs in the
    training forward, so gradients must reach both feature extractors."""

    @staticmethod
    def forward(ctx, fa, fv, m, mode, wa, wv):
        B, C, T = (int(s) for s in fa.shape)
        out = torch.empty((B, 2 * C, T) if mode == _lib.FUSE_CONCAT else (B, C, T), dtype=fa.dtype, device=fa.device)
        with torch.cuda.device(fa.device):
            _lib.call("avfe_fuse", _lib.ptr(fa), _lib.ptr(fv), _lib.ptr(m), mode, wa, wv, _DTYPES[fa.dtype],
                      B, C, T, _lib.ptr(out), _lib.stream_ptr())
        ctx.mask = m
        ctx.meta = (mode, wa, wv, B, C, T)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        mode, wa, wv, B, C, T = ctx.meta
        g = grad_out.contiguous()
        gfa = torch.empty((B, C, T), dtype=g.dtype, device=g.device)
        gfv = torch.empty_like(gfa)
        with torch.cuda.device(g.device):
            _lib.call("avfe_fuse_backward", _lib.ptr(g), _lib.ptr(ctx.mask), mode, wa, wv, _DTYPES[g.dtype],
                      B, C, T, _lib.ptr(gfa), _lib.ptr(gfv), _lib.stream_ptr())
        return (gfa if ctx.needs_input_grad[0] else None, gfv if ctx.needs_input_grad[1] else None,
                None, None, None, None)


class _FuseLayerNormFn(torch.autograd.Function):
    """avfe_fuse_layernorm with its backward kernel (avfe_fuse_layernorm_backward): gradients for
    both feature maps and for layer_norm.weight / .bias; the moments are recomputed in the backward
    pass, so only the inputs are kept."""

    @staticmethod
    def forward(ctx, fa, fv, weight, bias, m, mode, wa, wv, eps):
        B, C, T = (int(s) for s in fa.shape)
        Cout = 2 * C if mode == _lib.FUSE_CONCAT else C
        out = torch.empty((B, T, Cout), dtype=fa.dtype, device=fa.device)
        with torch.cuda.device(fa.device):
            _lib.call("avfe_fuse_layernorm", _lib.ptr(fa), _lib.ptr(fv), _lib.ptr(m), mode, wa, wv,
                      _DTYPES[fa.dtype], B, C, T, _lib.ptr(weight), _lib.ptr(bias), eps, _lib.ptr(out),
                      _lib.stream_ptr())
        ctx.save_for_backward(fa, fv, weight)
        ctx.mask = m
        ctx.meta = (mode, wa, wv, eps, B, C, T, Cout, bias is not None)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        fa, fv, weight = ctx.saved_tensors
        mode, wa, wv, eps, B, C, T, Cout, has_bias = ctx.meta
        g = grad_out.contiguous()
        gfa, gfv = torch.empty_like(fa), torch.empty_like(fv)
        want_w = weight is not None and ctx.needs_input_grad[2]
        want_b = has_bias and ctx.needs_input_grad[3]
        gw = torch.empty(Cout, dtype=torch.float32, device=g.device) if want_w else None
        gb = torch.empty(Cout, dtype=torch.float32, device=g.device) if want_b else None
        with torch.cuda.device(g.device):
            lib = _lib.load()
            ws_bytes = int(lib.avfe_fuse_layernorm_backward_workspace_bytes(B, C, T, mode))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=g.device)
            _lib.call("avfe_fuse_layernorm_backward", _lib.ptr(fa), _lib.ptr(fv), _lib.ptr(ctx.mask), mode, wa, wv,
                      _DTYPES[fa.dtype], B, C, T, _lib.ptr(weight), eps, _lib.ptr(g), _lib.ptr(gfa), _lib.ptr(gfv),
                      _lib.ptr(gw), _lib.ptr(gb), _lib.ptr(ws), ws_bytes, _lib.stream_ptr())
        return (gfa if ctx.needs_input_grad[0] else None, gfv if ctx.needs_input_grad[1] else None, gw, gb,
                None, None, None, None, None)


def fuse_modalities(features_audio: torch.Tensor, features_video: torch.Tensor,
                    mask=None, fusion_type: str = "concat", weights: Tuple[float, float] = (0.5, 0.5),
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """features_audio, features_video: [B, C, T] CUDA tensors of one dtype (fp32/fp16/bf16).
    mask: [B,2] (numpy / tensor; nonzero = present) or None.  Returns [B,2C,T] for "concat",
    [B,C,T] for "add" / "weighted_sum".  Differentiable: when an input requires grad the result
    carries a ``grad_fn`` whose backward is a libavfe kernel (``out=`` cannot be combined with that)."""
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
    m = _mask_tensor(mask, B, fa.device)
    if _needs_grad(fa, fv):
        if out is not None:
            raise ValueError("out= cannot be used when gradients are required")
        return _FuseFn.apply(fa, fv, m, mode, float(weights[0]), float(weights[1]))
    shape = (B, 2 * C, T) if mode == _lib.FUSE_CONCAT else (B, C, T)
    if out is None:
        out = torch.empty(shape, dtype=fa.dtype, device=fa.device)
    elif not (out.is_cuda and out.is_contiguous() and out.dtype == fa.dtype and tuple(out.shape) == shape):
        raise ValueError(f"out must be a contiguous {fa.dtype} CUDA tensor of shape {shape}")
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
    LayerNorm parameters (float32 [C']).  Differentiable with respect to both feature maps,
    ``weight`` and ``bias`` (one backward kernel; ``out=`` cannot be combined with that)."""
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
    B, C, T = (int(s) for s in fa.shape)
    mode = _MODES[fusion_type]
    Cout = 2 * C if mode == _lib.FUSE_CONCAT else C

    def pitched(t):             # unit time stride, rows `pitch` apart, samples C * pitch apart
        return t.stride(2) == 1 and t.stride(0) == C * t.stride(1) and t.stride(1) >= T
    pitch = T
    if pitched(fa) and pitched(fv) and fa.stride(1) == fv.stride(1) and fa.stride(1) != T and not _needs_grad(fa, fv, weight, bias):
        pitch = int(fa.stride(1))       # a padded allocation (alloc_features): TMA-addressable rows
    else:
        fa, fv = fa.contiguous(), fv.contiguous()
    for name, prm in (("weight", weight), ("bias", bias)):
        if prm is not None and not (prm.is_cuda and prm.dtype == torch.float32 and prm.is_contiguous()
                                    and tuple(prm.shape) == (Cout,)):
            raise ValueError(f"{name} must be a contiguous float32 CUDA tensor of shape ({Cout},)")
    m = _mask_tensor(mask, B, fa.device)
    if _needs_grad(fa, fv, weight, bias):
        if out is not None:
            raise ValueError("out= cannot be used when gradients are required")
        return _FuseLayerNormFn.apply(fa, fv, weight, bias, m, mode, float(weights[0]), float(weights[1]), float(eps))
    if out is None:
        out = torch.empty((B, T, Cout), dtype=fa.dtype, device=fa.device)
    elif not (out.is_cuda and out.is_contiguous() and out.dtype == fa.dtype and tuple(out.shape) == (B, T, Cout)):
        raise ValueError(f"out must be a contiguous {fa.dtype} CUDA tensor of shape {(B, T, Cout)}")
    with torch.cuda.device(fa.device):
        _lib.call("avfe_fuse_layernorm_pitched", _lib.ptr(fa), _lib.ptr(fv), _lib.ptr(m), mode, float(weights[0]),
                  float(weights[1]), _DTYPES[fa.dtype], B, C, T, pitch, _lib.ptr(weight), _lib.ptr(bias),
                  float(eps), _lib.ptr(out), _lib.stream_ptr())
    return out


# --------------------------------------------------------------------------- post_extract_proj
def alloc_features(B: int, C: int, T: int, dtype=torch.float16, device=None) -> torch.Tensor:
    """A ``[B, C, T]`` feature map whose time rows start on 16-byte boundaries (row pitch = T rounded
    up to a multiple of 8): what :func:`fuse_layernorm_project` needs for its TMA loads.  Let the
    feature extractor write into it (``out=`` / ``copy_``); the padding is never read."""
    pitch = (T + 7) // 8 * 8
    return torch.empty((B, C, pitch), dtype=dtype, device=device)[:, :, :T]


class FoldedProjection:
    """``post_extract_proj`` (``nn.Linear(2C, D)``) with the preceding LayerNorm's affine folded into
    it once (``avfe_proj_fold``): ``W' = gamma * W``, ``s = sum_k W'``, ``c = W beta + bias``."""

    def __init__(self, proj_weight: torch.Tensor, proj_bias: Optional[torch.Tensor], ln_weight: Optional[torch.Tensor],
                 ln_bias: Optional[torch.Tensor], dtype=torch.float16):
        _lib.require_cuda()
        if dtype not in (torch.float16, torch.bfloat16):
            raise ValueError("the tensor-core projection runs in float16 or bfloat16")
        if proj_weight.dim() != 2 or not proj_weight.is_cuda:
            raise ValueError("proj_weight must be a CUDA tensor [D, 2C]")
        W = proj_weight.detach().contiguous()
        if W.dtype not in (torch.float32, dtype):
            W = W.float()
        self.D, self.K = int(W.shape[0]), int(W.shape[1])
        self.dtype = dtype
        f32 = lambda t: None if t is None else t.detach().float().contiguous()
        g, b, bias = f32(ln_weight), f32(ln_bias), f32(proj_bias)
        lib = _lib.load()
        self.buffer = torch.empty(int(lib.avfe_proj_fold_bytes(self.D, self.K)), dtype=torch.uint8, device=W.device)
        with torch.cuda.device(W.device):
            _lib.call("avfe_proj_fold", _lib.ptr(W), _DTYPES[W.dtype], _lib.ptr(g), _lib.ptr(b), _lib.ptr(bias),
                      self.D, self.K, _DTYPES[dtype], _lib.ptr(self.buffer), _lib.stream_ptr())


def fuse_layernorm_project(features_audio: torch.Tensor, features_video: torch.Tensor, mask, folded: FoldedProjection,
                           eps: float = 1e-5, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``post_extract_proj(layer_norm(cat([fa, fv], 1).transpose(1, 2)))`` (av_hubert_encoder.py:315-334)
    in one pass on the tensor cores: ``[B,C,T] x 2 -> [B,T,D]`` (float16 / bfloat16).  The feature maps
    must have unit stride in time and a row pitch that is a multiple of 8 elements (``alloc_features``).
    This entry point is the inference form (weights folded once); :class:`FusedProjection` is the module
    that also trains (its backward recomputes the normalised features and reuses the LayerNorm backward kernel)."""
    _lib.require_cuda()
    fa, fv = features_audio, features_video
    if _needs_grad(fa, fv):
        raise RuntimeError("fuse_layernorm_project has no backward of its own; use avsl_b200.FusedProjection when training")
    if not (fa.is_cuda and fv.is_cuda) or fa.shape != fv.shape or fa.dim() != 3 or fa.dtype != fv.dtype:
        raise ValueError("features_audio and features_video must both be CUDA [B, C, T] of one dtype")
    if fa.dtype != folded.dtype:
        raise ValueError(f"features are {fa.dtype} but the projection was folded for {folded.dtype}")
    B, C, T = (int(s) for s in fa.shape)
    if 2 * C != folded.K:
        raise ValueError(f"projection expects {folded.K} fused channels, got 2 x {C}")
    for t in (fa, fv):
        if t.stride(2) != 1 or t.stride(1) % 8 != 0 or t.stride(0) != C * t.stride(1) or t.stride(1) != fa.stride(1):
            raise ValueError("feature maps need unit time stride and a row pitch that is a multiple of 8 elements "
                             "(allocate them with avsl_b200.alloc_features): tensor-map TMA loads 16-byte aligned rows")
    m = _mask_tensor(mask, B, fa.device)
    if out is None:
        out = torch.empty((B, T, folded.D), dtype=fa.dtype, device=fa.device)
    elif not (out.is_cuda and out.is_contiguous() and out.dtype == fa.dtype and tuple(out.shape) == (B, T, folded.D)):
        raise ValueError(f"out must be a contiguous {fa.dtype} CUDA tensor of shape {(B, T, folded.D)}")
    with torch.cuda.device(fa.device):
        lib = _lib.load()
        ws_bytes = int(lib.avfe_fuse_ln_proj_workspace_bytes(B, T))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=fa.device)
        _lib.call("avfe_fuse_ln_proj", _lib.ptr(fa), _lib.ptr(fv), _lib.ptr(m), _DTYPES[fa.dtype], B, C, T,
                  int(fa.stride(1)), _lib.ptr(folded.buffer), folded.D, float(eps), _lib.ptr(out), _lib.ptr(ws), ws_bytes,
                  _lib.stream_ptr())
    return out


def _padded(t: torch.Tensor) -> torch.Tensor:
    """``t`` itself when its rows are TMA-addressable, else a copy in an ``alloc_features`` buffer."""
    B, C, T = (int(s) for s in t.shape)
    if t.stride(2) == 1 and t.stride(1) % 8 == 0 and t.stride(0) == C * t.stride(1):
        return t
    p = alloc_features(B, C, T, t.dtype, t.device)
    p.copy_(t)
    return p


class _FuseLayerNormProjectFn(torch.autograd.Function):
    """Training form of :func:`fuse_layernorm_project`.  Forward: the LayerNorm affine is folded into the
    current projection weights (``avfe_proj_fold``, microseconds) and the tcgen05 kernel runs as in
    inference.  Backward (activation recomputation, nothing but the inputs is kept): the normalised
    features are rebuilt by the fused LayerNorm kernel, ``dLN = dY W`` and ``dW = dY^T LN`` are two
    library GEMMs, and ``avfe_fuse_layernorm_backward`` carries ``dLN`` back to both feature maps and
    to ``layer_norm.weight / .bias``."""

    @staticmethod
    def forward(ctx, fa, fv, ln_w, ln_b, W, pb, m, eps):
        folded = FoldedProjection(W, pb, ln_w, ln_b, fa.dtype)
        out = fuse_layernorm_project(_padded(fa), _padded(fv), m, folded, eps)
        ctx.save_for_backward(fa, fv, ln_w, ln_b, W)
        ctx.mask, ctx.eps, ctx.has_pb = m, eps, pb is not None
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        fa, fv, ln_w, ln_b, W = ctx.saved_tensors
        B, C, T = (int(s) for s in fa.shape)
        D = int(W.shape[0])
        dt = fa.dtype
        fa_c, fv_c = fa.contiguous(), fv.contiguous()
        g2 = grad_out.contiguous().reshape(B * T, D)
        ln = fuse_transpose_layernorm(fa_c, fv_c, ctx.mask, "concat", ln_w, ln_b, ctx.eps)       # [B, T, 2C], recomputed
        d_ln = (g2 @ W.to(dt)).view(B, T, 2 * C)
        gW = (g2.t() @ ln.reshape(B * T, 2 * C)).to(W.dtype) if ctx.needs_input_grad[4] else None
        gpb = g2.float().sum(dim=0) if (ctx.has_pb and ctx.needs_input_grad[5]) else None
        del ln
        gfa, gfv = torch.empty_like(fa_c), torch.empty_like(fv_c)
        want_w = ln_w is not None and ctx.needs_input_grad[2]
        want_b = ln_b is not None and ctx.needs_input_grad[3]
        gw = torch.empty(2 * C, dtype=torch.float32, device=fa.device) if want_w else None
        gb = torch.empty(2 * C, dtype=torch.float32, device=fa.device) if want_b else None
        with torch.cuda.device(fa.device):
            lib = _lib.load()
            ws_bytes = int(lib.avfe_fuse_layernorm_backward_workspace_bytes(B, C, T, _lib.FUSE_CONCAT))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=fa.device)
            _lib.call("avfe_fuse_layernorm_backward", _lib.ptr(fa_c), _lib.ptr(fv_c), _lib.ptr(ctx.mask), _lib.FUSE_CONCAT,
                      0.5, 0.5, _DTYPES[dt], B, C, T, _lib.ptr(ln_w), ctx.eps, _lib.ptr(d_ln.contiguous()), _lib.ptr(gfa),
                      _lib.ptr(gfv), _lib.ptr(gw), _lib.ptr(gb), _lib.ptr(ws), ws_bytes, _lib.stream_ptr())
        return (gfa if ctx.needs_input_grad[0] else None, gfv if ctx.needs_input_grad[1] else None, gw, gb, gW,
                None if gpb is None else gpb.to(W.dtype), None, None)


class FusedProjection(torch.nn.Module):
    """``cat`` + ``transpose(1, 2)`` + ``self.layer_norm`` + ``self.post_extract_proj`` of
    ``AVHuBERTEncoderWrapper.forward`` (av_hubert_encoder.py:315-334) as one module holding the same
    parameters (``layer_norm.weight / .bias`` float32 ``[2C]``, ``post_extract_proj.weight [D, 2C]`` /
    ``.bias [D]``).  Inference: the folded projection is cached and reused until a parameter changes.
    Training: every output carries a ``grad_fn`` (:class:`_FuseLayerNormProjectFn`), so gradients reach
    the feature extractors and all four parameters.  Features are float16 / bfloat16 ``[B, C, T]``."""

    def __init__(self, channels: int, out_dim: int, eps: float = 1e-5, dtype=torch.float16, device=None):
        super().__init__()
        self.layer_norm = torch.nn.LayerNorm(2 * channels, eps=eps, device=device)
        self.post_extract_proj = torch.nn.Linear(2 * channels, out_dim, device=device)
        self.dtype = dtype
        self._folded = None
        self._folded_key = None

    def _params(self):
        return (self.layer_norm.weight, self.layer_norm.bias, self.post_extract_proj.weight, self.post_extract_proj.bias)

    def forward(self, features_audio: torch.Tensor, features_video: torch.Tensor, mask=None) -> torch.Tensor:
        ln_w, ln_b, W, pb = self._params()
        fa, fv = features_audio, features_video
        if _needs_grad(fa, fv, ln_w, ln_b, W, pb):
            m = _mask_tensor(mask, int(fa.shape[0]), fa.device)
            return _FuseLayerNormProjectFn.apply(fa, fv, ln_w.float(), ln_b.float(), W, pb, m, float(self.layer_norm.eps))
        key = tuple((p.data_ptr(), p._version) for p in self._params()) + (fa.dtype,)
        if self._folded is None or self._folded_key != key:
            self._folded = FoldedProjection(W, pb, ln_w, ln_b, fa.dtype)
            self._folded_key = key
        return fuse_layernorm_project(_padded(fa), _padded(fv), mask, self._folded, float(self.layer_norm.eps))


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
