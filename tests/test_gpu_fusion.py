"""GPU parity: modality fusion through the C ABI vs the oracle (bit-exact)."""
import numpy as np
import pytest
import torch

import avsl_b200 as A
from avsl_b200 import synth
from oracle import fusion as O

pytestmark = pytest.mark.gpu


def _bits(t):
    return t.contiguous().view(torch.int32 if t.element_size() == 4 else torch.int16)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("mode", ["concat", "add", "weighted_sum"])
@pytest.mark.parametrize("masked", [False, True])
def test_fuse_bit_exact_small(dtype, mode, masked):
    fa, fv, mask = synth.fusion_inputs(6, 40, 50, seed=3, dtype=dtype)
    fa[0, 0, 0] = -0.0                                  # signed zeros survive exactly as in torch
    m = mask if masked else None
    out = A.fuse_modalities(fa.cuda(), fv.cuda(), m, mode, weights=(0.3, 0.7)).cpu()
    ref = O.fuse(fa, fv, m, mode, 0.3, 0.7)
    assert out.shape == ref.shape and out.dtype == ref.dtype
    assert torch.equal(_bits(out), _bits(ref))


@pytest.mark.parametrize("shape", [(3, 7, 13), (1, 1, 1), (2, 5, 3)])
def test_fuse_unaligned_shapes(shape):
    B, C, T = shape
    fa, fv, mask = synth.fusion_inputs(B, C, T, seed=5)
    for mode in ("concat", "add", "weighted_sum"):
        out = A.fuse_modalities(fa.cuda(), fv.cuda(), mask, mode).cpu()
        assert torch.equal(_bits(out), _bits(O.fuse(fa, fv, mask, mode)))
    # misaligned base pointers (views offset by one element) take the element-wise kernel
    big = torch.randn(B * C * T + 1).cuda()
    fa_m = big[1:].view(B, C, T)
    out = A.fuse_modalities(fa_m, fv.cuda(), None, "add").cpu()
    assert torch.equal(_bits(out), _bits(O.fuse(fa_m.cpu(), fv, None, "add")))


def test_config4_full_size_concat_and_sum():
    fa, fv, mask = synth.fusion_inputs(64, 1024, 750, device="cuda")
    cat = A.fuse_modalities(fa, fv, mask, "concat")
    assert cat.shape == (64, 2048, 750)
    m = torch.from_numpy(mask).cuda().bool()
    exp_a = torch.where(m[:, 0].view(-1, 1, 1), fa, torch.zeros((), device="cuda"))
    exp_v = torch.where(m[:, 1].view(-1, 1, 1), fv, torch.zeros((), device="cuda"))
    assert torch.equal(cat[:, :1024], exp_a) and torch.equal(cat[:, 1024:], exp_v)
    s = A.fuse_modalities(fa, fv, mask, "add")
    assert torch.equal(s, exp_a + exp_v)
    # size-independent property: sum == first half + second half of concat
    assert torch.equal(s, cat[:, :1024] + cat[:, 1024:])
    h = A.fuse_modalities(fa.half(), fv.half(), mask, "add")
    assert torch.equal(h, (exp_a.half().float() + exp_v.half().float()).half())


def test_errors_follow_reference():
    fa, fv, _ = synth.fusion_inputs(2, 4, 4, device="cuda")
    with pytest.raises(ValueError, match="Unsupported fusion type"):
        A.fuse_modalities(fa, fv, None, "gated")
    with pytest.raises(ValueError, match="At least one input modality"):
        A.fuse_modalities(fa, fv, np.array([[0, 0], [1, 1]]), "add")
    with pytest.raises(ValueError):
        A.fuse_modalities(fa, fv[:, :2], None, "add")


def test_modality_fusion_module_matches_reference_block():
    fa, fv, _ = synth.fusion_inputs(4, 8, 16, device="cuda")
    fuse = A.ModalityFusion("concat", modality_dropout=0.5, audio_dropout=0.5).train()
    np.random.seed(3407)
    outs = [fuse(fa, fv) for _ in range(8)]
    np.random.seed(3407)
    for o in outs:
        ua, uv = O.modality_dropout_flags(True, 0.5, 0.5)
        mask = np.tile(np.array([[ua, uv]], dtype=np.uint8), (4, 1))
        assert torch.equal(o.cpu(), O.fuse(fa.cpu(), fv.cpu(), mask, "concat"))
    fuse.eval()
    assert torch.equal(fuse(fa, fv).cpu(), torch.cat([fa, fv], dim=1).cpu())
    assert torch.equal(fuse(fa, fv, modality_override="visual")[:, :8].cpu(), torch.zeros(4, 8, 16))
    with pytest.raises(ValueError, match="At least one input modality"):
        A.ModalityFusion(use_audio=False)(fa, None)


# tolerance of the fused LayerNorm against torch's CPU kernel (both float32 arithmetic, different
# summation order): 2e-5 absolute on unit-variance output for fp32; one ulp of the output dtype
# (fp16 2^-10, bf16 2^-7 relative) plus that for the half types
_LN_TOL = {torch.float32: 2e-5, torch.float16: 2.5e-3, torch.bfloat16: 2e-2}


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("mode", ["concat", "add", "weighted_sum"])
def test_fuse_transpose_layernorm_vs_reference_ops(dtype, mode):
    B, C, T = 5, 72, 45                               # T not a multiple of the tile, C not of 32
    fa, fv, mask = synth.fusion_inputs(B, C, T, seed=11, dtype=dtype)
    fa = fa * 3.0 + 0.5                               # non-trivial mean and variance
    Cout = 2 * C if mode == "concat" else C
    g = torch.Generator().manual_seed(1)
    w = torch.randn(Cout, generator=g) * 0.5 + 1.0
    b = torch.randn(Cout, generator=g) * 0.1
    for m in (None, mask):
        ref = O.fuse_transpose_layernorm(fa, fv, m, mode, w, b, 1e-5, 0.3, 0.7)
        got = A.fuse_transpose_layernorm(fa.cuda(), fv.cuda(), m, mode, w.cuda(), b.cuda(), 1e-5,
                                         weights=(0.3, 0.7)).cpu()
        assert got.shape == (B, T, Cout) and got.dtype == dtype
        err = (got.float() - ref.float()).abs().max().item()
        assert err <= _LN_TOL[dtype] * max(1.0, ref.float().abs().max().item()), err
    # no affine parameters, and the unfused product path agrees
    got = A.fuse_transpose_layernorm(fa.cuda(), fv.cuda(), None, mode, weights=(0.3, 0.7)).cpu()
    ref = O.fuse_transpose_layernorm(fa, fv, None, mode, None, None, 1e-5, 0.3, 0.7)
    assert (got.float() - ref.float()).abs().max().item() <= _LN_TOL[dtype] * max(1.0, ref.float().abs().max().item())


def test_fuse_transpose_layernorm_config4_properties():
    """Full size (B 64, C 1024, T 750): every output row has zero mean and unit variance (no affine),
    a dropped modality contributes exact zeros before normalisation, and the result equals the
    unfused GPU chain fuse -> transpose -> torch layer_norm to float32 round-off."""
    fa, fv, mask = synth.fusion_inputs(64, 1024, 750, device="cuda")
    out = A.fuse_transpose_layernorm(fa, fv, mask, "concat")
    assert out.shape == (64, 750, 2048)
    assert out.mean(dim=-1).abs().max().item() < 1e-4
    assert (out.var(dim=-1, unbiased=False) - 1.0).abs().max().item() < 1e-3
    cat = A.fuse_modalities(fa, fv, mask, "concat")
    ref = torch.nn.functional.layer_norm(cat.transpose(1, 2), (2048,), None, None, 1e-5)
    assert (out - ref).abs().max().item() < 2e-5 * ref.abs().max().item() + 2e-5
    add = A.fuse_transpose_layernorm(fa.half(), fv.half(), mask, "add")
    ref = torch.nn.functional.layer_norm((A.fuse_modalities(fa.half(), fv.half(), mask, "add")).transpose(1, 2).float(),
                                         (1024,), None, None, 1e-5).half()
    assert (add.float() - ref.float()).abs().max().item() <= 2.5e-3 * ref.float().abs().max().item()


# ------------------------------------------------------------------ gradients
def _ref_fuse(fa, fv, mask, mode, wa=0.5, wv=0.5):
    """The reference's fusion expressions (av_hubert_encoder.py:315-326) with zero-filled missing
    modalities, in torch ops autograd can differentiate."""
    m = torch.as_tensor(np.asarray(mask), device=fa.device).to(fa.dtype)
    a = fa * m[:, 0].view(-1, 1, 1)
    v = fv * m[:, 1].view(-1, 1, 1)
    if mode == "concat":
        return torch.cat([a, v], dim=1)
    if mode == "add":
        return a + v
    return wa * a + wv * v


@pytest.mark.parametrize("mode", ["concat", "add", "weighted_sum"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_fuse_backward_matches_torch_autograd(mode, dtype):
    fa, fv, mask = synth.fusion_inputs(5, 24, 37, seed=11, dtype=dtype)
    fa, fv = fa.cuda().requires_grad_(), fv.cuda().requires_grad_()
    out = A.fuse_modalities(fa, fv, mask, mode, weights=(0.25, 0.75))
    assert out.grad_fn is not None
    g = torch.randn_like(out)
    out.backward(g)
    ra, rv = fa.detach().clone().requires_grad_(), fv.detach().clone().requires_grad_()
    _ref_fuse(ra, rv, mask, mode, 0.25, 0.75).backward(g)
    assert torch.equal(fa.grad, ra.grad) and torch.equal(fv.grad, rv.grad)      # exact: copies / one product


@pytest.mark.parametrize("mode", ["concat", "add", "weighted_sum"])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.float16, 4e-3), (torch.bfloat16, 3e-2)])
@pytest.mark.parametrize("shape", [(3, 40, 75), (2, 256, 33), (1, 520, 100)])
def test_fuse_layernorm_backward_matches_torch_autograd(mode, dtype, tol, shape):
    """Gradients of fusion + transpose + LayerNorm w.r.t. both feature maps and the LayerNorm
    parameters against torch autograd over the reference's own ops (av_hubert_encoder.py:315-330)."""
    B, C, T = shape
    fa, fv, mask = synth.fusion_inputs(B, C, T, seed=13, dtype=dtype)
    Cout = 2 * C if mode == "concat" else C
    gen = torch.Generator().manual_seed(3)
    w0 = (torch.rand(Cout, generator=gen) + 0.5).cuda()
    b0 = torch.randn(Cout, generator=gen).cuda()
    fa, fv = fa.cuda().requires_grad_(), fv.cuda().requires_grad_()
    w, b = w0.clone().requires_grad_(), b0.clone().requires_grad_()
    out = A.fuse_transpose_layernorm(fa, fv, mask, mode, w, b, weights=(0.4, 0.6))
    assert out.grad_fn is not None and out.shape == (B, T, Cout)
    g = torch.randn(out.shape, generator=gen).to(dtype).cuda()
    out.backward(g)
    # reference in float32 from the same (rounded) inputs
    ra, rv = fa.detach().float().requires_grad_(), fv.detach().float().requires_grad_()
    rw, rb = w0.clone().requires_grad_(), b0.clone().requires_grad_()
    fused = _ref_fuse(ra, rv, mask, mode, 0.4, 0.6)
    if dtype != torch.float32:
        fused = fused + (fused.to(dtype).float() - fused).detach()      # the forward rounds the fused value to dtype
    ref = torch.nn.functional.layer_norm(fused.transpose(1, 2), (Cout,), rw, rb, 1e-5)
    ref.backward(g.float())
    scale = max(1.0, ra.grad.abs().max().item())
    assert (fa.grad.float() - ra.grad).abs().max().item() <= tol * scale
    assert (fv.grad.float() - rv.grad).abs().max().item() <= tol * scale
    n_red = B * T
    assert (w.grad - rw.grad).abs().max().item() <= 2e-5 * n_red ** 0.5 * max(1.0, rw.grad.abs().max().item())
    assert (b.grad - rb.grad).abs().max().item() <= 2e-5 * n_red ** 0.5 * max(1.0, rb.grad.abs().max().item())
    # a masked-out modality receives an exactly zero gradient
    m = np.asarray(mask)
    for k in range(B):
        if not m[k, 0]:
            assert not fa.grad[k].any()
        if not m[k, 1]:
            assert not fv.grad[k].any()
    # deterministic: a second backward gives the same bits
    fa2, fv2 = fa.detach().clone().requires_grad_(), fv.detach().clone().requires_grad_()
    w2, b2 = w0.clone().requires_grad_(), b0.clone().requires_grad_()
    A.fuse_transpose_layernorm(fa2, fv2, mask, mode, w2, b2, weights=(0.4, 0.6)).backward(g)
    assert torch.equal(fa2.grad, fa.grad) and torch.equal(w2.grad, w.grad) and torch.equal(b2.grad, b.grad)


def test_modality_fusion_module_trains_and_no_grad_path_is_unchanged():
    fa, fv, mask = synth.fusion_inputs(4, 16, 20, seed=2)
    fa, fv = fa.cuda(), fv.cuda()
    plain = A.fuse_modalities(fa, fv, mask, "concat")
    assert plain.grad_fn is None
    fa_g = fa.clone().requires_grad_()
    with torch.no_grad():
        assert A.fuse_modalities(fa_g, fv, mask, "concat").grad_fn is None
    with pytest.raises(ValueError, match="out="):
        A.fuse_modalities(fa_g, fv, mask, "concat", out=torch.empty_like(plain))
    mod = A.ModalityFusion("add", modality_dropout=0.0).cuda().train()
    out = mod(fa_g, fv)
    out.sum().backward()
    assert torch.equal(fa_g.grad, torch.ones_like(fa_g))
    # a CUDA mask is taken without a host round trip
    cm = torch.as_tensor(np.asarray(mask)).cuda()
    assert torch.equal(A.fuse_modalities(fa, fv, cm, "concat"), plain)


# ------------------------------------------------------------------ TMA-filled fused LayerNorm (padded rows)
@pytest.mark.parametrize("mode", ["concat", "add", "weighted_sum"])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.float16, 2e-3), (torch.bfloat16, 1.6e-2)])
@pytest.mark.parametrize("B,C,T", [(3, 256, 37), (2, 1024, 750), (2, 512, 8), (24, 256, 750)])   # the last: more tiles than resident CTAs
def test_fuse_layernorm_tma_path_matches_oracle(mode, dtype, tol, B, C, T):
    """Feature maps in a padded allocation (row pitch a multiple of 16 bytes) take the tensor-map TMA
    kernel; same numbers as the oracle (the reference's torch ops) and as the LSU kernel."""
    from avsl_b200 import _lib
    fa0, fv0, mask = synth.fusion_inputs(B, C, T, seed=21, dtype=dtype)
    fa, fv = A.alloc_features(B, C, T, dtype, "cuda"), A.alloc_features(B, C, T, dtype, "cuda")
    fa.copy_(fa0); fv.copy_(fv0)
    assert fa.stride(1) % (16 // fa.element_size()) == 0
    assert _lib.load().avfe_fuse_layernorm_tma_ok(A.fusion._DTYPES[dtype], C, T, fa.stride(1)) == 1
    Cout = 2 * C if mode == "concat" else C
    gen = torch.Generator().manual_seed(1)
    w = (torch.rand(Cout, generator=gen) + 0.5).cuda()
    bz = torch.randn(Cout, generator=gen).cuda()
    got = A.fuse_transpose_layernorm(fa, fv, mask, mode, w, bz, weights=(0.3, 0.7))
    assert got.shape == (B, T, Cout) and got.dtype == dtype
    from oracle import fusion as OFU
    ref = OFU.fuse_transpose_layernorm(fa0, fv0, mask, mode, w.cpu(), bz.cpu(), w_a=0.3, w_v=0.7)
    assert (got.cpu().float() - ref.float()).abs().max().item() <= tol * max(1.0, ref.float().abs().max().item())
    lsu = A.fuse_transpose_layernorm(fa0.cuda(), fv0.cuda(), mask, mode, w, bz, weights=(0.3, 0.7))   # contiguous: LSU kernel
    assert (got.float() - lsu.float()).abs().max().item() <= tol * max(1.0, ref.float().abs().max().item())


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.float16, 2e-3)])
def test_fuse_layernorm_tma_full_size_equals_lsu_kernel(dtype, tol):
    """configs[3] size (B 64, C 1024, T 750, masked): the persistent TMA kernel walks ~27 tiles per CTA
    through its two-stage ring; same numbers as the LSU kernel on the contiguous tensors, and rows are
    normalised (zero mean / unit variance with unit weight, zero bias)."""
    B, C, T = 64, 1024, 750
    fa0, fv0, mask = synth.fusion_inputs(B, C, T, seed=5, dtype=dtype, device="cuda")
    fa, fv = A.alloc_features(B, C, T, dtype, "cuda"), A.alloc_features(B, C, T, dtype, "cuda")
    fa.copy_(fa0); fv.copy_(fv0)
    w, bz = torch.ones(2 * C, device="cuda"), torch.zeros(2 * C, device="cuda")
    got = A.fuse_transpose_layernorm(fa, fv, mask, "concat", w, bz)
    lsu = A.fuse_transpose_layernorm(fa0, fv0, mask, "concat", w, bz)
    assert (got.float() - lsu.float()).abs().max().item() <= tol * max(1.0, lsu.float().abs().max().item())
    g = got.float()
    assert g.mean(dim=-1).abs().max().item() < (1e-4 if dtype == torch.float32 else 2e-3)
    assert (g.var(dim=-1, unbiased=False) - 1.0).abs().max().item() < (1e-3 if dtype == torch.float32 else 1e-2)
