"""GPU parity: post_extract_proj fused behind concat + transpose + LayerNorm (tcgen05 path) vs
F.linear(F.layer_norm(...)) -- the reference's own ops, avsl/modules/av_hubert_encoder.py:315-334."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import avsl_b200 as A
from avsl_b200 import synth

pytestmark = pytest.mark.gpu


def _case(B, C, T, D, dtype, seed=0, mask=None, offset=0.0):
    g = torch.Generator().manual_seed(seed)
    fa = A.alloc_features(B, C, T, dtype, "cuda")
    fv = A.alloc_features(B, C, T, dtype, "cuda")
    fa.copy_((torch.randn(B, C, T, generator=g) * 1.5 + offset).to(dtype))
    fv.copy_((torch.randn(B, C, T, generator=g) * 0.7 - 0.3).to(dtype))
    W = (torch.randn(D, 2 * C, generator=g) / (2 * C) ** 0.5).cuda()
    bias = torch.randn(D, generator=g).cuda() * 0.1
    gamma = (torch.rand(2 * C, generator=g) + 0.5).cuda()
    beta = (torch.randn(2 * C, generator=g) * 0.2).cuda()
    return fa, fv, W, bias, gamma, beta


def _reference(fa, fv, mask, W, bias, gamma, beta):
    m = torch.ones(fa.shape[0], 2) if mask is None else torch.as_tensor(np.asarray(mask)).float()
    m = m.to(fa.device)
    x = torch.cat([fa.float() * m[:, 0].view(-1, 1, 1), fv.float() * m[:, 1].view(-1, 1, 1)], dim=1).transpose(1, 2)
    ln = F.layer_norm(x, (x.shape[-1],), gamma, beta, 1e-5)
    return F.linear(ln, W, bias), ln


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("B,C,T,D,masked", [(2, 64, 128, 256, False), (3, 128, 200, 512, True), (2, 1024, 750, 1024, True),
                                            (1, 64, 5, 256, False), (2, 192, 257, 768, True)])
def test_fuse_layernorm_project_matches_reference_ops(dtype, B, C, T, D, masked):
    fa, fv, W, bias, gamma, beta = _case(B, C, T, D, dtype, seed=B * 7 + T, offset=0.75)
    mask = None
    if masked:
        mask = np.ones((B, 2), np.uint8)
        mask[0, 1] = 0                      # first sample: video dropped
        if B > 1:
            mask[1, 0] = 0                  # second sample: audio dropped
    folded = A.FoldedProjection(W, bias, gamma, beta, dtype)
    out = A.fuse_layernorm_project(fa, fv, mask, folded)
    assert out.shape == (B, T, D) and out.dtype == dtype
    ref, ln = _reference(fa, fv, mask, W, bias, gamma, beta)
    # what the reference's own mixed-precision path (LayerNorm in fp32, Linear in `dtype`) loses
    ref_lp = F.linear(ln.to(dtype), W.to(dtype), bias.to(dtype)).float()
    e_ref = (ref_lp - ref).abs().max().item()
    e_ours = (out.float() - ref).abs().max().item()
    assert e_ours <= max(2.0 * e_ref, 4e-3 * ref.abs().max().item()), (e_ours, e_ref)
    rel = (out.float() - ref).norm() / ref.norm()
    assert rel.item() <= (2e-3 if dtype == torch.float16 else 1.2e-2)


def test_contract_errors():
    fa, fv, W, bias, gamma, beta = _case(2, 64, 100, 256, torch.float16)
    folded = A.FoldedProjection(W, bias, gamma, beta, torch.float16)
    bad = torch.randn(2, 64, 100, device="cuda").half()            # pitch 100: rows not 16-byte aligned
    with pytest.raises(ValueError, match="row pitch"):
        A.fuse_layernorm_project(bad, bad, None, folded)
    with pytest.raises(ValueError, match="folded for"):
        A.fuse_layernorm_project(fa.bfloat16(), fv.bfloat16(), None, folded)
    with pytest.raises(RuntimeError, match="no backward"):
        A.fuse_layernorm_project(fa.clone().requires_grad_(), fv, None, folded)
    # T a multiple of 8: a plain contiguous tensor qualifies
    x = torch.randn(2, 64, 104, device="cuda").half()
    assert A.fuse_layernorm_project(x, x, None, folded).shape == (2, 104, 256)


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 2e-2), (torch.bfloat16, 6e-2)])
@pytest.mark.parametrize("masked", [False, True])
def test_fused_projection_module_trains(dtype, tol, masked):
    """FusedProjection in the reference's TRAINING forward (av_hubert_encoder.py:292-334): the output
    equals the inference path, and the gradients of both feature maps, layer_norm.weight / .bias and
    post_extract_proj.weight / .bias agree with torch autograd through the reference's own ops
    (LayerNorm in float32, Linear in `dtype`)."""
    B, C, T, D = 3, 128, 50, 256
    g = torch.Generator().manual_seed(4)
    fa = (torch.randn(B, C, T, generator=g) + 0.5).to(dtype).cuda().requires_grad_()
    fv = (torch.randn(B, C, T, generator=g) - 0.25).to(dtype).cuda().requires_grad_()
    mod = A.FusedProjection(C, D, dtype=dtype, device="cuda")
    with torch.no_grad():
        mod.layer_norm.weight.copy_(torch.rand(2 * C, generator=g) + 0.5)
        mod.layer_norm.bias.copy_(torch.randn(2 * C, generator=g) * 0.2)
    mask = None
    if masked:
        mask = np.ones((B, 2), np.uint8)
        mask[0, 1] = 0
        mask[1, 0] = 0
    R = torch.randn(B, T, D, generator=g).cuda()
    out = mod(fa, fv, mask)
    assert out.grad_fn is not None and out.shape == (B, T, D) and out.dtype == dtype
    (out.float() * R).sum().backward()
    got = [fa.grad, fv.grad, mod.layer_norm.weight.grad, mod.layer_norm.bias.grad,
           mod.post_extract_proj.weight.grad, mod.post_extract_proj.bias.grad]
    # the reference's ops under torch autograd
    fa2, fv2 = fa.detach().clone().requires_grad_(), fv.detach().clone().requires_grad_()
    lw, lb = mod.layer_norm.weight.detach().clone().requires_grad_(), mod.layer_norm.bias.detach().clone().requires_grad_()
    W, pb = mod.post_extract_proj.weight.detach().clone().requires_grad_(), mod.post_extract_proj.bias.detach().clone().requires_grad_()
    m = torch.ones(B, 2) if mask is None else torch.as_tensor(mask).float()
    m = m.cuda()
    x = torch.cat([fa2.float() * m[:, 0].view(-1, 1, 1), fv2.float() * m[:, 1].view(-1, 1, 1)], dim=1).transpose(1, 2)
    ref = F.linear(F.layer_norm(x, (2 * C,), lw, lb, mod.layer_norm.eps).to(dtype), W.to(dtype), pb.to(dtype))
    (ref.float() * R).sum().backward()
    want = [fa2.grad, fv2.grad, lw.grad, lb.grad, W.grad, pb.grad]
    assert (out.float() - ref.float()).abs().max().item() <= tol * ref.float().abs().max().item()
    for name, a, b in zip(("fa", "fv", "ln.weight", "ln.bias", "proj.weight", "proj.bias"), got, want):
        assert a is not None and a.shape == b.shape, name
        scale = max(b.float().abs().max().item(), 1e-6)
        assert (a.float() - b.float()).abs().max().item() <= tol * scale, name
    # inference: same numbers without a graph, folded weights cached until a parameter changes
    mod.eval()
    with torch.no_grad():
        y1 = mod(fa.detach(), fv.detach(), mask)
        folded = mod._folded
        y2 = mod(fa.detach(), fv.detach(), mask)
        assert mod._folded is folded and torch.equal(y1, y2) and y1.grad_fn is None
        assert (y1.float() - out.float()).abs().max().item() <= 1e-3 * out.float().abs().max().item()
        mod.post_extract_proj.bias.add_(1.0)
        y3 = mod(fa.detach(), fv.detach(), mask)
        assert mod._folded is not folded and (y3.float() - y1.float() - 1.0).abs().max().item() < 5e-2
