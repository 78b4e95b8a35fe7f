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
