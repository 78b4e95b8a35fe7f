"""post_extract_proj: fused tcgen05 path vs fuse_ln_kernel + cuBLAS (torch F.linear), B 64 x T 750, C 1024 -> D 1024."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
import avsl_b200 as A
from avsl_b200 import synth

def timeit(fn, n=20, flush=None):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(n):
        if flush is not None: flush.zero_()
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / n

def run(dtype, B=64, C=1024, T=750, D=1024, masked=False):
    g = torch.Generator(device="cuda").manual_seed(0)
    fa = A.alloc_features(B, C, T, dtype, "cuda"); fv = A.alloc_features(B, C, T, dtype, "cuda")
    fa.copy_(torch.randn(B, C, T, generator=g, device="cuda").to(dtype)); fv.copy_(torch.randn(B, C, T, generator=g, device="cuda").to(dtype))
    W = (torch.randn(D, 2 * C, generator=g, device="cuda") / (2 * C) ** 0.5); bias = torch.randn(D, generator=g, device="cuda") * 0.1
    gamma = torch.rand(2 * C, generator=g, device="cuda") + 0.5; beta = torch.randn(2 * C, generator=g, device="cuda") * 0.2
    mask = None
    if masked:
        _, _, mask = synth.fusion_inputs(B, 2, 2, seed=3407)
    folded = A.FoldedProjection(W, bias, gamma, beta, dtype)
    out = torch.empty((B, T, D), dtype=dtype, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    flops = 2.0 * B * T * 2 * C * D
    res = {}
    ms = timeit(lambda: A.fuse_layernorm_project(fa, fv, mask, folded, out=out), flush=flush)
    res["fused_tcgen05 (stats + gemm)"] = ms
    fac, fvc = fa.contiguous(), fv.contiguous()
    Wl, bl = W.to(dtype), bias.to(dtype)
    ln_out = torch.empty((B, T, 2 * C), dtype=dtype, device="cuda")
    def unfused():
        A.fuse_transpose_layernorm(fac, fvc, mask, "concat", gamma, beta, out=ln_out)
        return F.linear(ln_out, Wl, bl)
    res["fuse_ln_kernel + cuBLAS"] = timeit(unfused, flush=flush)
    res["cuBLAS F.linear alone"] = timeit(lambda: F.linear(ln_out, Wl, bl), flush=flush)
    def torch_only():
        x = torch.cat([fac, fvc], 1).transpose(1, 2)
        return F.linear(F.layer_norm(x.float(), (2 * C,), gamma, beta).to(dtype), Wl, bl)
    if not masked:
        res["torch ops (cat, transpose, LN fp32, linear)"] = timeit(torch_only, n=5, flush=flush)
    print(f"--- {dtype}, B {B} C {C} T {T} D {D}, masked={masked}")
    for k, v in res.items():
        print(f"{k:48s} {v:8.4f} ms  {flops / v / 1e9:8.1f} TFLOP/s (dense-equivalent)", flush=True)
    return res

if __name__ == "__main__":
    for dt in (torch.float16, torch.bfloat16):
        run(dt)
    run(torch.float16, masked=True)
