"""Timing of the TMA LayerNorm kernels at configs[3] size (L2 flushed between iterations), for
profiles/fuseln_sweep.sh (same-box A/B of CTAs per SM)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, avsl_b200 as A
from avsl_b200 import synth
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); tot += a.elapsed_time(b)
    return tot / n
fa0, fv0, mask = synth.fusion_inputs(64, 1024, 750, seed=3407, device="cuda")
for dt, name in ((torch.float32, "f32"), (torch.float16, "f16")):
    fa, fv = A.alloc_features(64, 1024, 750, dt, "cuda"), A.alloc_features(64, 1024, 750, dt, "cuda")
    fa.copy_(fa0.to(dt)); fv.copy_(fv0.to(dt))
    for mode, C in (("concat", 2048), ("add", 1024)):
        w = torch.ones(C, device="cuda"); b = torch.zeros(C, device="cuda")
        out = torch.empty((64, 750, C), dtype=dt, device="cuda")
        ms = t(lambda: A.fuse_transpose_layernorm(fa, fv, mask, mode, w, b, out=out))
        print(f"{name} {mode}: {ms:.4f} ms", flush=True)
