import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, avsl_b200 as A
dtype = torch.float16
B, C, T, D = 64, 1024, 750, 1024
g = torch.Generator(device="cuda").manual_seed(0)
fa = A.alloc_features(B, C, T, dtype, "cuda"); fv = A.alloc_features(B, C, T, dtype, "cuda")
fa.copy_(torch.randn(B, C, T, generator=g, device="cuda").to(dtype)); fv.copy_(torch.randn(B, C, T, generator=g, device="cuda").to(dtype))
W = torch.randn(D, 2 * C, generator=g, device="cuda") / (2 * C) ** 0.5
folded = A.FoldedProjection(W, None, None, None, dtype)
out = torch.empty((B, T, D), dtype=dtype, device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    A.fuse_layernorm_project(fa, fv, None, folded, out=out)
torch.cuda.synchronize()
print("ok")
