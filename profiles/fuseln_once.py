import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, avsl_b200 as A
from avsl_b200 import synth
dt = {"f32": torch.float32, "f16": torch.float16}[sys.argv[1] if len(sys.argv) > 1 else "f32"]
fa0, fv0, mask = synth.fusion_inputs(64, 1024, 750, seed=3407, device="cuda")
fa, fv = A.alloc_features(64, 1024, 750, dt, "cuda"), A.alloc_features(64, 1024, 750, dt, "cuda")
fa.copy_(fa0.to(dt)); fv.copy_(fv0.to(dt))
w = torch.ones(2048, device="cuda"); b = torch.zeros(2048, device="cuda")
out = torch.empty((64, 750, 2048), dtype=dt, device="cuda")
for _ in range(3):
    A.fuse_transpose_layernorm(fa, fv, mask, "concat", w, b, out=out)
torch.cuda.synchronize(); print("ok")
