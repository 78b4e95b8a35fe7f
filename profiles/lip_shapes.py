"""Lip stage on other frame shapes than the benchmark's 224x224 (real AMI closeups are 352x288):
    python profiles/lip_shapes.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np, avsl_b200 as A
from avsl_b200 import synth
from avsl_b200.lips import lip_roi_batch
def run(H, W, n_clips=32, T=250, u8=False):
    fr = synth.video_frames_cuda(n_clips * T, H, W, device="cuda")
    lms, vs = [], []
    for c in range(n_clips):
        lm, v = synth.landmarks_for_clip(T, H, W, seed=100 + c)
        lms.append(lm); vs.append(v)
    lm = torch.from_numpy(np.concatenate(lms)).cuda(); v = torch.from_numpy(np.concatenate(vs)).cuda()
    offs = torch.arange(n_clips + 1, dtype=torch.int64, device="cuda") * T
    out = None
    def f():
        nonlocal out
        out = lip_roi_batch(fr, offs, lm, v, want_gray=True, want_u8=u8, want_f32=True, out=out)
    for _ in range(3): f()
    torch.cuda.synchronize(); s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): f()
    e.record(); torch.cuda.synchronize(); ms = s.elapsed_time(e)/10
    n = n_clips * T
    nb = n * (H*W*3 + 68*2*8 + H*W + 88*88*4 + (96*96 if u8 else 0))
    print(f"{H}x{W}{' +u8 ROI' if u8 else ''}: {ms:.4f} ms, {nb/ms/1e6:.0f} GB/s, frac {nb/ms/1e6/6552.6:.3f}")
if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "sweep":      # profiles/lip_sweep.sh: the two shapes that matter, both windows
        run(224, 224, 66, 250); run(288, 352); run(224, 224, 66, 250, u8=True)
    else:
        run(224, 224); run(288, 352); run(480, 640)
