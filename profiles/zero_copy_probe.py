"""Probe: lip ROI with frames left in pinned host memory (kernel pulls footprints over PCIe)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from avsl_b200 import synth
from avsl_b200 import lips as LP

dev = torch.device("cuda", 0)
idx, durs = bench.rank_utterances(0, 1, 128)
T = np.maximum(1, np.round(durs * 25).astype(np.int64)); clip_off = np.concatenate([[0], np.cumsum(T)]).astype(np.int64); N = int(clip_off[-1])
frames = synth.video_frames_cuda(N, 224, 224, seed=3407, device=dev)
lms, vals = zip(*[synth.landmarks_for_clip(int(T[k]), 224, 224, seed=3407 + int(idx[k]), invalid_frac=0.05) for k in range(len(idx))])
lm = torch.from_numpy(np.concatenate(lms)).to(dev); val = torch.from_numpy(np.concatenate(vals)).to(dev); off = torch.from_numpy(clip_off).to(dev)
host = frames.cpu().pin_memory()
ref = LP.lip_roi_batch(frames, off, lm, val, want_gray=False, want_f32=True)
def timeit(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
out = LP.lip_roi_batch(host, off, lm, val, want_gray=False, want_f32=True)
torch.cuda.synchronize()
print("zero-copy equals device-resident:", torch.equal(out.lip_f32, ref.lip_f32))
print("device frames, no gray  : %.3f ms" % timeit(lambda: LP.lip_roi_batch(frames, off, lm, val, want_gray=False, want_f32=True)))
print("pinned host frames      : %.3f ms" % timeit(lambda: LP.lip_roi_batch(host, off, lm, val, want_gray=False, want_f32=True)))
dbuf = torch.empty_like(frames)
print("H2D copy of all frames  : %.3f ms (%.1f GB/s)" % ((t := timeit(lambda: dbuf.copy_(host, non_blocking=True))), host.numel() / t / 1e6))
# outputs written straight to pinned host memory by the blend warps (no D2H copy afterwards)
lip_host = torch.empty((N, 88, 88), dtype=torch.float32).pin_memory()
res_h = LP.LipBatch(None, None, lip_host, None, None, off)
def zc_both():
    LP.lip_roi_batch(host, off, lm, val, want_gray=False, want_f32=True, out=res_h)
zc_both(); torch.cuda.synchronize()
print("zero-copy in AND out equals reference:", torch.equal(lip_host, ref.lip_f32.cpu()))
print("pinned frames in, pinned lip out : %.3f ms" % timeit(zc_both))
lip_dev = torch.empty((N, 88, 88), dtype=torch.float32, device=dev)
res_d = LP.LipBatch(None, None, lip_dev, None, None, off)
def zc_then_copy():
    LP.lip_roi_batch(host, off, lm, val, want_gray=False, want_f32=True, out=res_d)
    lip_host.copy_(lip_dev, non_blocking=True)
print("pinned frames in, device lip + D2H copy (serial) : %.3f ms" % timeit(zc_then_copy))
def dev_out_host():
    LP.lip_roi_batch(frames, off, lm, val, want_gray=False, want_f32=True, out=res_h)
print("device frames in, pinned lip out : %.3f ms" % timeit(dev_out_host))
print("D2H copy of lip alone : %.3f ms" % timeit(lambda: lip_host.copy_(lip_dev, non_blocking=True)))
