"""e2e features-only pipeline: lip features stored across PCIe by the kernel itself (direct) against a
device buffer drained by the copy engine, at several pipeline depths; and the lip kernel alone."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import avsl_b200 as A
from avsl_b200 import synth
from avsl_b200.frontend import PackedBatch
dev = torch.device("cuda", 0)
idx, durs = bench.rank_utterances(0, 1, 128)
T = np.maximum(1, np.round(durs * 25).astype(np.int64)); clip_off = np.concatenate([[0], np.cumsum(T)]).astype(np.int64); N = int(clip_off[-1])
a_len = np.round(durs * 16000).astype(np.int64); a_off = np.concatenate([[0], np.cumsum(a_len)]).astype(np.int64)
g = torch.Generator(device=dev).manual_seed(3407)
audio = (torch.randn(int(a_off[-1]), generator=g, device=dev) * 0.1).clamp_(-1, 1)
frames = synth.video_frames_cuda(N, 224, 224, seed=3407, device=dev)
lms, vals = zip(*[synth.landmarks_for_clip(int(T[k]), 224, 224, seed=3407 + int(idx[k]), invalid_frac=0.05) for k in range(len(idx))])
host = PackedBatch(audio, torch.from_numpy(a_off).to(dev), frames, torch.from_numpy(clip_off).to(dev), torch.from_numpy(np.concatenate(lms)).to(dev), torch.from_numpy(np.concatenate(vals)).to(dev)).pin()
del frames
for direct in (True, False):
    for depth in (2, 3):
        pipe = A.HostPipeline(depth=depth, n_mels=80, audio_max_length=480000, device=dev, want_gray=False)
        for fe in pipe.fes: fe.lip_direct = direct
        for i in range(depth): pipe.submit(i, host)
        pipe.drain(); torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(20): pipe.submit(i, host)
        pipe.drain(); t1.record(); torch.cuda.synchronize()
        print(f"lip_direct={direct} depth {depth}: {t0.elapsed_time(t1)/20:.3f} ms/step", flush=True)
        del pipe
# the lip kernel alone, frames in mapped pinned memory: output to device memory / to pinned memory
from avsl_b200 import lips as L
devb = host.to(dev, non_blocking=False, frames_stay_on_host=True)
for where in ("device", "pinned"):
    out = torch.empty((N, 88, 88), dtype=torch.float32, device=dev) if where == "device" else torch.empty((N, 88, 88), dtype=torch.float32).pin_memory()
    fe = A.AVFrontEnd(n_mels=80, audio_max_length=480000, device=dev, want_gray=False)
    for _ in range(3): fe.forward_device(devb, reuse=True, lip_out=out)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(10): fe.forward_device(devb, reuse=True, lip_out=out)
    t1.record(); torch.cuda.synchronize()
    print(f"forward_device zero-copy frames, lip -> {where}: {t0.elapsed_time(t1)/10:.3f} ms", flush=True)
# D2H of the result alone
src = torch.empty((N, 88, 88), dtype=torch.float32, device=dev); dst = torch.empty((N, 88, 88), dtype=torch.float32).pin_memory()
for _ in range(2): dst.copy_(src, non_blocking=True)
torch.cuda.synchronize(); t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); t0.record()
for _ in range(10): dst.copy_(src, non_blocking=True)
t1.record(); torch.cuda.synchronize()
print(f"copy engine D2H of the lip features ({src.numel()*4/1e6:.0f} MB): {t0.elapsed_time(t1)/10:.3f} ms = {src.numel()*4/1e6/(t0.elapsed_time(t1)/10):.1f} GB/s", flush=True)
