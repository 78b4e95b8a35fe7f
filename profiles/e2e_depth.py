"""e2e features-only pipeline: effect of the number of slots in flight."""
import sys, os, time
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
for depth in (1, 2, 3, 4):
    pipe = A.HostPipeline(depth=depth, n_mels=80, audio_max_length=480000, device=dev, want_gray=False)
    for i in range(depth): pipe.submit(i, host)
    pipe.drain(); torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(20): pipe.submit(i, host)
    pipe.drain(); t1.record(); torch.cuda.synchronize()
    print(f"depth {depth}: {t0.elapsed_time(t1)/20:.3f} ms/step", flush=True)
    del pipe
