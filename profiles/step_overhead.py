"""Host-side cost of one eager AVFrontEnd step vs its GPU time (is the eager loop launch-bound?)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from avsl_b200 import synth
from avsl_b200.frontend import AVFrontEnd, PackedBatch

dev = torch.device("cuda", 0)
idx, durs = bench.rank_utterances(0, 1, 128)
T = np.maximum(1, np.round(durs * 25).astype(np.int64)); clip_off = np.concatenate([[0], np.cumsum(T)]).astype(np.int64); N = int(clip_off[-1])
a_len = np.round(durs * 16000).astype(np.int64); a_off = np.concatenate([[0], np.cumsum(a_len)]).astype(np.int64)
g = torch.Generator(device=dev).manual_seed(3407)
audio = (torch.randn(int(a_off[-1]), generator=g, device=dev) * 0.1).clamp_(-1, 1)
frames = synth.video_frames_cuda(N, 224, 224, seed=3407, device=dev)
lms, vals = zip(*[synth.landmarks_for_clip(int(T[k]), 224, 224, seed=3407 + int(idx[k]), invalid_frac=0.05) for k in range(len(idx))])
b = PackedBatch(audio, torch.from_numpy(a_off).to(dev), frames, torch.from_numpy(clip_off).to(dev), torch.from_numpy(np.concatenate(lms)).to(dev), torch.from_numpy(np.concatenate(vals)).to(dev))
fe = AVFrontEnd(device=dev)
for reuse in (True, False):
    for _ in range(5): fe.forward_device(b, reuse=reuse)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): fe.forward_device(b, reuse=reuse)
    e1.record()
    t_cpu = (time.perf_counter() - t0) / 50
    torch.cuda.synchronize()
    print(f"reuse={reuse}: host enqueue {t_cpu*1e3:.3f} ms/step, GPU {e0.elapsed_time(e1)/50:.3f} ms/step", flush=True)
# per-call host cost, GPU idle
from avsl_b200 import audio as AU, lips as LP
torch.cuda.synchronize()
mel = torch.empty((128, 80, 3000), device=dev)
for name, fn in (("logmel_ragged", lambda: AU.log_mel_spectrogram_ragged(b.audio, b.audio_offsets, 480000, 80, filters=fe.filters, out=mel)),
                 ("lip_roi_batch", lambda: LP.lip_roi_batch(b.frames, b.clip_offsets, b.landmarks, b.lm_valid))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0); torch.cuda.synchronize()
    print(f"{name}: host time per call median {np.median(ts)*1e6:.0f} us, max {max(ts)*1e6:.0f} us", flush=True)
