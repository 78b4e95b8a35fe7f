import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, avsl_b200 as A
from avsl_b200 import synth
dense = synth.audio_batch(64, 480000, 3407, device="cuda")
out = torch.empty((64, 80, 3000), device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    A.log_mel_spectrogram(dense, 80, out=out)
torch.cuda.synchronize()
print("ok")
