"""Front-end only: the fused MFCC kernel on 8192 x 1 s clips, a few launches (ncu target)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from honk2_b200 import AudioProcessor, synth  # noqa: E402

dev = torch.device("cuda", 0)
ap = AudioProcessor()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
w = torch.from_numpy(synth.broadband(n, seed=3)).to(dev)
out = torch.empty((n, 101, 40), device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    ap.compute_mfccs_batch(w, out=out)
torch.cuda.synchronize()
e0.record()
for _ in range(5):
    ap.compute_mfccs_batch(w, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"mfcc front-end: {n} clips in {ms:.3f} ms = {n * 80160 / ms / 1e6:.0f} GB/s algorithmic")
