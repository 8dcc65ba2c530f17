"""Small forward passes of the kernels added in round 2 (fp32 resident-weight convolution kernels: row tile and column tile,
every dilation; int16 PCM front-end incl. odd clip lengths), meant for `compute-sanitizer --tool memcheck|racecheck python
tools/sanitize_small.py`.  compute-sanitizer is closed on this GPU pool, so the same cases are covered by the bit-identity
tests in tests/test_gpu_parity.py (three convolution kernels against each other and the oracle; PCM16 against float32)."""
import numpy as np
import torch

import honk2_b200
from honk2_b200 import AudioProcessor, synth

dev = torch.device("cuda", 0)
fe = AudioProcessor()
for n in (16000, 15999, 4007):
    pcm = np.random.default_rng(n).integers(-32768, 32768, size=(5, n), dtype=np.int16)
    a = fe.compute_mfccs_batch(torch.from_numpy(pcm).to(dev))
    b = fe.compute_mfccs_batch(torch.from_numpy(pcm.astype(np.float32) / 32768.0).to(dev))
    assert torch.equal(a, b)
waves = torch.from_numpy(synth.speechlike(5, seed=3)).to(dev)
for name in ("res15", "res8", "res26", "res15_narrow", "res26_narrow"):
    m = honk2_b200.build_model(name, precision="fp32").to(dev)
    with torch.no_grad():
        y = m.forward_wave(waves, fe)
    assert torch.isfinite(y).all()
    print(name, "ok", float(y.abs().max()))
torch.cuda.synchronize()
print("done")
