"""One front-end batch (B=256) for profiling under ncu."""
import sys
import torch
sys.path.insert(0, '.')
from preset_gen_vae_b200 import synthetic
from preset_gen_vae_b200.utils.audio import MelSpectrogram
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
mel = MelSpectrogram(1024, 256, -120.0, 257, 22050)
x = (torch.rand(n, 88576, device='cuda') - 0.5)
for _ in range(reps):
    y = mel.compute(x)
torch.cuda.synchronize()
print(float(y.mean()))
