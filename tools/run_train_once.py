"""A few eager training steps (no CUDA graph) for ncu launch lists."""
import sys
import torch
sys.path.insert(0, '.')
from preset_gen_vae_b200 import config as pcfg, synthetic
from preset_gen_vae_b200.data.preset import DexedLearnableLayout
from preset_gen_vae_b200.train import TrainStep
B = int(sys.argv[1]) if len(sys.argv) > 1 else 160
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
h = DexedLearnableLayout().preset_indexes_helper
m, t = pcfg.make_default(minibatch_size=B); pcfg.apply_dataset_dims(m, h)
tr = TrainStep(m, t, h, use_cuda_graph=False)
audio = (torch.rand(B, 1, 88576, device='cuda') - 0.5); v = synthetic.make_preset_targets(h, B).cuda(); info = synthetic.make_sample_info(B).cuda()
for _ in range(steps):
    out = tr.step(audio, v, info)
torch.cuda.synchronize()
print(out.tolist())
