"""Target for an ncu capture of the stand-alone column-slice kernels of the latent flow (program kernel off)."""
import sys

import torch

sys.path.insert(0, '.')
from preset_gen_vae_b200.model import flows, ops

ops.use_flow_program = 'program' in sys.argv
torch.manual_seed(0)
flow = flows.SimpleRealNVP(610, 300, 6, 2, batch_norm_within_layers=True)._transform.cuda().train()
x = torch.randn(160, 610, device='cuda').requires_grad_()
for it in range(3):
    flow.zero_grad()
    y, ld = flow(x)
    (y.sum() + ld.sum()).backward()
torch.cuda.synchronize()
