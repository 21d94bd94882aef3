"""Tiny driver for ncu: a few launches of the channels-last tensor-core conv kernel (fwd, quad dgrad, wgrad) at B=160."""
import sys
import torch
sys.path.insert(0, '.')
from preset_gen_vae_b200.model import ops
ops.set_precision('tf32')
B = 160
dev = 'cuda'
x5 = ops.to_cl(torch.randn(B, 64, 17, 23, device=dev), True); w5 = torch.randn(128, 64, 4, 4, device=dev) * 0.1; b5 = torch.randn(128, device=dev)
dy5 = ops.to_cl(torch.randn(B, 128, 9, 12, device=dev), True)
x2 = ops.to_cl(torch.randn(B, 8, 129, 174, device=dev), True); w2 = torch.randn(16, 8, 4, 4, device=dev) * 0.1; b2 = torch.randn(16, device=dev)
wf5, wq5 = ops.prep_conv_weights(w5, 2, 2)
wf2, wq2 = ops.prep_conv_weights(w2, 2, 2)
for _ in range(2):
    ops.conv2d_fwd(x5, w5, b5, 2, 2, 0.1, wf=wf5)
    ops.conv2d_dgrad(dy5, w5, (17, 23), 2, 2, wq=wq5)
    ops.conv2d_wgrad(x5, dy5, w5.shape, 2, 2, want_bias=False)
    ops.conv2d_fwd(x2, w2, b2, 2, 2, 0.1, wf=wf2)
torch.cuda.synchronize()
