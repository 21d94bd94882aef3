"""Tiny driver for ncu: a few launches of representative tensor-core kernels."""
import sys
import torch
sys.path.insert(0, '.')
from preset_gen_vae_b200.model import ops
ops.set_precision('tf32')
B = 160
dev = 'cuda'
a = torch.randn(B, 300, device=dev); w = torch.randn(300, 300, device=dev) * 0.05; bias = torch.randn(300, device=dev)
dy7 = torch.randn(B, 16, 65, 88, device=dev); w7 = torch.randn(16, 8, 4, 4, device=dev) * 0.1
x4 = torch.randn(B, 32, 33, 45, device=dev); w4 = torch.randn(64, 32, 4, 4, device=dev) * 0.1; b4 = torch.randn(64, device=dev)
dy4 = torch.randn(B, 64, 17, 23, device=dev)
for _ in range(3):
    ops.linear_fwd(a, w, bias)
    ops.conv2d_dgrad(dy7, w7, (129, 174), 2, 2)
    ops.conv2d_fwd(x4, w4, b4, 2, 2, 0.1)
    ops.conv2d_wgrad(x4, dy4, w4.shape, 2, 2, want_bias=False)
torch.cuda.synchronize()
