"""ncu driver: the output-heavy (8/16/32-channel) launches of the channels-last conv kernel at B=160."""
import sys
import torch
sys.path.insert(0, '.')
from preset_gen_vae_b200.model import ops
ops.set_precision('tf32')
B = 160
dev = 'cuda'
x2 = ops.to_cl(torch.randn(B, 8, 129, 174, device=dev), True); w2 = torch.randn(16, 8, 4, 4, device=dev) * 0.1; b2 = torch.randn(16, device=dev)
dy2 = ops.to_cl(torch.randn(B, 16, 65, 88, device=dev), True); bt = torch.randn(8, device=dev)
x3 = ops.to_cl(torch.randn(B, 16, 65, 88, device=dev), True); w3 = torch.randn(32, 16, 4, 4, device=dev) * 0.1; b3 = torch.randn(32, device=dev)
wf2, wq2 = ops.prep_conv_weights(w2, 2, 2)
wf3, wq3 = ops.prep_conv_weights(w3, 2, 2)
for _ in range(2):
    ops.conv2d_fwd(x2, w2, b2, 2, 2, 0.1, wf=wf2)                                   # enc2 forward: N=16, direct epilogue
    ops.conv2d_dgrad(dy2, w2, (129, 174), 2, 2, bias=bt, slope=0.1, wq=wq2)         # dec7 forward: quad epilogue, N=32
    ops.conv2d_fwd(x3, w3, b3, 2, 2, 0.1, wf=wf3)                                   # enc3 forward: N=32, staged epilogue
torch.cuda.synchronize()
