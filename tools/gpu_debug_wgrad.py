"""Decodes what the MN-major weight-gradient MMA actually computes: one-hot operands identify (m, n, k) positions."""
import os
import sys

import torch

sys.path.insert(0, '.')
from preset_gen_vae_b200.model import ops  # noqa: E402

torch.set_printoptions(linewidth=220, precision=1, sci_mode=False)
variant = os.environ.get('PGV_WGRAD_VARIANT', '0')
Cin = Cout = 32
B = 32
for b0 in (0, 1, 5, 9, 31):
    x = torch.zeros(B, Cin, 1, 1, device='cuda'); dy = torch.zeros(B, Cout, 1, 1, device='cuda')
    x[b0, :, 0, 0] = torch.arange(1, Cin + 1, device='cuda').float()
    dy[b0, :, 0, 0] = torch.arange(1, Cout + 1, device='cuda').float() * 100
    dw, _ = ops.conv2d_wgrad(x, dy, (Cout, Cin, 1, 1), 1, 0, want_bias=False)
    want = torch.einsum('bn,bm->nm', dy.view(B, Cout), x.view(B, Cin))
    ok = torch.allclose(dw.view(Cout, Cin), want)
    print('variant', variant, 'b0', b0, 'match', ok, 'nonzero', int((dw != 0).sum()), 'sum', float(dw.sum()), 'want sum', float(want.sum()))
    if not ok and b0 in (0, 5):
        print(dw.view(Cout, Cin)[:6, :12])
g = torch.Generator(device='cuda').manual_seed(0)
for (cin, cout, k, H, W, Bn) in [(32, 32, 1, 1, 1, 32), (64, 128, 4, 17, 23, 3), (8, 16, 4, 129, 174, 2), (512, 2048, 1, 3, 4, 5)]:
    s, p = (2, 2) if k == 4 else (1, 0)
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    x = torch.randn(Bn, cin, H, W, device='cuda', generator=g); dy = torch.randn(Bn, cout, Ho, Wo, device='cuda', generator=g)
    dw, _ = ops.conv2d_wgrad(x, dy, (cout, cin, k, k), s, p, want_bias=False)
    xd = x.double(); wd = torch.zeros(cout, cin, k, k, device='cuda', dtype=torch.float64, requires_grad=True)
    out = torch.nn.functional.conv2d(xd, wd, None, s, p)
    gw, = torch.autograd.grad(out, wd, dy.double())
    print('variant', variant, (cin, cout, k, H, W, Bn), 'rel', float((dw.double() - gw).norm() / gw.norm()), 'norm ratio', float(dw.norm() / gw.norm()))
