"""Per-op timing inside the flow program kernel (CTA 0): body vs grid barrier, for the latent flow forward and backward at B = 160."""
import sys

import torch

sys.path.insert(0, '.')
from preset_gen_vae_b200 import _lib
from preset_gen_vae_b200.model import flows, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 160
KIND = ['gather', 'cs_fwd', 'cs_bn_fwd', 'cs_dgrad', 'cs_bn_bwd', 'coupling_fwd', 'coupling_bwd', 'scatter', 'wgrad']
torch.manual_seed(0)
flow = flows.SimpleRealNVP(610, 300, 6, 2, batch_norm_within_layers=True)._transform.cuda().train()
x = torch.randn(B, 610, device='cuda').requires_grad_()
L = _lib.lib()
for it in range(3):
    flow.zero_grad()
    y, ld = flow(x)
    (y.sum() + ld.sum()).backward()
torch.cuda.synchronize()
tr = torch.zeros(8 * 150, dtype=torch.int64, device='cuda')
L.pgv_debug_set_flow_trace(_lib.ptr(tr))
y, ld = flow(x)
torch.cuda.synchronize()
fwd = tr.clone().view(-1, 4).cpu()
inner = fwd[150:]
fwd = fwd[:150]
tr.zero_()
(y.sum() + ld.sum()).backward()
torch.cuda.synchronize()
bwd = tr.view(-1, 4).cpu()[:150]
L.pgv_debug_set_flow_trace(None)
for name, t in (('forward', fwd), ('backward', bwd)):
    rows = [r for r in t.tolist() if r[0] > 0]
    t0 = rows[0][0]
    print("==== latent flow %s: %d ops, %.1f us in the kernel" % (name, len(rows), (max(r[2] or r[1] for r in rows) - t0) / 1e3))
    agg = {}
    for i, (a, b, c, k) in enumerate(rows):
        body, bar = (b - a) / 1e3, ((c - b) / 1e3 if c else 0.0)
        g = agg.setdefault(KIND[k], [0, 0.0, 0.0]); g[0] += 1; g[1] += body; g[2] += bar
        if i < 20:
            print("  op %3d %-12s start %8.1f  body %6.1f  barrier %6.1f" % (i, KIND[k], (a - t0) / 1e3, body, bar))
    for k, (n, body, bar) in agg.items():
        print("  %-12s x%3d   body %7.1f us (%.1f each)   barrier %7.1f us (%.1f each)" % (k, n, body, body / n, bar, bar / n))

print("==== inside the column-slice tile (forward, thread 0 of CTA 0): K loop / BatchNorm cluster reduction / rest")
for i in range(1, 8):
    a, b, c, g0 = inner[i].tolist()
    if a:
        end = fwd[i][1].item()
        print("  op %d  first chunk group landed after %.1f us" % (i, (g0 - a) / 1e3), end='')
        print("  op %d  k-loop %.1f us  reduce %.1f us  rest %.1f us" % (i, (b - a) / 1e3, ((c - b) / 1e3) if c else 0.0, (end - (c or b)) / 1e3))

