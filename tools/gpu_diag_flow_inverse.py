"""Round trip of the regression flow (BatchNorm transforms between couplings) at B = 160: this repo vs the fp64 / fp32 oracle."""
import sys

import torch

sys.path.insert(0, '.')
from oracle import model as omodel
from preset_gen_vae_b200.model import flows, ops

torch.manual_seed(3)
flow = flows.CustomRealNVP(610, 300, 6, 2, batch_norm_within_layers=True, batch_norm_between_layers=True).cuda()
g = torch.Generator(device='cuda').manual_seed(5)
z = torch.randn(160, 610, device='cuda', generator=g)
flow.train()
with torch.no_grad():
    for _ in range(2):
        flow(z)
flow.eval()
orc = omodel.CustomRealNVP(610, 300, 6, 2, batch_norm_within_layers=True, batch_norm_between_layers=True)
orc.load_state_dict(flow.state_dict())
orc.eval()
with torch.no_grad():
    for prog in (True, False):
        ops.use_flow_program = prog
        y, ld = flow(z)
        back, ldb = flow.inverse(y)
        print("program kernel %s: pgv round trip max err %.3e, mean %.3e" % (prog, float((back - z).abs().max()), float((back - z).abs().mean())))
    for dt in (torch.float64, torch.float32):
        o = orc.to(dt)
        zc = z.cpu().to(dt)
        yo, ldo = o(zc)
        bo, _ = o.inverse(yo)
        print("oracle %s: round trip max err %.3e mean %.3e; pgv forward vs oracle max %.3e; pgv inverse(oracle y) vs z max %.3e" %
              (dt, float((bo - zc).abs().max()), float((bo - zc).abs().mean()), float((y.cpu().to(dt) - yo).abs().max()),
               float((flow.inverse(yo.float().cuda())[0].cpu().to(dt) - zc).abs().max())))
