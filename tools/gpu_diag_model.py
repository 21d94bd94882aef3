"""Per-tensor gradient accuracy of the CUDA model path and of the fp32 CPU oracle, both against an fp64 oracle."""
import copy
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from oracle import losses as oloss, model as omodel  # noqa: E402
from preset_gen_vae_b200 import config as pcfg, synthetic  # noqa: E402
from preset_gen_vae_b200.data.preset import DexedLearnableLayout  # noqa: E402
from preset_gen_vae_b200.model import build, loss as ploss, ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
MODE = sys.argv[2] if len(sys.argv) > 2 else 'c1'           # c1 | c6 | c1fe (real front-end input)
SIX = ((40, 85), (50, 85), (60, 42), (60, 85), (60, 127), (70, 85))
h = DexedLearnableLayout().preset_indexes_helper
m_cfg, t_cfg = pcfg.make_default(minibatch_size=B, midi_notes=SIX if MODE == 'c6' else None, stack_spectrograms=(MODE == 'c6'))
print("##### B=%d mode=%s" % (B, MODE))
pcfg.apply_dataset_dims(m_cfg, h)
torch.manual_seed(0)
orc = omodel.build_extended_ae_model(m_cfg, t_cfg, h)[3]
orc64 = copy.deepcopy(orc).double()
torch.manual_seed(0)
mine = build.build_extended_ae_model(m_cfg, t_cfg, h)[3]
mine.load_state_dict(orc.state_dict())
mine.cuda()
C = m_cfg.input_tensor_size[1]
x = synthetic.make_spectrogram_like(B, C, seed=0)
if MODE == 'c1fe':
    from oracle import frontend as ofe
    x = ofe.batch_front_end(synthetic.make_audio(B, 1, seed=0), 1024, 256, -120.0, 257, -120.0, 0.0)
v_in = synthetic.make_preset_targets(h, B, seed=0)
info = synthetic.make_sample_info(B)
noise = synthetic.make_noise(B, 610, t_cfg.fc_dropout, t_cfg.reg_fc_dropout, seed=1, enc_fc_in=orc.ae_model.encoder.mlp[1].in_features)
noise64 = {k: (v.double() if torch.is_tensor(v) else [[m.double() for m in l] for l in v]) for k, v in noise.items()}


def oracle_step(model, x_, v_, n_):
    model.train()
    outs, losses, total = oloss.train_step_losses(model, x_, v_, info, n_, beta=0.2)
    total.backward()
    return outs, losses


o32, l32 = oracle_step(orc, x, v_in, noise)
o64, l64 = oracle_step(orc64, x.double(), v_in.double(), noise64)
ref = {n: p.grad for n, p in orc64.named_parameters()}


def report(tag, named_grads, outs, losses):
    rows = []
    dot = n1 = n2 = 0.0
    for n, g in named_grads:
        r = ref[n]
        g = g.double().cpu()
        dot += float((g * r).sum()); n1 += float((g * g).sum()); n2 += float((r * r).sum())
        if float(r.norm()) > 1e-7:
            rows.append((float((g - r).norm() / r.norm()), n, float(r.norm())))
    rows.sort(reverse=True)
    print("== %s: global rel-L2 %.3e cosine %.7f | median per-tensor %.2e" %
          (tag, np.sqrt(max(n1 + n2 - 2 * dot, 0) / n2), dot / np.sqrt(n1 * n2), np.median([r[0] for r in rows])))
    for e, n, nr in rows[:8]:
        print("   %.3e  %s  (|g|=%.2e)" % (e, n, nr))
    for k in o64:
        a, b = outs[k].detach().double().cpu(), o64[k].detach()
        print("   out %-13s rel %.2e" % (k, float((a - b).norm() / b.norm())), end='')
    print()
    print("   losses", {k: float(v) for k, v in losses.items()}, "fp64", {k: float(v) for k, v in l64.items()})


report("fp32 CPU oracle vs fp64", [(n, p.grad) for n, p in orc.named_parameters()], o32, l32)
dn = {k: v.cuda() for k, v in noise.items() if torch.is_tensor(v)}
masks = [[m.cuda() for m in l] for l in noise['reg_masks']]
for prec in ('fp32', 'tf32'):
    ops.set_precision(prec)
    mine.zero_grad(set_to_none=True)
    mine.train()
    z0_ml, z0, zk, logdet, x_out = mine(x.cuda(), info.cuda(), dn)
    v_out = mine.reg_model(zk, dropout_masks=masks)
    recons = ploss.MSELoss()(x_out, x.cuda())
    lat = mine.latent_loss(z0_ml, z0, zk, logdet)
    cont = ploss.SynthParamsLoss(h, True, cat_bce=False, cat_softmax=True, cat_softmax_t=0.2)(v_out, v_in.cuda())
    (recons + 0.2 * lat + cont).backward()
    torch.cuda.synchronize()
    report("CUDA path (%s) vs fp64" % prec, [(n, p.grad) for n, p in mine.named_parameters()],
           dict(z0_mu_logvar=z0_ml, z0=z0, zK=zk, logdet=logdet, x_out=x_out, v_out=v_out),
           dict(recons=recons, latent=lat, controls=cont))
