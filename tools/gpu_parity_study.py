"""Parity study that sets the gates of the default (TF32) precision.  TEST INFRASTRUCTURE: imports the oracle.

For B in argv (default 4 8 160) the SAME seeded inputs, weights, eps and dropout masks go through
  (R) the oracle in fp64 on the GPU (torch eager; the checker),
  (a) the oracle in fp32 on the GPU, TF32 off     (cuDNN / cuBLAS fp32: what the reference's own GPU path computes),
  (b) the oracle in fp32 on the GPU, TF32 on      (cuDNN / cuBLAS TF32: the library's idea of a TF32 training step),
  (c) this package, precision 'tf32'              (the benchmarked default),
  (d) this package, precision 'fp32'.
Printed per arm: relative loss errors, relative-L2 of every output, global gradient cosine / relative-L2, worst tensor, and
the relative-L2 per parameter group.  Also times (a) and (b) (ms per training step incl. torch Adam, inputs resident).

    gpurun --timeout 900 -- 'python tools/gpu_parity_study.py 4 8 160 > gpurun_out/parity_study.log 2>&1'
"""
import copy
import json
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from oracle import losses as oloss, model as omodel  # noqa: E402
from preset_gen_vae_b200 import config as pcfg, synthetic  # noqa: E402
from preset_gen_vae_b200.data.preset import DexedLearnableLayout  # noqa: E402
from preset_gen_vae_b200.model import build, loss as ploss, ops  # noqa: E402

import os
CPU_DRY = bool(os.environ.get('PGV_STUDY_CPU'))        # logic check of the torch arms on a box without a GPU
dev = torch.device('cpu' if CPU_DRY else 'cuda')
helper = DexedLearnableLayout().preset_indexes_helper
GROUPS = (('enc.cnn', 'ae_model.encoder.single_ch_cnn'), ('enc.mixer', 'ae_model.encoder.features_mixer_cnn'),
          ('enc.fc', 'ae_model.encoder.mlp'), ('latent flow', 'ae_model.flow_transform'), ('dec.fc', 'ae_model.decoder.mlp'),
          ('dec.unmix', 'ae_model.decoder.features_unmixer_cnn'), ('dec.cnn', 'ae_model.decoder.single_ch_cnn'),
          ('reg flow', 'reg_model'))


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def cast_noise(noise, dtype):
    out = {k: v.to(dev, dtype) for k, v in noise.items() if torch.is_tensor(v)}
    out['reg_masks'] = [[m.to(dev, dtype) for m in layer] for layer in noise['reg_masks']]
    return out


def oracle_step(model, x, v_in, info, noise, dtype, beta):
    model.train()
    model.zero_grad()
    outs, losses, total = oloss.train_step_losses(model, x.to(dtype), v_in.to(dtype), info, cast_noise(noise, dtype), beta=beta)
    total.backward()
    return outs, losses, {n: p.grad for n, p in model.named_parameters()}


def mine_step(mine, x, v_in, info, noise, precision, beta):
    ops.set_precision(precision)
    mine.train()
    mine.zero_grad()
    dn = cast_noise(noise, torch.float32)
    z0_ml, z0, zk, logdet, x_out = mine(x, info, dn)
    v_out = mine.reg_model(zk, dropout_masks=dn['reg_masks'])
    recons = ploss.MSELoss()(x_out, x)
    lat = mine.latent_loss(z0_ml, z0, zk, logdet)
    cont = ploss.SynthParamsLoss(helper, True, cat_bce=False, cat_softmax=True, cat_softmax_t=0.2)(v_out, v_in)
    (recons + beta * lat + cont).backward()
    torch.cuda.synchronize()
    ops.set_precision('tf32')
    outs = dict(z0_mu_logvar=z0_ml, z0=z0, zK=zk, logdet=logdet, x_out=x_out, v_out=v_out)
    return outs, dict(recons=recons, latent=lat, controls=cont), {n: p.grad for n, p in mine.named_parameters()}


def compare(tag, ref, got, B, table):
    (r_out, r_loss, r_g), (g_out, g_loss, g_g) = ref, got
    row = {'arm': tag, 'B': B}
    row['loss_rel'] = {k: abs(float(g_loss[k]) - float(r_loss[k])) / abs(float(r_loss[k])) for k in r_loss}
    row['out_rel'] = {k: rel(g_out[k], r_out[k]) for k in r_out}
    dot = n1 = n2 = 0.0
    worst = ('', 0.0)
    grp = {g: [0.0, 0.0] for g, _ in GROUPS}
    for name, r in r_g.items():
        g = g_g[name].double()
        r = r.double()
        d2, r2 = float(((g - r) ** 2).sum()), float((r * r).sum())
        dot += float((g * r).sum()); n1 += float((g * g).sum()); n2 += r2
        if r2 > 1e-14 and (d2 / r2) ** 0.5 > worst[1] and d2 ** 0.5 > 1e-4 * 1.0:
            worst = (name, (d2 / r2) ** 0.5)
        for gname, prefix in GROUPS:
            if name.startswith(prefix):
                grp[gname][0] += d2; grp[gname][1] += r2
    row['grad_cos'] = dot / np.sqrt(n1 * n2)
    row['grad_rel'] = float(np.sqrt(max(n1 + n2 - 2 * dot, 0.0) / n2))
    row['worst'] = worst
    row['group_rel'] = {g: (v[0] / max(v[1], 1e-300)) ** 0.5 for g, v in grp.items()}
    table.append(row)
    print("B=%3d %-22s cos %.6f rel %.3e worst %.2e (%s)\n      losses %s\n      outs %s\n      groups %s" % (
        B, tag, row['grad_cos'], row['grad_rel'], worst[1], worst[0].replace('ae_model.', ''),
        ' '.join('%s %.1e' % kv for kv in row['loss_rel'].items()), ' '.join('%s %.1e' % kv for kv in row['out_rel'].items()),
        ' '.join('%s %.1e' % kv for kv in row['group_rel'].items())), flush=True)


def time_eager(B, tf32, m_cfg, t_cfg):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    ext = omodel.build_extended_ae_model(m_cfg, t_cfg, helper)[3].to(dev).train()
    opt = torch.optim.Adam(ext.parameters(), lr=t_cfg.initial_learning_rate, weight_decay=t_cfg.weight_decay, betas=t_cfg.adam_betas)
    x = synthetic.make_spectrogram_like(B, 1, seed=0).to(dev)
    v_in = synthetic.make_preset_targets(helper, B, seed=0).to(dev)
    info = synthetic.make_sample_info(B).to(dev)

    def step():
        opt.zero_grad()
        _, _, total = oloss.train_step_losses(ext, x, v_in, info, None, beta=t_cfg.beta)
        total.backward()
        opt.step()
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 20
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print("TIMING reference modules, PyTorch eager on one GPU (spectrograms resident, no front end), B=%d, TF32 %s: %.2f ms/step = %.0f samples/s"
          % (B, 'on' if tf32 else 'off', ms, B / ms * 1e3), flush=True)
    return ms


def main():
    batches = [int(a) for a in sys.argv[1:]] or [4, 8, 160]
    table, timing = [], {}
    for B in batches:
        m_cfg, t_cfg = pcfg.make_default(minibatch_size=B)
        pcfg.apply_dataset_dims(m_cfg, helper)
        torch.manual_seed(0)
        orc = omodel.build_extended_ae_model(m_cfg, t_cfg, helper)[3]
        torch.manual_seed(0)
        mine = build.build_extended_ae_model(m_cfg, t_cfg, helper)[3]
        mine.load_state_dict(orc.state_dict())
        if not CPU_DRY:
            mine.to(dev)
        x = synthetic.make_spectrogram_like(B, 1, seed=0).to(dev)
        v_in = synthetic.make_preset_targets(helper, B, seed=0).to(dev)
        info = synthetic.make_sample_info(B).to(dev)
        noise = synthetic.make_noise(B, m_cfg.dim_z, t_cfg.fc_dropout, t_cfg.reg_fc_dropout, seed=1)
        beta = t_cfg.beta
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        orc64 = copy.deepcopy(orc).double().to(dev)
        ref = oracle_step(orc64, x, v_in, info, noise, torch.float64, beta)
        orc32 = copy.deepcopy(orc).to(dev)
        compare('torch fp32 (TF32 off)', ref, oracle_step(orc32, x, v_in, info, noise, torch.float32, beta), B, table)
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True
        orc32b = copy.deepcopy(orc).to(dev)
        compare('torch TF32 (cuDNN)', ref, oracle_step(orc32b, x, v_in, info, noise, torch.float32, beta), B, table)
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        if CPU_DRY:
            continue
        compare('pgv tf32', ref, mine_step(mine, x, v_in, info, noise, 'tf32', beta), B, table)
        compare('pgv tf32 (2nd run)', ref, mine_step(mine, x, v_in, info, noise, 'tf32', beta), B, table)
        if hasattr(ops, 'PRECISIONS'):
            for pr in ops.PRECISIONS:
                if pr not in ('tf32', 'fp32'):
                    compare('pgv ' + pr, ref, mine_step(mine, x, v_in, info, noise, pr, beta), B, table)
        if B <= 16 or '--fp32-all' in sys.argv:
            compare('pgv fp32', ref, mine_step(mine, x, v_in, info, noise, 'fp32', beta), B, table)
        del orc64, orc32, orc32b, mine, ref
        torch.cuda.empty_cache()
    Bt = max(batches)
    m_cfg, t_cfg = pcfg.make_default(minibatch_size=Bt)
    pcfg.apply_dataset_dims(m_cfg, helper)
    for tf32 in (() if CPU_DRY else (False, True)):
        timing['tf32_on' if tf32 else 'tf32_off'] = time_eager(Bt, tf32, m_cfg, t_cfg)
    print('JSON ' + json.dumps({'rows': table, 'eager_ms_per_step': timing, 'B_timed': Bt}))


if __name__ == '__main__':
    main()
