"""CPU: the oracle's monitoring metrics / inference tail / FlowParamsLoss restatements against the REFERENCE's numbers committed in
tests/golden/metrics.npz (written by oracle/make_golden.py from model/loss.py:187-346 and data/preset.py:341-369, unmodified)."""
import os

import numpy as np
import torch

from oracle import losses as oloss, model as omodel
from preset_gen_vae_b200 import config as pcfg
from preset_gen_vae_b200.data import preset as ppreset


def _g(golden_dir):
    return np.load(os.path.join(golden_dir, 'metrics.npz'))


def test_monitoring_metrics_match_the_reference(golden_dir, idx_helper):
    g = _g(golden_dir)
    v_in, v_out = torch.from_numpy(g['v_in']), torch.from_numpy(g['v_out'])
    assert abs(oloss.quantized_numerical_params_loss(idx_helper, v_out, v_in).item() - float(g['qloss_mse'])) < 1e-7
    assert abs(oloss.quantized_numerical_params_loss(idx_helper, v_out, v_in, l1=True).item() - float(g['qloss_l1'])) < 1e-7
    assert abs(oloss.categorical_params_accuracy(idx_helper, v_out, v_in) - float(g['accuracy_pct'])) < 1e-9
    per = oloss.categorical_params_accuracy(idx_helper, v_out, v_in, reduce=False, percentage_output=False)
    assert list(per.keys()) == g['acc_keys'].tolist() and np.allclose(list(per.values()), g['acc_vals'], atol=1e-12)
    full = oloss.learnable_to_full(idx_helper, v_out, ppreset.DexedLearnableLayout().params_default_values)
    assert torch.equal(full, torch.from_numpy(g['full_presets']))
    # the host-side product path of the same conversion (CPU tensors)
    assert torch.equal(ppreset.learnable_to_full_presets(idx_helper, v_out, ppreset.DexedLearnableLayout().params_default_values), full)


def test_dkl_and_l2_match_the_reference(golden_dir):
    g = _g(golden_dir)
    ml = torch.from_numpy(g['mu_logvar'])
    assert abs(omodel.gaussian_dkl(ml[:, 0], ml[:, 1], True).item() - float(g['dkl'])) < 1e-6
    assert abs(omodel.gaussian_dkl(ml[:, 0], ml[:, 1], False).item() - float(g['dkl_raw'])) < 1e-3
    a, b = torch.from_numpy(g['l2_a']), torch.from_numpy(g['l2_b'])
    got = [oloss.l2_loss(a, b, c, ba).item() for c in (False, True) for ba in (False, True)]
    assert np.allclose(got, g['l2'], rtol=1e-6)


def test_flow_params_loss_matches_the_reference(golden_dir, idx_helper):
    """model/loss.py:318-346 through the oracle model built with forward_controls_loss=False (regression.py:179-184)."""
    g = _g(golden_dir)
    m_cfg, t_cfg = pcfg.make_default(minibatch_size=6, forward_controls_loss=False)
    pcfg.apply_dataset_dims(m_cfg, idx_helper)
    torch.manual_seed(0)
    ext = omodel.build_extended_ae_model(m_cfg, t_cfg, idx_helper)[3].train()
    for blk in [m for m in ext.modules() if type(m).__name__ == 'ResidualBlock']:
        blk.dropout.p = 0.0
    ml = torch.from_numpy(g['mu_logvar'][:6]).clone().requires_grad_()
    v = torch.from_numpy(g['flow_params_v'])
    loss = oloss.flow_params_loss(ext.ae_model.flow_transform.inverse, ext.reg_model.flow_inverse_function, ml, v)
    loss.backward()
    assert abs(loss.item() - float(g['flow_params_loss'])) < 1e-6
    assert np.allclose(ml.grad.numpy(), g['flow_params_dml'], rtol=1e-4, atol=1e-9)
    params = dict(ext.named_parameters())
    names = g['flow_params_grad_names'].tolist()
    norms = np.asarray([float(params[n].grad.double().norm()) for n in names])
    assert np.allclose(norms, g['flow_params_grad_norms'], rtol=1e-4, atol=1e-12)
    assert all(p.grad is None for n, p in params.items() if n not in names)            # encoder / decoder are not on this loss' path


def test_spectrogram_statistics_rule():
    """abstractbasedataset.py:357-360: data-set min / max / mean of means / sqrt(mean of variances)."""
    g = torch.Generator().manual_seed(3)
    specs = [torch.randn(257, 347, generator=g) * (i + 1) - 60.0 for i in range(5)]
    per, ds = oloss.spectrogram_stats(specs)
    assert per.shape == (5, 4)
    allv = torch.stack(specs)
    assert ds['min'] == allv.min().item() and ds['max'] == allv.max().item()
    assert abs(ds['mean'] - allv.mean().item()) < 1e-4
    assert abs(ds['std'] - np.sqrt(np.mean([s.var().item() for s in specs]))) < 1e-9
