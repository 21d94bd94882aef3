"""GPU: monitoring metrics, inference tail, FlowParamsLoss (with autograd through the inverse latent flow), NaN flags and the
spectrogram statistics pass through the C ABI, against the reference's committed numbers
(tests/golden/metrics.npz) and the oracle.  Integer / index results (accuracies, argmax conversions) must be exact."""
import os

import numpy as np
import pytest
import torch

from oracle import losses as oloss, model as omodel
from preset_gen_vae_b200 import config as pcfg
from preset_gen_vae_b200.data import preset as ppreset
from preset_gen_vae_b200.model import build, loss as ploss, ops

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _g(golden_dir):
    return np.load(os.path.join(golden_dir, 'metrics.npz'))


def test_monitoring_metrics_match_the_reference(golden_dir, idx_helper):
    g = _g(golden_dir)
    v_in, v_out = torch.from_numpy(g['v_in']).to(DEV), torch.from_numpy(g['v_out']).to(DEV)
    q = ploss.QuantizedNumericalParamsLoss(idx_helper, numerical_loss=torch.nn.MSELoss(reduction='mean'))(v_out, v_in)
    assert q.is_cuda and abs(q.item() - float(g['qloss_mse'])) < 1e-6 * float(g['qloss_mse'])
    q1 = ploss.QuantizedNumericalParamsLoss(idx_helper, numerical_loss=torch.nn.L1Loss())(v_out, v_in)
    assert abs(q1.item() - float(g['qloss_l1'])) < 1e-6 * float(g['qloss_l1'])
    acc = ploss.CategoricalParamsAccuracy(idx_helper, reduce=True, percentage_output=True)(v_out, v_in)
    assert acc.is_cuda and abs(acc.item() - float(g['accuracy_pct'])) < 1e-4
    per = ploss.CategoricalParamsAccuracy(idx_helper, reduce=False, percentage_output=False)(v_out, v_in)
    assert list(per.keys()) == g['acc_keys'].tolist()
    assert np.array_equal(np.round(np.asarray(list(per.values())) * 48), np.round(g['acc_vals'] * 48))     # exact match counts out of B = 48
    lim = g['limited'].tolist()
    ql = ploss.QuantizedNumericalParamsLoss(idx_helper, limited_vst_params_indexes=lim)(v_out, v_in)
    assert abs(ql.item() - float(g['qloss_limited'])) < 1e-6 * max(float(g['qloss_limited']), 1e-3)
    n_all = len(idx_helper.num_idx_learned_as_num) + len(idx_helper.num_idx_learned_as_cat)
    al = ploss.CategoricalParamsAccuracy(idx_helper, limited_vst_params_indexes=lim)(v_out, v_in)
    assert abs(al.item() - float(g['accuracy_limited'])) < 1e-4
    both = ploss.PresetMetrics(idx_helper)(v_out, v_in)
    assert abs(both[0].item() - q.item()) == 0 and abs(both[1].item() - acc.item()) == 0 and both[2].item() == n_all


def test_learnable_to_full_presets_on_device(golden_dir, idx_helper):
    g = _g(golden_dir)
    v_out = torch.from_numpy(g['v_out']).to(DEV)
    full = ppreset.learnable_to_full_presets(idx_helper, v_out, ppreset.DexedLearnableLayout().params_default_values)
    assert full.is_cuda and full.shape == (48, 155)
    assert torch.equal(full.cpu(), torch.from_numpy(g['full_presets']))                  # argmax / copy / defaults: bit-exact
    ties = torch.zeros(3, 610, device=DEV)                                               # all-equal groups: first index wins, like torch.argmax
    full_t = ppreset.learnable_to_full_presets(idx_helper, ties, {})
    assert torch.equal(full_t.cpu(), oloss.learnable_to_full(idx_helper, ties.cpu(), {}))


def test_dkl_and_l2_wrappers_match_the_reference(golden_dir):
    g = _g(golden_dir)
    ml = torch.from_numpy(g['mu_logvar']).to(DEV).requires_grad_()
    d = ploss.GaussianDkl(normalize=True)(ml[:, 0], ml[:, 1])
    assert abs(d.item() - float(g['dkl'])) < 2e-6 * float(g['dkl'])
    assert abs(ploss.GaussianDkl(normalize=False)(ml[:, 0], ml[:, 1]).item() - float(g['dkl_raw'])) < 2e-6 * float(g['dkl_raw'])
    d.backward()
    mlc = torch.from_numpy(g['mu_logvar']).double().requires_grad_()
    omodel.gaussian_dkl(mlc[:, 0], mlc[:, 1], True).backward()
    assert float((ml.grad.cpu().double() - mlc.grad).norm() / mlc.grad.norm()) < 1e-5
    a, b = torch.from_numpy(g['l2_a']).to(DEV).requires_grad_(), torch.from_numpy(g['l2_b']).to(DEV)
    got = [ploss.L2Loss(c, ba)(a, b).item() for c in (False, True) for ba in (False, True)]          # model/loss.py:15-43
    assert np.allclose(got, g['l2'], rtol=2e-6)
    ploss.L2Loss()(a, b).backward()
    assert float((a.grad - 2 * (a.detach() - b) / 3).abs().max()) < 1e-6


def test_flow_params_loss_and_inverse_flow_gradients(golden_dir, idx_helper):
    """forward_controls_loss=False (train.py:117-119): FlowParamsLoss differentiates through the INVERSE latent flow (new coupling
    inverse backward kernel) and the regression flow run 'backwards' (regression.py:179-184)."""
    g = _g(golden_dir)
    m_cfg, t_cfg = pcfg.make_default(minibatch_size=6, forward_controls_loss=False)
    pcfg.apply_dataset_dims(m_cfg, idx_helper)
    torch.manual_seed(0)
    orc = omodel.build_extended_ae_model(m_cfg, t_cfg, idx_helper)[3].train()
    torch.manual_seed(0)
    mine = build.build_extended_ae_model(m_cfg, t_cfg, idx_helper)[3]
    mine.load_state_dict(orc.state_dict())
    mine.to(DEV).train()
    for mdl in (orc, mine):
        for blk in [m for m in mdl.modules() if type(m).__name__ == 'ResidualBlock']:
            blk.dropout.p = 0.0
    assert not mine.reg_model.is_flow_fast_forward
    ops.set_precision('fp32')
    try:
        ml = torch.from_numpy(g['mu_logvar'][:6]).to(DEV).requires_grad_()
        v = torch.from_numpy(g['flow_params_v']).to(DEV)
        crit = ploss.FlowParamsLoss(idx_helper, mine.ae_model.flow_inverse_function, mine.reg_model.flow_inverse_function)
        loss = crit(ml, v)
        loss.backward()
        torch.cuda.synchronize()
    finally:
        ops.set_precision('tf32')
    assert abs(loss.item() - float(g['flow_params_loss'])) < 2e-5 * abs(float(g['flow_params_loss']))
    assert float((ml.grad.cpu() - torch.from_numpy(g['flow_params_dml'])).norm() / np.linalg.norm(g['flow_params_dml'])) < 1e-4
    mlo = torch.from_numpy(g['mu_logvar'][:6]).double().requires_grad_()
    orc = orc.double()
    oloss.flow_params_loss(orc.ae_model.flow_transform.inverse, orc.reg_model.flow_inverse_function, mlo, v.cpu().double()).backward()
    ref = dict(orc.named_parameters())
    worst = 0.0
    for n, p in mine.named_parameters():
        r = ref[n].grad
        assert (p.grad is None) == (r is None), n
        if r is not None and r.norm() > 1e-12:
            worst = max(worst, float((p.grad.cpu().double() - r).norm() / r.norm()))
    assert worst < 2e-3, worst
    # eval-mode inverse still runs without autograd, and undoes the forward map
    mine.eval()
    with torch.no_grad():
        z = torch.randn(6, 610, device=DEV)
        y, ld = mine.ae_model.flow_transform(z)
        back, ldi = mine.ae_model.flow_transform.inverse(y)
    assert not back.requires_grad and float((back - z).abs().max()) < 1e-4 and float((ld + ldi).abs().max()) < 1e-3


def test_nan_flags_and_spectrogram_statistics():
    flags = torch.zeros(1, dtype=torch.int32, device=DEV)
    a, b, c = torch.tensor([1.0], device=DEV), torch.tensor([float('nan')], device=DEV), torch.tensor([float('inf')], device=DEV)
    ops.nan_flags_(flags, a, b, c, b)
    assert flags.item() == 0b1010
    g = torch.Generator().manual_seed(3)
    specs = torch.stack([torch.randn(257, 347, generator=g) * (i + 1) - 60.0 for i in range(7)])
    per, ds = ops.spectrogram_stats(specs.to(DEV))
    per_ref, ds_ref = oloss.spectrogram_stats(list(specs))
    assert np.array_equal(per[:, :2].cpu().numpy(), per_ref[:, :2].astype(np.float32))               # min / max exact
    assert np.allclose(per[:, 2:].cpu().numpy(), per_ref[:, 2:], rtol=2e-6)
    assert ds[0].item() == np.float32(ds_ref['min']) and ds[1].item() == np.float32(ds_ref['max'])
    assert abs(ds[2].item() - ds_ref['mean']) < 1e-5 * abs(ds_ref['mean']) and abs(ds[3].item() - ds_ref['std']) < 1e-5 * ds_ref['std']
