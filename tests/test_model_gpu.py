"""GPU: the full ExtendedAE path (this package's modules -> C ABI -> CUDA kernels) against the CPU oracle on identical
seeded inputs, weights, eps and dropout masks.

The checker is the oracle evaluated in fp64 (same code as the fp32 oracle that tests/test_oracle_model.py pins to the
reference's golden outputs); measured on B200 (tools/gpu_diag_model.py, B=4) the reference's own fp32 CPU path is
5.1e-5 global / 7.3e-4 worst-tensor relative-L2 away from it, the CUDA fp32 path 1.9e-5 / 5.5e-5.

Tolerances (north_star: "per-step losses and gradients within a stated fp32/TF32 tolerance"):
  precision 'fp32' (exact-fp32 products everywhere): outputs 1e-4 relative-L2, losses 1e-5 relative, every parameter
      gradient within max(2e-3, 3 x the deviation of the reference's own fp32 path from fp64 on that tensor) relative-L2
      - small batches are ill-conditioned (BatchNorm over 2-4 samples): at B=2 the fp32 reference itself is 1e-2 away
      from fp64; tensors whose gradient is structurally zero (a bias feeding a train-mode BatchNorm) are
      compared absolutely.
  precision 'tf32' (every conv / transposed conv / Linear multiplies in TF32 on tcgen05, operands rounded to nearest):
      each kernel is 3e-4 relative-L2 from fp64 (tests/test_kernels_gpu.py); through 40 layers with BatchNorm over only
      B=4 samples this amplifies to outputs <= 2e-2, losses <= 5e-3, global gradient cosine >= 0.98 (see DESIGN.md for the
      measured values at B=4 / 16 / 160: the deviation shrinks with the batch as BatchNorm becomes well conditioned).
"""
import copy
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import losses as oloss, model as omodel
from preset_gen_vae_b200 import config as pcfg, synthetic
from preset_gen_vae_b200.model import build, loss as ploss, ops

pytestmark = pytest.mark.gpu
SIX_NOTES = ((40, 85), (50, 85), (60, 42), (60, 85), (60, 127), (70, 85))


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def make_pair(idx_helper, B, notes=None, stack=False, train_over=None, **model_over):
    m_cfg, t_cfg = pcfg.make_default(minibatch_size=B, midi_notes=notes, stack_spectrograms=stack, **model_over)
    for k, v in (train_over or {}).items():
        assert hasattr(t_cfg, k), k
        setattr(t_cfg, k, v)
    pcfg.apply_dataset_dims(m_cfg, idx_helper)
    torch.manual_seed(0)
    orc = omodel.build_extended_ae_model(m_cfg, t_cfg, idx_helper)[3]
    torch.manual_seed(0)
    mine = build.build_extended_ae_model(m_cfg, t_cfg, idx_helper)[3]
    return orc, mine, m_cfg, t_cfg


def to_dev(noise):
    out = {k: v.cuda() for k, v in noise.items() if torch.is_tensor(v)}
    out['reg_masks'] = [[m.cuda() for m in layer] for layer in noise['reg_masks']]
    return out


def run_step(orc, mine, idx_helper, m_cfg, t_cfg, B, precision, orc32=None):
    C = m_cfg.input_tensor_size[1]
    x = synthetic.make_spectrogram_like(B, C, seed=0)
    v_in = synthetic.make_preset_targets(idx_helper, B, seed=0)
    info = synthetic.make_sample_info(B)
    noise = synthetic.make_noise(B, m_cfg.dim_z, t_cfg.fc_dropout, t_cfg.reg_fc_dropout, seed=1,
                                 enc_fc_in=orc.ae_model.encoder.mlp[1].in_features)
    orc.train()
    noise64 = {k: (v.double() if torch.is_tensor(v) else [[m.double() for m in l] for l in v]) for k, v in noise.items()}
    outs, losses, total = oloss.train_step_losses(orc, x.double(), v_in.double(), info, noise64, beta=0.2,
                                                  params_reg_softmax=m_cfg.params_reg_softmax)
    total.backward()
    if orc32 is not None:                      # the reference's own fp32 path: its distance to fp64 is the noise floor
        orc32.train()
        _, _, t32 = oloss.train_step_losses(orc32, x, v_in, info, noise, beta=0.2, params_reg_softmax=m_cfg.params_reg_softmax)
        t32.backward()
    ops.set_precision(precision)
    mine.train()
    dn = to_dev(noise)
    z0_ml, z0, zk, logdet, x_out = mine(x.cuda(), info.cuda(), dn)
    flow_head = m_cfg.params_regression_architecture.startswith('flow_')
    v_out = mine.reg_model(zk, dropout_masks=dn['reg_masks'] if flow_head else None)
    recons = ploss.MSELoss()(x_out, x.cuda())
    lat = mine.latent_loss(z0_ml, z0, zk, logdet)
    crit = ploss.SynthParamsLoss(idx_helper, True, cat_bce=False, cat_softmax=not m_cfg.params_reg_softmax, cat_softmax_t=0.2)
    cont = crit(v_out, v_in.cuda())
    (recons + 0.2 * lat + cont).backward()
    torch.cuda.synchronize()
    got = dict(z0_mu_logvar=z0_ml, z0=z0, zK=zk, logdet=logdet, x_out=x_out, v_out=v_out)
    return outs, losses, got, dict(recons=recons, latent=lat, controls=cont)


def check(orc, mine, outs, losses, got, got_losses, out_tol, loss_tol, grad_tol, min_cos, orc32=None):
    """grad_tol: every gradient tensor must be within max(grad_tol, 3 x the fp32 reference's own deviation from fp64)."""
    ref32 = None if orc32 is None else {n: p.grad for n, p in orc32.named_parameters()}
    for k in outs:
        assert got[k].shape == outs[k].shape, k
        assert rel(got[k], outs[k]) < out_tol, (k, rel(got[k], outs[k]))
    for k in losses:
        assert abs(got_losses[k].item() - losses[k].item()) <= loss_tol * abs(losses[k].item()), (k, got_losses[k].item(), losses[k].item())
    ref_g = dict(orc.named_parameters())
    dot = n1 = n2 = 0.0
    worst = ('', 0.0)
    pending = []
    for name, p in mine.named_parameters():
        g, r = p.grad, ref_g[name].grad
        assert g is not None and r is not None, name
        g, r = g.double().cpu(), r.double()
        dot += float((g * r).sum()); n1 += float((g * g).sum()); n2 += float((r * r).sum())
        if r.norm() < 1e-7:                                                  # structurally zero gradient
            assert g.abs().max() < 1e-5, name
            continue
        e = float((g - r).norm() / r.norm())
        if e > worst[1]:
            worst = (name, e)
        if grad_tol is not None:
            floor = 0.0 if ref32 is None else 3.0 * float((ref32[name].double() - r).norm() / r.norm())
            pending.append((name, e, floor, float((g - r).norm())))
    cos = dot / np.sqrt(n1 * n2)
    for name, e, floor, abs_err in pending:       # tiny-norm tensors: an absolute bound relative to the whole gradient
        assert e < max(grad_tol, floor) or abs_err < 1e-4 * np.sqrt(n2), (name, e, floor, abs_err)
    print("worst per-tensor gradient rel-L2: %s %.3e | global cosine %.6f | global rel-L2 %.3e"
          % (worst[0], worst[1], cos, np.sqrt(max(n1 + n2 - 2 * dot, 0.0) / n2)))
    assert cos >= min_cos


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_train_step_parity_default_config(idx_helper, precision):
    B = 4
    orc, mine, m_cfg, t_cfg = make_pair(idx_helper, B)
    mine.load_state_dict(orc.state_dict())
    mine.cuda()
    orc32, orc = orc, copy.deepcopy(orc).double()
    res = run_step(orc, mine, idx_helper, m_cfg, t_cfg, B, precision, orc32 if precision == 'fp32' else None)
    if precision == 'fp32':
        check(orc, mine, *res, out_tol=1e-4, loss_tol=1e-5, grad_tol=2e-3, min_cos=0.999999, orc32=orc32)
    else:
        # measured on B200 (profiles/parity_study_r02.log): 0.99766 here, 0.99824 for cuDNN's own TF32 step on the same inputs;
        # tests/test_parity_default_gpu.py gates this precision against the live cuDNN-TF32 deviation at B = 8 and B = 160
        check(orc, mine, *res, out_tol=1e-2, loss_tol=1e-3, grad_tol=None, min_cos=0.996)
    ops.set_precision('tf32')
    # running statistics after one training forward
    sd_o, sd_m = orc.state_dict(), mine.state_dict()
    for k in sd_o:
        if 'running' in k:
            assert rel(sd_m[k], sd_o[k]) < (1e-4 if precision == 'fp32' else 2e-2), k
        if 'num_batches_tracked' in k:
            assert int(sd_m[k]) == int(sd_o[k]), k


def test_constructor_side_effect_and_eval_mode(idx_helper):
    B = 3
    orc, mine, m_cfg, t_cfg = make_pair(idx_helper, B)
    assert mine.ae_model.encoder.constructor_bn_side_effect_applied
    sd_o, sd_m = orc.state_dict(), mine.state_dict()
    assert list(sd_o.keys()) == list(sd_m.keys())
    for k in sd_o:           # freshly built models agree, incl. the BN statistics touched by the shape-inference forward
        assert rel(sd_m[k].float(), sd_o[k].float()) < 1e-3 or float(sd_o[k].float().norm()) == 0, k
    # nflows initialises the flow-BatchNorm running_var to ZERO, so a never-trained model in eval mode divides by sqrt(eps)
    # and saturates the Hardtanh: give both models the same sane statistics before comparing eval outputs
    for mdl in (orc, mine):
        for t in mdl.reg_model._forward_flow_transform._transforms:
            if hasattr(t, 'running_var'):
                t.running_var.fill_(1.0)
    mine.cuda().eval()
    orc.eval()
    ops.set_precision('fp32')
    x = synthetic.make_spectrogram_like(B, 1, seed=3)
    info = synthetic.make_sample_info(B)
    with torch.no_grad():
        ev_o = orc(x, info)
        v_o = orc.reg_model(ev_o[2])
        ev_m = mine(x.cuda(), info.cuda())
        v_m = mine.reg_model(ev_m[2])
    ops.set_precision('tf32')
    for a, b in zip(ev_m, ev_o):
        assert rel(a, b) < 5e-4
    assert rel(v_m, v_o) < 5e-4
    assert torch.equal(ev_m[1], ev_m[0][:, 0, :])            # eval: z0 = mu (VAE.py:175-176)
    # inverse flow (evaluation only): inverse(forward(z)) == z, log-dets cancel
    flow = mine.ae_model.flow_transform
    with torch.no_grad():
        y, ld = flow(ev_m[1])
        back, ldi = flow.inverse(y)
    assert rel(back, ev_m[1]) < 1e-4 and float((ld + ldi).abs().max()) < 1e-3


def test_stacked_six_channel_config(idx_helper):
    B = 3
    orc, mine, m_cfg, t_cfg = make_pair(idx_helper, B, SIX_NOTES, True)
    assert m_cfg.input_tensor_size[1] == 6
    mine.load_state_dict(orc.state_dict())
    mine.cuda()
    orc32, orc = orc, copy.deepcopy(orc).double()
    res = run_step(orc, mine, idx_helper, m_cfg, t_cfg, B, 'fp32', orc32)
    ops.set_precision('tf32')
    check(orc, mine, *res, out_tol=2e-4, loss_tol=2e-5, grad_tol=2e-3, min_cos=0.99999, orc32=orc32)


def test_midi_concat_and_softmax_head_config(idx_helper):
    """Six notes, not stacked: concat_midi_to_z (VAE.py:155-165), bigger network (1800 channels), and
    params_reg_softmax=True (regression.py:47-50 + loss without its own softmax)."""
    B = 3
    orc, mine, m_cfg, t_cfg = make_pair(idx_helper, B, SIX_NOTES, False, params_reg_softmax=True)
    assert m_cfg.concat_midi_to_z and m_cfg.input_tensor_size[1] == 1
    mine.load_state_dict(orc.state_dict())
    mine.cuda()
    orc32, orc = orc, copy.deepcopy(orc).double()
    res = run_step(orc, mine, idx_helper, m_cfg, t_cfg, B, 'fp32', orc32)
    ops.set_precision('tf32')
    check(orc, mine, *res, out_tol=2e-4, loss_tol=2e-5, grad_tol=2e-3, min_cos=0.99999, orc32=orc32)


def test_deepest_features_mix_layout(idx_helper):
    """stack_specs_deepest_features_mix=True (encoder.py:55-58): the shared CNN keeps enc7 and the mixer is the single 1x1 convolution
    512*C -> 1024; state_dict keys `single_ch_cnn.enc_nn.4x4conv.enc7conv.*` / `features_mixer_cnn.enc8conv.*`."""
    B = 3
    orc, mine, m_cfg, t_cfg = make_pair(idx_helper, B, SIX_NOTES, True, stack_specs_deepest_features_mix=True)
    keys = list(orc.state_dict().keys())
    assert keys == list(mine.state_dict().keys())
    assert 'ae_model.encoder.single_ch_cnn.enc_nn.4x4conv.enc7conv.weight' in keys and 'ae_model.encoder.features_mixer_cnn.enc8conv.weight' in keys
    assert mine.state_dict()['ae_model.encoder.features_mixer_cnn.enc8conv.weight'].shape == (1024, 512 * 6, 1, 1)
    mine.load_state_dict(orc.state_dict())
    mine.cuda()
    orc32, orc = orc, copy.deepcopy(orc).double()
    res = run_step(orc, mine, idx_helper, m_cfg, t_cfg, B, 'fp32', orc32)
    ops.set_precision('tf32')
    check(orc, mine, *res, out_tol=2e-4, loss_tol=2e-5, grad_tol=2e-3, min_cos=0.99999, orc32=orc32)


def test_basic_vae(idx_helper):
    """BasicVAE (VAE.py:19-66; latent_flow_arch=None, build.py:45-47): 5-tuple with z_K = z_0 and a zero [B, 1] log-det, Dkl latent loss;
    forward and backward against the oracle with shared eps / dropout masks."""
    B = 4
    m_cfg, t_cfg = pcfg.make_default(minibatch_size=B, latent_flow_arch=None, params_regression_architecture='mlp_3l1024')
    torch.manual_seed(0)
    o_enc = omodel.Encoder(m_cfg.encoder_architecture, 256, m_cfg.input_tensor_size, t_cfg.fc_dropout, output_bn=True, deepest_features_mix=False)
    o_dec = omodel.Decoder(m_cfg.encoder_architecture, 256, m_cfg.input_tensor_size, t_cfg.fc_dropout)
    orc = omodel.BasicVAE(o_enc, 256, o_dec, t_cfg.normalize_losses).double().train()
    torch.manual_seed(0)
    _, _, mine = build.build_ae_model(m_cfg, t_cfg)
    assert type(mine).__name__ == 'BasicVAE' and list(mine.state_dict().keys()) == list(orc.state_dict().keys())
    mine.load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in orc.state_dict().items()})
    mine.cuda().train()
    x = synthetic.make_spectrogram_like(B, 1, seed=4)
    noise = synthetic.make_noise(B, 256, t_cfg.fc_dropout, 0.0, seed=2)
    n64 = {k: v.double() for k, v in noise.items() if torch.is_tensor(v)}
    o = orc(x.double(), n64)
    (F.mse_loss(o[4], x.double()) + 0.2 * orc.latent_loss(o[0])).backward()
    ops.set_precision('fp32')
    try:
        xm = x.cuda()
        m = mine(xm, {k: v.cuda() for k, v in noise.items() if torch.is_tensor(v)})
        lat = mine.latent_loss(m[0])
        (ploss.MSELoss()(m[4], xm) + 0.2 * lat).backward()
        torch.cuda.synchronize()
    finally:
        ops.set_precision('tf32')
    assert m[3].shape == (B, 1) and float(m[3].abs().max()) == 0.0 and torch.equal(m[1], m[2])
    for a, b in zip((m[0], m[1], m[4]), (o[0], o[1], o[4])):
        assert rel(a, b) < 1e-4
    assert abs(lat.item() - orc.latent_loss(o[0]).item()) < 1e-5 * abs(orc.latent_loss(o[0]).item())
    ref = dict(orc.named_parameters())
    for n, p in mine.named_parameters():
        r = ref[n].grad
        if r.norm() > 1e-7:
            assert rel(p.grad, r) < 1e-2, (n, rel(p.grad, r))      # the first layer's 8-element bias gradient sits at 4e-3..5e-3


def test_dataparallel_wrap_of_the_reference_train_script(idx_helper):
    """train.py:95-97 wraps the model (and its regression head) in nn.DataParallel; with one device id that must be transparent."""
    B = 3
    _, mine, m_cfg, t_cfg = make_pair(idx_helper, B)
    mine.cuda().eval()
    ae_par = torch.nn.DataParallel(mine, device_ids=[0], output_device='cuda:0')
    reg_par = torch.nn.DataParallel(mine.reg_model, device_ids=[0], output_device='cuda:0')
    x, info = synthetic.make_spectrogram_like(B, 1, seed=3).cuda(), synthetic.make_sample_info(B).cuda()
    with torch.no_grad():
        a = ae_par(x, info)
        b = mine(x, info)
        va, vb = reg_par(a[2]), mine.reg_model(b[2])
    for u, v in zip(a, b):
        assert torch.equal(u, v)
    assert torch.equal(va, vb) and va.shape == (B, 610)


def test_mlp_regression_head_config(idx_helper):
    """params_regression_architecture = 'mlp_3l1024' (regression.py:61-102: Linear -> BatchNorm1d -> Dropout -> ReLU stack and the
    PresetActivation) in the full training step; reg_fc_dropout = 0 because the reference draws this head's masks from nn.Dropout."""
    B = 4
    orc, mine, m_cfg, t_cfg = make_pair(idx_helper, B, train_over={'reg_fc_dropout': 0.0}, params_regression_architecture='mlp_3l1024')
    assert type(mine.reg_model).__name__ == 'MLPRegression' and type(orc.reg_model).__name__ == 'MLPRegression'
    mine.load_state_dict(orc.state_dict())
    mine.cuda()
    orc32, orc = orc, copy.deepcopy(orc).double()
    res = run_step(orc, mine, idx_helper, m_cfg, t_cfg, B, 'fp32', orc32)
    ops.set_precision('tf32')
    check(orc, mine, *res, out_tol=2e-4, loss_tol=2e-5, grad_tol=2e-3, min_cos=0.99999, orc32=orc32)


def test_reference_checkpoint_layout_round_trip(idx_helper, tmp_path):
    """state_dict keys / shapes are the reference's: a checkpoint dict in the reference's format
    (logs/logger.py:199-202) written from the oracle loads into this package's model and back."""
    orc, mine, m_cfg, t_cfg = make_pair(idx_helper, 2)
    path = tmp_path / '00000.tar'
    torch.save({'epoch': 0, 'ae_model_state_dict': orc.state_dict()}, path)
    ckpt = torch.load(path, map_location='cpu')
    missing = mine.load_state_dict(ckpt['ae_model_state_dict'])
    assert not missing.missing_keys and not missing.unexpected_keys
    sd = mine.state_dict()
    assert sd['ae_model.encoder.mlp.1.weight'].shape == (1220, 24576)
    assert sd['ae_model.decoder.single_ch_cnn.dec_nn.6.weight'].shape == (8, 1, 5, 5)
    assert sd['ae_model.flow_transform._transforms.0.identity_features'].dtype == torch.int64
    assert sd['reg_model._forward_flow_transform._transforms.1.running_var'].shape == (610,)
    orc.load_state_dict(sd)


def test_unsupported_configurations_raise(idx_helper):
    from preset_gen_vae_b200.model import encoder, decoder, VAE
    with pytest.raises(NotImplementedError):
        encoder.SpectrogramEncoder('flow_synth', 610, (4, 1, 257, 347), 0.3)
    with pytest.raises(NotImplementedError):
        decoder.SpectrogramDecoder('wavenet_baseline', 610, (4, 1, 513, 433), 0.3)
    with pytest.raises(NotImplementedError):
        VAE.FlowVAE(None, 610, None, True, 'maf_6l300')
    with pytest.raises(AssertionError):
        VAE.FlowVAE(None, 610, None, True, 'realnvp')
    with pytest.raises(ValueError):
        ploss.SynthParamsLoss(idx_helper, True, cat_bce=True, cat_softmax=True)
