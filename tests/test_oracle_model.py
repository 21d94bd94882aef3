"""CPU: the model oracle against fixtures written from the REFERENCE's outputs (oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import losses as oloss, model as omodel
from preset_gen_vae_b200 import config as pcfg, synthetic

SIX_NOTES = ((40, 85), (50, 85), (60, 42), (60, 85), (60, 127), (70, 85))


def build_oracle(idx_helper, B, midi_notes=None, stack=False):
    m_cfg, t_cfg = pcfg.make_default(minibatch_size=B, midi_notes=midi_notes, stack_spectrograms=stack)
    pcfg.apply_dataset_dims(m_cfg, idx_helper)
    torch.manual_seed(0)
    return omodel.build_extended_ae_model(m_cfg, t_cfg, idx_helper)[3], m_cfg, t_cfg


@pytest.mark.parametrize("tag,B,notes,stack", [("c1_b4", 4, None, False), ("c6_b2", 2, SIX_NOTES, True)])
def test_oracle_reproduces_reference_fixture(golden_dir, idx_helper, tag, B, notes, stack):
    g = np.load(os.path.join(golden_dir, 'model_%s.npz' % tag))
    ext, m_cfg, t_cfg = build_oracle(idx_helper, B, notes, stack)
    sd = ext.state_dict()
    assert len(sd) == int(g['n_state_entries']) == 530
    assert sum(p.numel() for p in ext.parameters()) == int(g['n_params'])
    params = dict(ext.named_parameters())
    names = [str(n) for n in g['param_names']]
    fp = g['param_fingerprints']
    assert names == list(params.keys())
    for i, n in enumerate(names):      # seeded init == the reference's init
        assert abs(float(params[n].double().sum()) - fp[i, 0]) <= 1e-6 * max(1.0, abs(fp[i, 0])), n
        assert abs(float(params[n].double().norm()) - fp[i, 1]) <= 1e-6 * max(1.0, fp[i, 1]), n
    C = m_cfg.input_tensor_size[1]
    x = synthetic.make_spectrogram_like(B, C, seed=0)
    v_in = synthetic.make_preset_targets(idx_helper, B, seed=0)
    info = synthetic.make_sample_info(B)
    noise = synthetic.make_noise(B, m_cfg.dim_z, t_cfg.fc_dropout, t_cfg.reg_fc_dropout, seed=1,
                                 enc_fc_in=ext.ae_model.encoder.mlp[1].in_features)
    ext.train()
    outs, losses, total = oloss.train_step_losses(ext, x, v_in, info, noise, beta=0.2)
    total.backward()
    tol = dict(rtol=2e-4, atol=2e-5)   # fixture written on another host: conv/BLAS summation order may differ
    for k in ('z0_mu_logvar', 'z0', 'zK', 'logdet', 'v_out'):
        assert np.allclose(outs[k].detach().numpy(), g[k], **tol), k
    assert np.allclose(outs['x_out'].detach().numpy()[:, :, ::4, ::4], g['x_out_sub'], **tol)
    got = np.asarray([losses[k].item() for k in ('recons', 'latent', 'controls')])
    assert np.allclose(got, g['losses'], rtol=1e-4)
    for i, n in enumerate(names):      # gradient norms
        gn = float(params[n].grad.double().norm())
        assert abs(gn - fp[i, 2]) <= 5e-3 * fp[i, 2] + 1e-7, (n, gn, fp[i, 2])
    ext.eval()
    with torch.no_grad():
        ev = ext(x, info)
        v_ev = ext.reg_model(ev[2])
    assert np.allclose(ev[2].numpy(), g['eval_zK'], **tol)
    assert np.allclose(v_ev.numpy(), g['eval_v_out'], **tol)
    assert np.allclose(ev[4].numpy()[:, :, ::4, ::4], g['eval_x_out_sub'], **tol)


def test_flow_known_answers(idx_helper):
    """inverse(forward(z)) == z and logdet_fwd == -logdet_inv in eval mode; GaussianDkl(0,0) == 0."""
    ext, m_cfg, _ = build_oracle(idx_helper, 4)
    ext.eval()
    torch.manual_seed(5)
    z = torch.randn(4, m_cfg.dim_z)
    with torch.no_grad():
        for flow in (ext.ae_model.flow_transform, ext.reg_model._forward_flow_transform):
            # give the between-layer BatchNorm transforms non-degenerate running stats
            for t in flow._transforms:
                if hasattr(t, 'running_var') and t.running_var.abs().sum() == 0:
                    t.running_var.fill_(1.0)
            y, ld = flow.forward(z)
            zi, ldi = flow.inverse(y)
            assert torch.allclose(zi, z, atol=1e-4)
            assert torch.allclose(ld, -ldi, atol=1e-4)
    assert float(omodel.gaussian_dkl(torch.zeros(3, 7), torch.zeros(3, 7))) == 0.0


def test_latent_loss_identity_flow():
    """With an identity flow latent_loss reduces to mean(0.5*sum(zK^2 - logvar - eps^2)) / D (probability.py:18,28-29)."""
    torch.manual_seed(0)
    B, D = 5, 12
    mu, logvar, eps = torch.randn(B, D), torch.randn(B, D) * 0.3, torch.randn(B, D)
    z0 = mu + torch.exp(logvar / 2) * eps
    vae = omodel.FlowVAE.__new__(omodel.FlowVAE)
    torch.nn.Module.__init__(vae)
    vae.normalize_latent_loss = True
    got = vae.latent_loss(torch.stack([mu, logvar], dim=1), z0, z0, torch.zeros(B))
    want = (0.5 * (z0 ** 2 - logvar - eps ** 2).sum(dim=1)).mean() / D
    assert torch.allclose(got, want, atol=1e-5)


@pytest.mark.parametrize("between_bn", [False, True])
def test_flow_logdet_equals_the_log_abs_determinant_of_the_autograd_jacobian(between_bn):
    """An anchor for the nflows restatement that does not depend on nflows (which is absent here, DESIGN.md §3 "parity unpinned"):
    for a small RealNVP (affine couplings with ResidualNet conditioners, optional flow BatchNorm in eval mode) the logabsdet the
    transform returns must be log|det J| of the map itself, J from torch.autograd, and inverse(forward(z)) == z."""
    from oracle import nflows_port as nf
    torch.manual_seed(3)
    D = 6
    flow = nf.SimpleRealNVP(D, 8, num_layers=4, num_blocks_per_layer=2, batch_norm_within_layers=True,
                            batch_norm_between_layers=between_bn).double()
    t = flow._transform
    with torch.no_grad():                                  # non-trivial parameters everywhere (final conditioner layers start near zero)
        for p in t.parameters():
            p.add_(torch.randn_like(p) * 0.3)
        for m in t.modules():
            if hasattr(m, 'running_var'):
                m.running_var.uniform_(0.5, 2.0)
                m.running_mean.normal_()
    t.eval()
    z = torch.randn(3, D, dtype=torch.float64)
    y, ld = t.forward(z)
    for i in range(z.shape[0]):
        J = torch.autograd.functional.jacobian(lambda v: t.forward(v[None])[0][0], z[i])
        assert torch.allclose(torch.linalg.slogdet(J)[1], ld[i], atol=1e-9)
    zi, ldi = t.inverse(y)
    assert torch.allclose(zi, z, atol=1e-9) and torch.allclose(ldi, -ld, atol=1e-9)


def test_basic_vae_matches_the_reference(golden_dir):
    """BasicVAE (VAE.py:19-66), the model build.py:45-47 makes when latent_flow_arch is None: eval forward and Dkl latent loss of the
    oracle against the reference's committed outputs (tests/golden/basic_vae.npz)."""
    import os
    import numpy as np
    from preset_gen_vae_b200 import config as pcfg, synthetic
    g = np.load(os.path.join(golden_dir, 'basic_vae.npz'))
    m_cfg, t_cfg = pcfg.make_default(minibatch_size=3, latent_flow_arch=None, params_regression_architecture='mlp_3l1024')
    torch.manual_seed(0)
    enc = omodel.Encoder(m_cfg.encoder_architecture, 256, m_cfg.input_tensor_size, t_cfg.fc_dropout, output_bn=True, deepest_features_mix=False)
    dec = omodel.Decoder(m_cfg.encoder_architecture, 256, m_cfg.input_tensor_size, t_cfg.fc_dropout)
    ae = omodel.BasicVAE(enc, 256, dec, t_cfg.normalize_losses).eval()
    assert len(ae.state_dict()) == int(g['n_state_entries'])
    with torch.no_grad():
        out = ae(synthetic.make_spectrogram_like(3, 1, seed=2))
    assert out[3].shape == (3, 1) and torch.equal(out[1], out[2]) and torch.equal(out[1], out[0][:, 0, :])
    assert np.allclose(out[0].numpy(), g['z_mu_logvar'], rtol=1e-5, atol=1e-6)
    assert np.allclose(out[4].numpy()[:, :, ::4, ::4], g['x_out_sub'], rtol=1e-5, atol=1e-6)
    assert abs(ae.latent_loss(out[0]).item() - float(g['latent_loss'])) < 1e-6 * abs(float(g['latent_loss'])) + 1e-9
