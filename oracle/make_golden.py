"""Validate the oracle against the REFERENCE and write tests/golden/*.  Runs only in the build container.

    python -m oracle.make_golden            # needs /root/reference (read-only); writes tests/golden/

What it does
  1. imports the reference's own modules from /root/reference unmodified (sys.path), with stand-ins for the
     packages that are absent here: `nflows` -> oracle.nflows_port (PARITY UNPINNED, see that file),
     `librosa.feature.melspectrogram` -> oracle.frontend.slaney_mel_filterbank (PARITY UNPINNED), and inert
     stubs for matplotlib / soundfile / librenderman (never executed on the hot path);
  2. runs the reference on the seeded synthetic inputs of preset_gen_vae_b200/synthetic.py;
  3. asserts that the oracle restatement reproduces the reference (state_dict, outputs, losses, gradients);
  4. writes the REFERENCE's outputs as small fixtures, so that on the GPU box (where /root/reference does not
     exist) tests can pin the oracle first and then use it to check the CUDA path.
"""
import json
import os
import sys
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
GOLDEN = os.path.join(REPO, 'tests', 'golden')
sys.path.insert(0, REPO)

from oracle import frontend as ofe, losses as oloss, model as omodel, nflows_port  # noqa: E402
from preset_gen_vae_b200 import synthetic  # noqa: E402
from preset_gen_vae_b200.data import preset as ppreset  # noqa: E402


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference():
    nflows_port.install_as_nflows()

    def melspectrogram(S=None, n_mels=128, norm='slaney', **kw):
        assert norm is None
        S = np.asarray(S)
        return np.dot(ofe.slaney_mel_filterbank(2 * (S.shape[0] - 1), n_mels), S)
    feature = _stub('librosa.feature', melspectrogram=melspectrogram)
    _stub('librosa.display')
    _stub('librosa', feature=feature, display=sys.modules['librosa.display'])
    _stub('matplotlib.pyplot')
    _stub('matplotlib', pyplot=sys.modules['matplotlib.pyplot'])
    _stub('soundfile')
    _stub('librenderman')
    sys.path.insert(0, REF)
    import config as ref_config
    import model.build as ref_build
    import model.loss as ref_loss
    import utils.audio as ref_audio
    import data.preset as ref_preset
    import synth.dexed as ref_dexed
    return ref_config, ref_build, ref_loss, ref_audio, ref_preset, ref_dexed


class _FakeDexedDataset:
    """What PresetIndexesHelper(dataset) reads, built from the reference's static Dexed tables with the rules of
    data/dexeddataset.py:79-167 (the real DexedDataset needs the LFS SQLite DB and 10.5 GB of wavs)."""
    synth_name = 'Dexed'

    def __init__(self, ref_dexed, learned_as_categorical='all<=32'):
        D = ref_dexed.Dexed
        self.total_nb_params = 155
        self.preset_param_names = ['p%d' % i for i in range(155)]
        self.algos = []
        learnable = list(range(155))
        for i in [0, 1, 2, 3, 13] + [44, 66, 88, 110, 132, 154]:
            learnable.remove(i)
        self.learnable_params_idx = learnable
        card = np.asarray([D.get_param_cardinality(i) for i in range(155)])
        card[[44, 66, 88, 110, 132, 154]] = 1
        card[[0, 1, 2, 3, 13]] = 1
        self._card = card
        self.params_default_values = {0: 1.0, 1: 0.0, 2: 1.0, 3: 0.5, 13: 0.5,
                                      **{i: 1.0 for i in [44, 66, 88, 110, 132, 154]}}
        self.numerical_vst_params = D.get_numerical_params_indexes()
        self.categorical_vst_params = D.get_categorical_params_indexes()
        thr = int(learned_as_categorical.replace('all<=', ''))
        self.vst_param_learnable_model = []
        for i in range(155):
            if i not in learnable:
                self.vst_param_learnable_model.append(None)
            elif i in self.numerical_vst_params:
                self.vst_param_learnable_model.append('cat' if 1 < card[i] <= thr else 'num')
            else:
                self.vst_param_learnable_model.append('cat')

    def get_preset_param_cardinality(self, idx, learnable_representation=True):
        return int(self._card[idx])


def relerr(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def golden_frontend(ref_audio):
    audio = synthetic.make_audio(4, 1, seed=0)
    spec = ref_audio.Spectrogram(1024, 256, -120.0)
    mel = ref_audio.MelSpectrogram(1024, 256, -120.0, 257, 22050)
    assert abs(spec.spectrogram_norm_factor - 511.5) < 1e-3
    lin_ref = [spec(audio[b, 0].numpy()) for b in range(4)]
    mel_ref = [mel(audio[b, 0].numpy()) for b in range(4)]
    for b in range(4):
        assert torch.equal(lin_ref[b], ofe.spectrogram_db(audio[b, 0], 1024, 256, -120.0)), "linear dB oracle"
        d = (mel_ref[b] - ofe.mel_spectrogram_db(audio[b, 0], 1024, 256, -120.0, 257)).abs().max().item()
        assert d < 2e-4, ("mel dB oracle", d)   # BLAS (numpy) vs torch.matmul summation order only
    assert lin_ref[0].shape == (513, 347) and mel_ref[0].shape == (257, 347)
    # KATs: silence -> floor everywhere; unit sine at a bin centre -> ~ -6.02 dB peak
    assert torch.all(spec(np.zeros(synthetic.CLIP_SAMPLES, dtype=np.float32)) == -120.0)
    n = np.arange(synthetic.CLIP_SAMPLES)
    sine = np.sin(2 * np.pi * 64 * n / 1024).astype(np.float32)
    peak = spec(sine)[64, 100].item()
    assert abs(peak - 20 * np.log10(0.5)) < 1e-2, peak
    mel64 = ofe.mel_spectrogram_db(audio[:, 0], 1024, 256, -120.0, 257, dtype=torch.float64)
    np.savez_compressed(os.path.join(GOLDEN, 'frontend.npz'),
                        lin_db_clip0=lin_ref[0].numpy(), mel_db=torch.stack(mel_ref).numpy(),
                        mel_db_fp64_max=np.float64(mel64.max().item()), sine_peak_db=np.float32(peak),
                        norm_factor=np.float64(spec.spectrogram_norm_factor))
    print("front end: oracle == reference (linear dB bit-exact, mel dB within %.1e)" % 2e-4,
          "| mel dB range", float(torch.stack(mel_ref).min()), float(torch.stack(mel_ref).max()))


def golden_index_tables(ref_preset, ref_dexed):
    fake = _FakeDexedDataset(ref_dexed)
    ref_helper = ref_preset.PresetIndexesHelper(fake)
    fake.preset_indexes_helper = ref_helper
    mine = ppreset.DexedLearnableLayout().preset_indexes_helper
    assert ref_helper.learnable_preset_size == mine.learnable_preset_size == 610
    assert ref_helper.full_to_learnable == mine.full_to_learnable
    assert list(ref_helper.learnable_to_full) == list(mine.learnable_to_full)
    assert ref_helper.get_numerical_learnable_indexes() == mine.get_numerical_learnable_indexes()
    assert ref_helper.get_categorical_learnable_indexes() == mine.get_categorical_learnable_indexes()
    assert list(ref_helper.vst_param_cardinals) == list(mine.vst_param_cardinals)
    v = synthetic.make_preset_targets(mine, 32, seed=0)
    for row in range(32):
        a, b = ref_helper.get_useless_learned_params_indexes(v[row]), mine.get_useless_learned_params_indexes(v[row])
        assert a == b
    t_ref, t_mine = oloss._tables_from_reference(ref_helper), mine.device_tables()
    for k in t_ref:
        assert np.array_equal(t_ref[k], t_mine[k]), k
    with open(os.path.join(GOLDEN, 'dexed_layout.json'), 'w') as f:
        json.dump({'full_to_learnable': ref_helper.full_to_learnable,
                   'vst_param_learnable_model': ref_helper.vst_param_learnable_model,
                   'vst_param_cardinals': [int(c) for c in ref_helper.vst_param_cardinals],
                   'learnable_preset_size': ref_helper.learnable_preset_size,
                   'device_tables': {k: v_.tolist() for k, v_ in t_ref.items()}}, f)
    # inference tail (data/preset.py:350-369)
    full_ref = ref_preset.DexedPresetsParams(fake, learnable_presets=v).get_full()
    full_mine = ppreset.learnable_to_full_presets(mine, v, mine_defaults())
    assert torch.allclose(full_ref, full_mine), "learnable->full preset conversion"
    print("index tables: product helper == reference helper (610 columns, 90 num + 54 cat groups)")
    return fake, ref_helper, mine


def mine_defaults():
    return ppreset.DexedLearnableLayout().params_default_values


def _configure(cfg, B, midi_notes=None, stack=False):
    cfg.model.midi_notes = ((60, 85),) if midi_notes is None else midi_notes
    cfg.model.stack_spectrograms = stack
    cfg.train.minibatch_size = B
    cfg.model.synth_params_count = 144
    cfg.model.learnable_params_tensor_length = 610
    cfg.model.dim_z = 610
    cfg.update_dynamic_config_params()


def _ref_step(ref_build, ref_loss, cfg, helper, x_in, v_in, info, beta, seed_noise):
    torch.manual_seed(0)
    _, _, _, ext = ref_build.build_extended_ae_model(cfg.model, cfg.train, helper)
    ext.train()
    crit = ref_loss.SynthParamsLoss(helper, cfg.train.normalize_losses, cat_bce=cfg.train.params_cat_bceloss,
                                    cat_softmax=(not cfg.model.params_reg_softmax and not cfg.train.params_cat_bceloss),
                                    cat_softmax_t=cfg.train.params_cat_softmax_temperature)
    torch.manual_seed(seed_noise)
    out = ext(x_in, info)
    v_out = ext.reg_model(out[2])
    recons = torch.nn.MSELoss(reduction='mean')(out[4], x_in)
    lat = ext.latent_loss(*out[:4])
    v_out_for_loss, v_in_for_loss = v_out, v_in.clone()       # criterion mutates both (loss.py:134-135)
    v_out_snapshot = v_out.detach().clone()
    cont = crit(v_out_for_loss, v_in_for_loss)
    total = recons + beta * lat + cont
    total.backward()
    return ext, out, v_out_snapshot, dict(recons=recons, latent=lat, controls=cont), total


def golden_model(ref_config, ref_build, ref_loss, ref_helper, my_helper, tag, B, midi_notes=None, stack=False):
    cfg = ref_config
    _configure(cfg, B, midi_notes, stack)
    C = cfg.model.input_tensor_size[1]
    x_in = synthetic.make_spectrogram_like(B, C, seed=0)
    v_in = synthetic.make_preset_targets(my_helper, B, seed=0)
    info = synthetic.make_sample_info(B)
    beta = 0.2
    ext_ref, out_ref, v_out_ref, losses_ref, total_ref = _ref_step(ref_build, ref_loss, cfg, ref_helper, x_in, v_in,
                                                                  info, beta, seed_noise=1)
    sd_ref_after = {k: v.clone() for k, v in ext_ref.state_dict().items()}   # running stats after 1 train forward
    # ---- oracle, same seeds, explicit noise ----
    from preset_gen_vae_b200 import config as pcfg
    m_cfg, t_cfg = pcfg.make_default(minibatch_size=B, midi_notes=midi_notes, stack_spectrograms=stack)
    pcfg.apply_dataset_dims(m_cfg, my_helper)
    torch.manual_seed(0)
    _, _, _, ext_or = omodel.build_extended_ae_model(m_cfg, t_cfg, my_helper)
    torch.manual_seed(0)
    _, _, _, ext_ref_fresh = ref_build.build_extended_ae_model(cfg.model, cfg.train, ref_helper)
    sd_init = ext_ref_fresh.state_dict()
    sd_or = ext_or.state_dict()
    assert list(sd_init.keys()) == list(sd_or.keys()), "state_dict keys/order differ"
    for k in sd_init:
        assert torch.equal(sd_init[k], sd_or[k]), ("init differs", k)
    ext_or.train()
    noise = synthetic.make_noise(B, 610, t_cfg.fc_dropout, t_cfg.reg_fc_dropout, seed=1,
                                 enc_fc_in=ext_or.ae_model.encoder.mlp[1].in_features)
    outs, losses, total = oloss.train_step_losses(ext_or, x_in, v_in, info, noise, beta)
    total.backward()
    names = ['z0_mu_logvar', 'z0', 'zK', 'logdet', 'x_out']
    for n, r in zip(names, out_ref):
        assert relerr(outs[n], r) < 1e-5, (tag, n, relerr(outs[n], r))
    assert relerr(outs['v_out'], v_out_ref) < 1e-5
    for k in losses_ref:
        assert abs(losses[k].item() - losses_ref[k].item()) <= 1e-5 * abs(losses_ref[k].item()), (k, losses[k], losses_ref[k])
    g_ref = dict(ext_ref.named_parameters())
    worst = 0.0
    for n, p in ext_or.named_parameters():
        gr = g_ref[n].grad
        assert gr is not None and p.grad is not None, n
        denom = gr.norm().item()
        if denom > 1e-12:
            worst = max(worst, (p.grad - gr).norm().item() / denom)
    assert worst < 1e-3, worst
    for k, v in ext_or.state_dict().items():
        if 'running' in k or 'num_batches' in k:
            assert torch.allclose(v.float(), sd_ref_after[k].float(), rtol=1e-5, atol=1e-6), k
    # ---- eval mode ----
    ext_ref.eval()
    ext_or.eval()
    with torch.no_grad():
        ev_ref = ext_ref(x_in, info)
        v_ev_ref = ext_ref.reg_model(ev_ref[2])
        ev_or = ext_or(x_in, info)
        v_ev_or = ext_or.reg_model(ev_or[2])
    for a, b in zip(ev_or, ev_ref):
        assert relerr(a, b) < 1e-5
    assert relerr(v_ev_or, v_ev_ref) < 1e-5
    # ---- fixtures: reference outputs + per-tensor fingerprints of weights and gradients ----
    fp = {}
    for n, p in ext_ref.named_parameters():
        g = p.grad.double()
        fp[n] = [float(sd_init[n].double().sum()), float(sd_init[n].double().norm()), float(g.norm()),
                 float(g.flatten()[:: max(1, g.numel() // 7)][:7].sum())]
    np.savez_compressed(
        os.path.join(GOLDEN, 'model_%s.npz' % tag),
        z0_mu_logvar=out_ref[0].detach().numpy(), z0=out_ref[1].detach().numpy(), zK=out_ref[2].detach().numpy(),
        logdet=out_ref[3].detach().numpy(), x_out_sub=out_ref[4].detach().numpy()[:, :, ::4, ::4],
        x_out_sum=np.float64(out_ref[4].double().sum().item()), v_out=v_out_ref.numpy(),
        losses=np.asarray([losses_ref[k].item() for k in ('recons', 'latent', 'controls')], dtype=np.float64),
        eval_zK=ev_ref[2].numpy(), eval_v_out=v_ev_ref.numpy(), eval_x_out_sub=ev_ref[4].numpy()[:, :, ::4, ::4],
        param_names=np.asarray(list(fp.keys())), param_fingerprints=np.asarray(list(fp.values()), dtype=np.float64),
        n_state_entries=np.int64(len(sd_init)), n_params=np.int64(sum(p.numel() for p in ext_ref.parameters())))
    print("model[%s] B=%d C=%d: oracle == reference | losses" % (tag, B, C),
          {k: round(v.item(), 6) for k, v in losses_ref.items()}, "| worst grad rel-L2 vs ref %.2e" % worst,
          "| params", sum(p.numel() for p in ext_ref.parameters()), "state entries", len(sd_init))


def golden_metrics(ref_loss, ref_preset, ref_build, ref_config, fake, ref_helper, my_helper):
    """Monitoring metrics (loss.py:187-315), FlowParamsLoss (loss.py:318-346), GaussianDkl / L2Loss (loss.py:15-66): the reference
    classes on seeded inputs; the oracle restatements must reproduce them; the reference's numbers become tests/golden/metrics.npz."""
    B = 48
    g = torch.Generator().manual_seed(11)
    v_in = synthetic.make_preset_targets(my_helper, B, seed=3)
    v_out = (v_in + 0.35 * torch.randn(B, 610, generator=g)).clamp_(0.0, 1.0)      # a plausible, partly wrong inference
    v_out[:, 5] = 0.5                                                               # exact .5 products exercise round-half-even
    qmse = ref_loss.QuantizedNumericalParamsLoss(ref_helper, numerical_loss=torch.nn.MSELoss(reduction='mean'))(v_out, v_in)
    ql1 = ref_loss.QuantizedNumericalParamsLoss(ref_helper, numerical_loss=torch.nn.L1Loss())(v_out, v_in)
    acc = ref_loss.CategoricalParamsAccuracy(ref_helper, reduce=True, percentage_output=True)(v_out, v_in)
    acc_d = ref_loss.CategoricalParamsAccuracy(ref_helper, reduce=False, percentage_output=False)(v_out, v_in)
    lim = [4, 5, 6, 30, 40, 52, 100]
    q_lim = ref_loss.QuantizedNumericalParamsLoss(ref_helper, limited_vst_params_indexes=lim)(v_out, v_in)
    acc_lim = ref_loss.CategoricalParamsAccuracy(ref_helper, limited_vst_params_indexes=lim)(v_out, v_in)
    assert abs(oloss.quantized_numerical_params_loss(my_helper, v_out, v_in).item() - qmse.item()) < 1e-7
    assert abs(oloss.quantized_numerical_params_loss(my_helper, v_out, v_in, l1=True).item() - ql1.item()) < 1e-7
    assert abs(oloss.categorical_params_accuracy(my_helper, v_out, v_in) - float(acc)) < 1e-9
    assert oloss.categorical_params_accuracy(my_helper, v_out, v_in, reduce=False, percentage_output=False) == acc_d
    full_ref = ref_preset.DexedPresetsParams(fake, learnable_presets=v_out).get_full()
    assert torch.equal(full_ref, oloss.learnable_to_full(my_helper, v_out, mine_defaults()))
    # GaussianDkl / L2Loss
    ml = 0.5 * torch.randn(B, 2, 610, generator=g)
    dkl = ref_loss.GaussianDkl(normalize=True)(ml[:, 0], ml[:, 1])
    dkl_raw = ref_loss.GaussianDkl(normalize=False)(ml[:, 0], ml[:, 1])
    assert abs(omodel.gaussian_dkl(ml[:, 0], ml[:, 1], True).item() - dkl.item()) < 1e-6
    a, b = torch.randn(3, 2, 9, 7, generator=g), torch.randn(3, 2, 9, 7, generator=g)
    l2 = [ref_loss.L2Loss(c, ba)(a, b).item() for c in (False, True) for ba in (False, True)]
    assert all(abs(oloss.l2_loss(a, b, c, ba).item() - l2[2 * i + j]) < 1e-5 for i, c in enumerate((False, True)) for j, ba in enumerate((False, True)))
    # FlowParamsLoss through the reference class, with the reference-built model's inverse-flow functions (forward_controls_loss=False
    # builds the regression flow 'backwards', regression.py:179-184)
    cfg = ref_config
    _configure(cfg, 6)
    cfg.model.forward_controls_loss = False
    torch.manual_seed(0)
    _, _, _, ext = ref_build.build_extended_ae_model(cfg.model, cfg.train, ref_helper)
    cfg.model.forward_controls_loss = True
    ext.train()
    for blk in [m for m in ext.modules() if type(m).__name__ == 'ResidualBlock']:
        blk.dropout.p = 0.0                                                          # no RNG in the fixture
    crit = ref_loss.FlowParamsLoss(ref_helper, ext.ae_model.flow_inverse_function, ext.reg_model.flow_inverse_function)
    ml6, v6 = ml[:6].clone().requires_grad_(), synthetic.make_preset_targets(my_helper, 6, seed=5)
    fpl = crit(ml6, v6)
    fpl.backward()
    assert abs(oloss.flow_params_loss(ext.ae_model.flow_inverse_function, ext.reg_model.flow_inverse_function, ml6, v6).item() - fpl.item()) < 1e-6
    g_names = [n for n, p in ext.named_parameters() if p.grad is not None]
    np.savez_compressed(os.path.join(GOLDEN, 'metrics.npz'), v_in=v_in.numpy(), v_out=v_out.numpy(), qloss_mse=np.float64(qmse.item()),
                        qloss_l1=np.float64(ql1.item()), accuracy_pct=np.float64(acc), qloss_limited=np.float64(q_lim.item()),
                        accuracy_limited=np.float64(acc_lim), limited=np.asarray(lim),
                        acc_keys=np.asarray(list(acc_d.keys())), acc_vals=np.asarray(list(acc_d.values()), dtype=np.float64),
                        full_presets=full_ref.numpy(), mu_logvar=ml.numpy(), dkl=np.float64(dkl.item()), dkl_raw=np.float64(dkl_raw.item()),
                        l2_a=a.numpy(), l2_b=b.numpy(), l2=np.asarray(l2), flow_params_loss=np.float64(fpl.item()),
                        flow_params_v=v6.numpy(), flow_params_dml=ml6.grad.numpy(),
                        flow_params_grad_norms=np.asarray([float(dict(ext.named_parameters())[n].grad.double().norm()) for n in g_names]),
                        flow_params_grad_names=np.asarray(g_names))
    print("metrics: oracle == reference | QLoss %.6f / L1 %.6f, accuracy %.3f %%, FlowParamsLoss %.6f (%d parameter tensors with gradients)"
          % (qmse.item(), ql1.item(), float(acc), fpl.item(), len(g_names)))


def golden_basic_vae(ref_config, ref_build, my_helper):
    """BasicVAE (VAE.py:19-66; built when latent_flow_arch is None, build.py:45-47): reference vs oracle on the same seeds (eval forward and
    Dkl latent loss; the training forward draws eps from the global generator, so it is compared in eval mode)."""
    cfg = ref_config
    _configure(cfg, 3)
    cfg.model.latent_flow_arch = None
    cfg.model.dim_z = 256
    torch.manual_seed(0)
    enc, dec, ae = ref_build.build_ae_model(cfg.model, cfg.train)
    cfg.model.latent_flow_arch = 'realnvp_6l300'
    assert type(ae).__name__ == 'BasicVAE'
    from preset_gen_vae_b200 import config as pcfg
    m_cfg, t_cfg = pcfg.make_default(minibatch_size=3, latent_flow_arch=None, params_regression_architecture='mlp_3l1024')
    m_cfg.dim_z = 256
    torch.manual_seed(0)
    o_enc = omodel.Encoder(m_cfg.encoder_architecture, 256, m_cfg.input_tensor_size, t_cfg.fc_dropout,
                           output_bn=(t_cfg.latent_flow_input_regularization.lower() == 'bn'), deepest_features_mix=False)
    o_dec = omodel.Decoder(m_cfg.encoder_architecture, 256, m_cfg.input_tensor_size, t_cfg.fc_dropout)
    o_ae = omodel.BasicVAE(o_enc, 256, o_dec, t_cfg.normalize_losses)
    sd = ae.state_dict()
    assert list(sd.keys()) == list(o_ae.state_dict().keys())
    for k, v in o_ae.state_dict().items():
        assert torch.equal(v, sd[k]), k
    x = synthetic.make_spectrogram_like(3, 1, seed=2)
    ae.eval(); o_ae.eval()
    with torch.no_grad():
        r, o = ae(x), o_ae(x)
    for a, b in zip(o, r):
        assert a.shape == b.shape and relerr(a, b) < 1e-6
    lat = ae.latent_loss(r[0]).item()
    assert abs(o_ae.latent_loss(o[0]).item() - lat) < 1e-6 * abs(lat)
    np.savez_compressed(os.path.join(GOLDEN, 'basic_vae.npz'), z_mu_logvar=r[0].numpy(), x_out_sub=r[4].numpy()[:, :, ::4, ::4],
                        x_out_sum=np.float64(r[4].double().sum().item()), latent_loss=np.float64(lat), n_state_entries=np.int64(len(sd)))
    print("BasicVAE: oracle == reference (eval forward, Dkl latent loss %.6f, %d state entries)" % (lat, len(sd)))


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(8)
    ref_config, ref_build, ref_loss, ref_audio, ref_preset, ref_dexed = import_reference()
    golden_frontend(ref_audio)
    fake, ref_helper, my_helper = golden_index_tables(ref_preset, ref_dexed)
    golden_model(ref_config, ref_build, ref_loss, ref_helper, my_helper, 'c1_b4', B=4)
    six = ((40, 85), (50, 85), (60, 42), (60, 85), (60, 127), (70, 85))
    golden_model(ref_config, ref_build, ref_loss, ref_helper, my_helper, 'c6_b2', B=2, midi_notes=six, stack=True)
    golden_metrics(ref_loss, ref_preset, ref_build, ref_config, fake, ref_helper, my_helper)
    golden_basic_vae(ref_config, ref_build, my_helper)
    print("golden fixtures written to", GOLDEN)


if __name__ == '__main__':
    main()
