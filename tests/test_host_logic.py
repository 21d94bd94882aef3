"""CPU: host-side logic of the package (index tables, config format, ABI surface) — no GPU, no compute calls."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

from preset_gen_vae_b200 import _lib, config as pcfg, synthetic
from preset_gen_vae_b200.data import preset as ppreset
from preset_gen_vae_b200.utils import config as ucfg


def test_dexed_layout_matches_reference_fixture(golden_dir, idx_helper):
    g = json.load(open(os.path.join(golden_dir, 'dexed_layout.json')))
    assert idx_helper.learnable_preset_size == g['learnable_preset_size'] == 610
    assert idx_helper.full_to_learnable == g['full_to_learnable']
    assert idx_helper.vst_param_learnable_model == g['vst_param_learnable_model']
    assert list(idx_helper.vst_param_cardinals) == g['vst_param_cardinals']
    t = idx_helper.device_tables()
    for k, v in g['device_tables'].items():
        assert t[k].tolist() == v, k
    assert len(t['num_cols']) == 90 and len(t['grp_start']) == 54
    widths = t['grp_len'].tolist()
    assert widths[:6] == [32, 8, 2, 2, 6, 8] and widths[6:14] == [2, 32, 15, 4, 4, 8, 4, 8]


@pytest.mark.parametrize("mode,size", [('vst_cat', 224), ('all<=8', 340), (None, 144)])
def test_other_learnable_layouts(mode, size):
    assert ppreset.DexedLearnableLayout(mode).preset_indexes_helper.learnable_preset_size == size


def test_three_operators_layout():
    assert ppreset.DexedLearnableLayout('all<=32', operators=(1, 2, 3)).preset_indexes_helper.learnable_preset_size == 340


def test_useless_params_and_inference_tail(idx_helper):
    v = synthetic.make_preset_targets(idx_helper, 64, seed=0)
    silent = 0
    for row in range(64):
        num, cat = idx_helper.get_useless_learned_params_indexes(v[row])
        n_silent_ops = sum(v[row, idx_helper.full_to_learnable[31 + 22 * op]].item() < 1e-3 for op in range(6))
        assert len(num) == 12 * n_silent_ops and len(cat) == 8 * n_silent_ops
        silent += n_silent_ops
    assert silent > 0
    full = ppreset.learnable_to_full_presets(idx_helper, v, ppreset.DexedLearnableLayout().params_default_values)
    assert full.shape == (64, 155) and float(full.min()) >= 0.0 and float(full.max()) <= 1.0
    assert torch.all(full[:, [44, 66, 88, 110, 132, 154]] == 1.0) and torch.all(full[:, 3] == 0.5)


def test_config_format_round_trip(tmp_path, idx_helper):
    m, t = pcfg.make_default()
    assert m.input_tensor_size == (160, 1, 257, 347) and m.dim_z == 256 and not m.concat_midi_to_z
    pcfg.apply_dataset_dims(m, idx_helper)
    assert m.dim_z == 610 and m.learnable_params_tensor_length == 610 and m.synth_params_count == 144
    assert m.synth_args_str == 'al*_op123456_lab*' and abs(t.early_stop_lr_threshold - 2e-7) < 1e-12
    path = tmp_path / 'config.json'
    pcfg.dump_config_json(m, t, path)
    m2, t2 = ucfg.get_config_from_file(path)
    assert m2.stft_args == (1024, 256) and isinstance(m2.spectrogram_size, tuple) and t2.minibatch_size == 160
    m6, t6 = pcfg.make_default(midi_notes=((40, 85), (50, 85), (60, 42), (60, 85), (60, 127), (70, 85)),
                               stack_spectrograms=True)
    assert m6.input_tensor_size[1] == 6 and not m6.concat_midi_to_z and t6.n_epochs == 400
    m6b, t6b = pcfg.make_default(midi_notes=((40, 85), (50, 85), (60, 42), (60, 85), (60, 127), (70, 85)))
    assert m6b.concat_midi_to_z and m6b.increased_dataset_size and t6b.n_epochs == 81 and t6b.lr_warmup_epochs == 2


def test_abi_library_exports_every_declared_symbol():
    protos = _lib.parse_header()
    assert len(protos) >= 15
    cdll = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(cdll, name), name
    L = _lib.lib()
    assert L.pgv_version() == 102
    assert L.pgv_frontend_num_frames(88576, 256) == 347 and L.pgv_frontend_num_frames(88200, 256) == 345
    assert L.pgv_frontend_mel_ld(1024) == 544
    assert L.pgv_frontend_workspace_bytes(1, 88576, 1000, 256, 257) == 0      # unsupported n_fft
    assert L.pgv_frontend_workspace_bytes(8, 88576, 1024, 256, 257) > 8 * 88576 * 8
    decls = re.sub(r'/\*.*?\*/', '', open(_lib.HEADER_PATH).read(), flags=re.S)
    assert not re.search(r'torch|at::|Tensor|std::', decls)                     # plain C types only


def test_no_gpu_means_loud_failure():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.PgvError):
        _lib.handle()


def test_synthetic_inputs_are_deterministic(idx_helper):
    a, b = synthetic.make_audio(2, 1, seed=0), synthetic.make_audio(2, 1, seed=0)
    assert torch.equal(a, b) and a.shape == (2, 1, 88576) and float(a.abs().max()) < 1.5
    n1 = synthetic.make_noise(3, 610, 0.3, 0.4, seed=1)
    torch.manual_seed(1)   # same stream as the global generator the reference draws from
    assert torch.equal(n1['enc_fc_mask'], torch.empty(3, 24576).bernoulli_(0.7) / 0.7)
    assert torch.equal(n1['eps'], torch.randn(3, 610))


@pytest.mark.parametrize("arch", ["flow_realnvp_6l300", "mlp_3l1024"])
def test_state_dict_layout_matches_the_oracle_for_both_regression_heads(idx_helper, arch):
    """Checkpoint compatibility (row b of the scope table): same keys, same shapes, loadable with strict=True, for the flow
    head of the default config and for the MLP head (regression.py:61-102) - modules are only constructed here, no kernels run."""
    from oracle import model as omodel
    from preset_gen_vae_b200 import config as pcfg
    from preset_gen_vae_b200.model import build
    m_cfg, t_cfg = pcfg.make_default(minibatch_size=4, params_regression_architecture=arch)
    pcfg.apply_dataset_dims(m_cfg, idx_helper)
    ref = omodel.build_extended_ae_model(m_cfg, t_cfg, idx_helper)[3].state_dict()
    mine = build.build_extended_ae_model(m_cfg, t_cfg, idx_helper)[3]
    sd = mine.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    assert all(sd[k].shape == ref[k].shape and sd[k].dtype == ref[k].dtype for k in ref)
    mine.load_state_dict(ref, strict=True)
