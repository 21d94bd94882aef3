"""Run configuration in the reference's format (config.py:19-198): plain attribute bags `model`, `train`,
`evaluate`, derived fields recomputed by `update_dynamic_config_params()`, serialisable to the reference's
`config.json` (`{'model': model.__dict__, 'train': train.__dict__}`, logs/logger.py:160-162) and readable back with
`utils.config.get_config_from_file`.  Field names, defaults and derivation rules are the reference's; a reference
`config.py` / `config.json` can be used in place of this module unchanged.

`make_default()` returns fresh, independent (model, train) bags so tests do not share module-level state, and
`apply_dataset_dims()` performs the mutation that `data.build.get_dataset` does on the reference (data/build.py:35-39).
"""
import copy
import datetime
import json

from .utils.config import _Config


def _default_model():
    m = _Config()
    m.name = "FlVAE2"
    m.run_name = '00_debug'
    m.allow_erase_run = True
    m.encoder_architecture = 'speccnn8l1_bn'
    m.params_regression_architecture = 'flow_realnvp_6l300'
    m.params_reg_softmax = False
    m.note_duration = (3.0, 1.0)
    m.sampling_rate = 22050
    m.stft_args = (1024, 256)
    m.mel_bins = 257
    m.mel_f_limits = (0, 11050)
    m.midi_notes = ((60, 85), )
    m.stack_spectrograms = False
    m.stack_specs_deepest_features_mix = False
    m.increased_dataset_size = None
    m.spectrogram_min_dB = -120.0
    m.spectrogram_size = (257, 347)
    m.input_tensor_size = None
    m.concat_midi_to_z = None
    m.dim_z = 256
    m.latent_flow_arch = 'realnvp_6l300'
    m.forward_controls_loss = True
    m.synth = 'dexed'
    m.synth_args_str = 'al*_op*_lab*'
    m.synth_params_count = -1
    m.learnable_params_tensor_length = -1
    m.synth_vst_params_learned_as_categorical = 'all<=32'
    m.dataset_labels = None
    m.dataset_synth_args = (None, [1, 2, 3, 4, 5, 6])
    m.logs_root_dir = "saved"
    return m


def _default_train():
    t = _Config()
    t.start_datetime = datetime.datetime.now().isoformat()
    t.minibatch_size = 160
    t.main_cuda_device_idx = 1
    t.test_holdout_proportion = 0.2
    t.k_folds = 5
    t.current_k_fold = 0
    t.start_epoch = 0
    t.n_epochs = 400
    t.save_period = 50
    t.plot_period = 20
    t.latent_loss = 'Dkl'
    t.latent_flow_input_regularization = 'bn'
    t.params_cat_bceloss = False
    t.params_cat_softmax_temperature = 0.2
    t.normalize_losses = True
    t.optimizer = 'Adam'
    t.initial_learning_rate = 2e-4
    t.lr_warmup_epochs = 6
    t.lr_warmup_start_factor = 0.1
    t.adam_betas = (0.9, 0.999)
    t.weight_decay = 1e-4
    t.fc_dropout = 0.3
    t.reg_fc_dropout = 0.4
    t.beta = 0.2
    t.beta_start_value = 0.1
    t.beta_warmup_epochs = 25
    t.beta_cycle_epochs = -1
    t.scheduler_name = 'ReduceLROnPlateau'
    t.scheduler_loss = ('ReconsLoss/Backprop', 'Controls/BackpropLoss')
    t.scheduler_lr_factor = 0.2
    t.scheduler_patience = 6
    t.scheduler_cooldown = 6
    t.scheduler_threshold = 1e-4
    t.early_stop_lr_threshold = None
    t.verbosity = 1
    t.init_security_pause = 0.0
    t.logged_samples_count = 4
    t.profiler_args = {'enabled': False, 'use_cuda': True, 'record_shapes': False,
                       'profile_memory': False, 'with_stack': False}
    t.profiler_full_trace = False
    t.profiler_1_GPU = False
    return t


def update_dynamic_config_params(model_config=None, train_config=None):
    """Derivation rules of config.py:148-198, applied to the given bags (default: this module's globals)."""
    m = model if model_config is None else model_config
    t = train if train_config is None else train_config
    m.stack_spectrograms = m.stack_spectrograms and (len(m.midi_notes) > 1)
    m.increased_dataset_size = (len(m.midi_notes) > 1) and not m.stack_spectrograms
    m.concat_midi_to_z = (len(m.midi_notes) > 1) and not m.stack_spectrograms
    m.input_tensor_size = (t.minibatch_size, 1 if not m.stack_spectrograms else len(m.midi_notes),
                           m.spectrogram_size[0], m.spectrogram_size[1])
    t.early_stop_lr_threshold = t.initial_learning_rate * 1e-3
    t.logged_samples_count = max(t.logged_samples_count, len(m.midi_notes))
    if m.dataset_synth_args[0] is not None:
        t.n_epochs, t.lr_warmup_epochs, t.scheduler_patience, t.scheduler_cooldown, t.beta_warmup_epochs \
            = 700, 10, 10, 10, 40
    if m.increased_dataset_size:
        n = len(m.midi_notes) - 1
        t.n_epochs = 1 + t.n_epochs // n
        t.lr_warmup_epochs = 1 + t.lr_warmup_epochs // n
        t.scheduler_patience = 1 + t.scheduler_patience // n
        t.scheduler_cooldown = 1 + t.scheduler_cooldown // n
        t.beta_warmup_epochs = 1 + t.beta_warmup_epochs // n
    if m.synth == "dexed":
        if m.dataset_synth_args[0] is not None:
            m.synth_args_str = m.synth_args_str.replace("al*", "al" + '.'.join(str(a) for a in m.dataset_synth_args[0]))
        if m.dataset_synth_args[1] is not None:
            m.synth_args_str = m.synth_args_str.replace("_op*", "_op" + ''.join(str(o) for o in m.dataset_synth_args[1]))
        if m.dataset_labels is not None:
            m.synth_args_str = m.synth_args_str.replace("_lab*", '_' + '_'.join(l[0:4] for l in m.dataset_labels))
    else:
        raise NotImplementedError("Unknown synth prefix for model.synth '{}'".format(m.synth))


def make_default(minibatch_size=None, midi_notes=None, stack_spectrograms=None, **model_overrides):
    """Fresh (model, train) bags with the reference defaults, optional overrides, dynamic fields updated."""
    m, t = _default_model(), _default_train()
    if minibatch_size is not None:
        t.minibatch_size = minibatch_size
    if midi_notes is not None:
        m.midi_notes = tuple(tuple(n) for n in midi_notes)
    if stack_spectrograms is not None:
        m.stack_spectrograms = stack_spectrograms
    for k, v in model_overrides.items():
        setattr(m, k, v)
    update_dynamic_config_params(m, t)
    return m, t


def apply_dataset_dims(model_config, idx_helper):
    """What data.build.get_dataset writes into the model config (data/build.py:35-39)."""
    model_config.synth_params_count = sum(1 for mdl in idx_helper.vst_param_learnable_model if mdl is not None)
    model_config.learnable_params_tensor_length = idx_helper.learnable_preset_size
    if model_config.params_regression_architecture.startswith("flow_"):
        model_config.dim_z = model_config.learnable_params_tensor_length
    return model_config


def dump_config_json(model_config, train_config, path):
    """Same file the reference's RunLogger writes (logs/logger.py:160-162)."""
    with open(path, 'w') as f:
        json.dump({'model': copy.deepcopy(model_config.__dict__), 'train': copy.deepcopy(train_config.__dict__)}, f)


model = _default_model()
train = _default_train()
evaluate = _Config()
evaluate.epoch = -1
update_dynamic_config_params()
