"""Model builders with the reference's signatures and config handling (model/build.py:11-122)."""
from . import VAE, encoder, decoder, extendedAE, regression


def build_encoder_and_decoder_models(model_config, train_config):
    if not hasattr(model_config, 'stack_specs_deepest_features_mix'):
        model_config.stack_specs_deepest_features_mix = True             # build.py:13-14 backward compatibility
    force_bigger_network = ((len(model_config.midi_notes) > 1) and not model_config.stack_spectrograms)
    enc_z_length = (model_config.dim_z - 2 if model_config.concat_midi_to_z else model_config.dim_z)
    encoder_model = encoder.SpectrogramEncoder(
        model_config.encoder_architecture, enc_z_length, model_config.input_tensor_size, train_config.fc_dropout,
        output_bn=(train_config.latent_flow_input_regularization.lower() == 'bn'),
        deepest_features_mix=model_config.stack_specs_deepest_features_mix, force_bigger_network=force_bigger_network)
    decoder_model = decoder.SpectrogramDecoder(model_config.encoder_architecture, model_config.dim_z,
                                               model_config.input_tensor_size, train_config.fc_dropout,
                                               force_bigger_network=force_bigger_network)
    return encoder_model, decoder_model


def build_ae_model(model_config, train_config):
    """Returns (encoder, decoder, full AE model)."""
    encoder_model, decoder_model = build_encoder_and_decoder_models(model_config, train_config)
    if model_config.latent_flow_arch is None:
        ae_model = VAE.BasicVAE(encoder_model, model_config.dim_z, decoder_model, train_config.normalize_losses,
                                train_config.latent_loss)
    else:
        ae_model = VAE.FlowVAE(encoder_model, model_config.dim_z, decoder_model, train_config.normalize_losses,
                               model_config.latent_flow_arch, concat_midi_to_z0=model_config.concat_midi_to_z)
    return encoder_model, decoder_model, ae_model


def build_extended_ae_model(model_config, train_config, idx_helper):
    """Returns (encoder, decoder, ae_model, extended_ae_model)."""
    encoder_model, decoder_model, ae_model = build_ae_model(model_config, train_config)
    if not hasattr(model_config, 'params_reg_softmax'):
        model_config.params_reg_softmax = True                            # build.py:61-62 legacy default
    if model_config.params_regression_architecture.startswith("mlp_"):
        assert model_config.forward_controls_loss is True
        reg_arch = model_config.params_regression_architecture.replace("mlp_", "")
        reg_model = regression.MLPRegression(reg_arch, model_config.dim_z, idx_helper, train_config.reg_fc_dropout,
                                             cat_softmax_activation=model_config.params_reg_softmax)
    elif model_config.params_regression_architecture.startswith("flow_"):
        assert model_config.learnable_params_tensor_length > 0
        reg_arch = model_config.params_regression_architecture.replace("flow_", "")
        reg_model = regression.FlowRegression(reg_arch, model_config.dim_z, idx_helper,
                                              fast_forward_flow=model_config.forward_controls_loss,
                                              dropout_p=train_config.reg_fc_dropout,
                                              cat_softmax_activation=model_config.params_reg_softmax)
    else:
        raise NotImplementedError("Synth param regression arch '{}' not implemented"
                                  .format(model_config.params_regression_architecture))
    extended_ae_model = extendedAE.ExtendedAE(ae_model, reg_model, idx_helper, train_config.fc_dropout)
    return encoder_model, decoder_model, ae_model, extended_ae_model


def _is_attr_equal(attr1, attr2):
    _attr1 = tuple(attr1) if isinstance(attr1, list) else attr1
    _attr2 = tuple(attr2) if isinstance(attr2, list) else attr2
    return _attr1 == _attr2


def check_configs_on_resume_from_checkpoint(new_model_config, new_train_config, config_json_checkpoint):
    """Raises ValueError if the saved config.json and the new config disagree (build.py:90-122)."""
    for section, new_cfg, attrs in (
            ('model', new_model_config, ['name', 'run_name', 'encoder_architecture', 'dim_z', 'concat_midi_to_z',
                                         'latent_flow_arch', 'logs_root_dir', 'note_duration', 'stack_spectrograms',
                                         'increased_dataset_size', 'stft_args', 'spectrogram_size', 'mel_bins']),
            ('train', new_train_config, ['minibatch_size', 'test_holdout_proportion', 'normalize_losses', 'optimizer',
                                         'scheduler_name'])):
        prev = config_json_checkpoint[section]
        for attr in attrs:
            if not _is_attr_equal(prev[attr], new_cfg.__dict__[attr]):
                raise ValueError("{} attribute '{}' is different in the new config.py ({}) and the old config.json ({})"
                                 .format(section.capitalize(), attr, new_cfg.__dict__[attr], prev[attr]))
