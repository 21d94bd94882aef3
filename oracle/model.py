"""CPU oracle of the spectral-VAE model path.  TEST INFRASTRUCTURE (see oracle/__init__.py).

A plain-PyTorch restatement of the reference modules, table-driven, with the SAME submodule names (so the
state_dict keys and the RNG consumption order at construction are those of the reference) and one addition:
`forward(..., noise=...)` accepts the eps / dropout masks explicitly so that the oracle and the CUDA path can
be fed identical randomness.  With `noise=None` it draws from torch's global generator in the reference's order.

Reference files restated:
  model/layer.py:10-46        Conv2D / TConv2D = conv -> LeakyReLU(0.1) -> BatchNorm2d
  model/encoder.py:23-108     SpectrogramEncoder, :233-259 'speccnn8l1_bn' CNN
  model/decoder.py:9-92       SpectrogramDecoder, :199-220 'speccnn8l1_bn' CNN
  model/VAE.py:69-193         FlowVAE (forward, latent_loss); :19-66 BasicVAE
  model/flows.py:42-90        CustomRealNVP
  model/regression.py:20-189  PresetActivation, FlowRegression, MLPRegression
  model/extendedAE.py:13-51   ExtendedAE
  model/build.py:11-80        build_* functions
Only the in-scope configuration space is covered ('speccnn8l1_bn', RealNVP flows).
"""
import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

from . import nflows_port as nf


def _conv_block(prefix, cin, cout, k, stride, pad, bn=True):
    seq = nn.Sequential()
    seq.add_module(prefix + 'conv', nn.Conv2d(cin, cout, k, stride, pad))
    seq.add_module(prefix + 'act', nn.LeakyReLU(0.1))
    if bn:
        seq.add_module(prefix + 'bn', nn.BatchNorm2d(cout))
    return seq


def _tconv_block(prefix, cin, cout, k, stride, pad, out_pad):
    seq = nn.Sequential()
    seq.add_module(prefix + 'tconv', nn.ConvTranspose2d(cin, cout, k, stride, pad, out_pad))
    seq.add_module(prefix + 'act', nn.LeakyReLU(0.1))
    seq.add_module(prefix + 'bn', nn.BatchNorm2d(cout))
    return seq


class EncoderCNN(nn.Module):                                   # encoder.py:233-259
    CHANNELS = (1, 8, 16, 32, 64, 128, 256)

    def __init__(self, last_layers_to_remove):
        super().__init__()
        c = self.CHANNELS
        self.enc_nn = nn.Sequential(_conv_block('enc1', c[0], c[1], 5, 2, 2, bn=False),
                                    *[_conv_block('enc%d' % (i + 1), c[i], c[i + 1], 4, 2, 2) for i in range(1, 6)])
        if last_layers_to_remove <= 1:
            self.enc_nn.add_module('4x4conv', _conv_block('enc7', 256, 512, 4, 2, 2))
        if last_layers_to_remove == 0:
            self.enc_nn.add_module('1x1conv', _conv_block('enc8', 512, 1024, 1, 1, 0, bn=False))

    def forward(self, x):
        return self.enc_nn(x)


class Encoder(nn.Module):                                      # encoder.py:23-108
    def __init__(self, architecture, dim_z, input_tensor_size, fc_dropout, output_bn=False,
                 deepest_features_mix=True, force_bigger_network=False):
        super().__init__()
        assert architecture == 'speccnn8l1_bn'
        self.dim_z = dim_z
        self.spectrogram_channels = C = input_tensor_size[1]
        mix_ch = 1024 if C > 1 else 2048
        self.single_ch_cnn = EncoderCNN(1 if deepest_features_mix else 2)
        if deepest_features_mix:
            self.features_mixer_cnn = _conv_block('enc8', 512 * C, mix_ch, 1, 1, 0, bn=False)
        else:
            n4 = 1800 if force_bigger_network else (512 if C == 1 else 768)
            self.features_mixer_cnn = nn.Sequential(_conv_block('enc7', 256 * C, n4, 4, 2, 2),
                                                    _conv_block('enc8', n4, mix_ch, 1, 1, 0, bn=False))
        with torch.no_grad():                                  # encoder.py:73-78: training-mode dummy forward
            shape = list(input_tensor_size)
            shape[0] = 1
            self.cnn_out_size = self._forward_cnns(torch.zeros(shape)).size()
        n_items = self.cnn_out_size[1] * self.cnn_out_size[2] * self.cnn_out_size[3]
        self.mlp = nn.Sequential(nn.Dropout(fc_dropout), nn.Linear(n_items, 2 * dim_z))
        if output_bn:
            self.mlp.add_module('lat_in_regularization', nn.BatchNorm1d(2 * dim_z))

    def _forward_cnns(self, x):
        per_channel = [self.single_ch_cnn(x[:, ch:ch + 1]) for ch in range(self.spectrogram_channels)]
        return self.features_mixer_cnn(torch.cat(per_channel, dim=1))

    def forward(self, x, dropout_mask=None):
        h = self._forward_cnns(x).reshape(x.shape[0], -1)
        if dropout_mask is None:
            h = self.mlp[0](h)
        elif self.training:
            h = h * dropout_mask
        h = self.mlp[1](h)
        if hasattr(self.mlp, 'lat_in_regularization'):
            h = self.mlp.lat_in_regularization(h)
        return h.reshape(x.shape[0], 2, self.dim_z)


class DecoderCNN(nn.Module):                                   # decoder.py:199-220
    SPEC = ((512, 256, (1, 1)), (256, 128, (1, 0)), (128, 64, (1, 1)), (64, 32, (1, 1)), (32, 16, (1, 0)),
            (16, 8, (1, 0)))

    def __init__(self, force_bigger_network=False):
        super().__init__()
        blocks = []
        for i, (cin, cout, op) in enumerate(self.SPEC):
            if i == 0 and force_bigger_network:
                cin = 1800
            blocks.append(_tconv_block('dec%d' % (i + 2), cin, cout, 4, 2, 2, op))
        self.dec_nn = nn.Sequential(*blocks, nn.ConvTranspose2d(8, 1, 5, 2, 2), nn.Hardtanh())

    def forward(self, x):
        return self.dec_nn(x)


class Decoder(nn.Module):                                      # decoder.py:9-92
    def __init__(self, architecture, dim_z, output_tensor_size, fc_dropout, force_bigger_network=False):
        super().__init__()
        assert architecture == 'speccnn8l1_bn'
        assert tuple(output_tensor_size[2:]) == (257, 347)
        self.spectrogram_channels = output_tensor_size[1]
        self.cnn_input_shape = (2048, 3, 4)
        self.last_4x4conv_ch = 1800 if force_bigger_network else 512
        self.mlp = nn.Sequential(nn.Linear(dim_z, int(np.prod(self.cnn_input_shape))), nn.Dropout(fc_dropout))
        self.features_unmixer_cnn = _tconv_block('dec1', 2048, self.spectrogram_channels * self.last_4x4conv_ch,
                                                 1, 1, 0, 0)
        self.single_ch_cnn = DecoderCNN(force_bigger_network)

    def forward(self, z, dropout_mask=None):
        h = self.mlp[0](z)
        if dropout_mask is None:
            h = self.mlp[1](h)
        elif self.training:
            h = h * dropout_mask
        h = self.features_unmixer_cnn(h.view(-1, *self.cnn_input_shape))
        outs = [self.single_ch_cnn(part) for part in torch.split(h, self.last_4x4conv_ch, dim=1)]
        return torch.cat(outs, dim=1)


_LOG_2PI = np.log(2 * np.pi)


def standard_gaussian_log_probability(samples):                # utils/probability.py:13-18
    return -0.5 * (samples.shape[1] * _LOG_2PI + torch.sum(samples ** 2, dim=1))


def gaussian_log_probability(samples, mu, log_var):            # utils/probability.py:21-29
    return -0.5 * (samples.shape[1] * _LOG_2PI
                   + torch.sum(log_var + ((samples - mu) ** 2 / torch.exp(log_var)), dim=1))


def gaussian_dkl(mu, logvar, normalize=True):                  # model/loss.py:46-66
    dkl = 0.5 * torch.sum(torch.exp(logvar) + torch.square(mu) - logvar - 1.0) / mu.size(0)
    return dkl / mu.size(1) if normalize else dkl


def _parse_flow_arch(arch):
    kind, layers = arch.split('_')
    n_layers, hidden = layers.split('l')
    assert kind.lower() in ('realnvp', 'rnvp')
    return int(n_layers), int(hidden)


class FlowVAE(nn.Module):                                      # VAE.py:69-193
    def __init__(self, encoder, dim_z, decoder, normalize_latent_loss, flow_arch, concat_midi_to_z0=False):
        super().__init__()
        self.encoder, self.dim_z, self.decoder = encoder, dim_z, decoder
        self.concat_midi_to_z0 = concat_midi_to_z0
        self.normalize_latent_loss = normalize_latent_loss
        n_layers, hidden = _parse_flow_arch(flow_arch)
        self.flow_transform = nf.SimpleRealNVP(dim_z, hidden, n_layers, 2, batch_norm_within_layers=True,
                                               batch_norm_between_layers=False)._transform

    def forward(self, x, sample_info=None, noise=None):
        B = x.shape[0]
        enc = self.encoder(x, None if noise is None else noise['enc_fc_mask'])
        if not self.concat_midi_to_z0:
            z0_mu_logvar = enc
        else:                                                  # VAE.py:155-165
            z0_mu_logvar = enc.new_empty((B, 2, self.dim_z))
            z0_mu_logvar[:, :, 2:] = enc
            if sample_info is None:
                z0_mu_logvar[:, :, [0, 1]] = 0.0
            else:
                z0_mu_logvar[:, 0, [0, 1]] = (-1.0 + 2.0 * sample_info[:, [1, 2]].float() / 127.0).to(enc.dtype)
                z0_mu_logvar[:, 1, [0, 1]] = np.log(4.0 / (127 ** 2))
        mu0 = z0_mu_logvar[:, 0, :]
        sigma0 = torch.exp(z0_mu_logvar[:, 1, :] / 2.0)
        if self.training:
            eps = torch.normal(torch.zeros(B, self.dim_z, device=mu0.device), torch.ones(B, self.dim_z, device=mu0.device)).to(mu0.dtype) if noise is None \
                else noise['eps'].to(mu0.dtype)
            z0 = mu0 + sigma0 * eps
        else:
            z0 = mu0
        zK, logdet = self.flow_transform(z0)
        x_out = self.decoder(zK, None if noise is None else noise['dec_fc_mask'])
        return z0_mu_logvar, z0, zK, logdet, x_out

    def latent_loss(self, z0_mu_logvar, z0, zK, logdet):       # VAE.py:183-193
        log_q = gaussian_log_probability(z0, z0_mu_logvar[:, 0, :], z0_mu_logvar[:, 1, :])
        log_p = standard_gaussian_log_probability(zK)
        loss = -(log_p - log_q + logdet).mean()
        return loss / z0.shape[1] if self.normalize_latent_loss else loss


class BasicVAE(nn.Module):                                     # VAE.py:19-66
    def __init__(self, encoder, dim_z, decoder, normalize_latent_loss, latent_loss_type='Dkl'):
        super().__init__()
        assert latent_loss_type.lower() == 'dkl'
        self.encoder, self.dim_z, self.decoder = encoder, dim_z, decoder
        self.normalize_latent_loss = normalize_latent_loss

    def forward(self, x, noise=None):
        ml = self.encoder(x, None if noise is None else noise['enc_fc_mask'])
        mu, sigma = ml[:, 0, :], torch.exp(ml[:, 1, :] / 2.0)
        if self.training:
            eps = torch.randn(x.shape[0], self.dim_z, device=mu.device).to(mu.dtype) if noise is None else noise['eps'].to(mu.dtype)
            z = mu + sigma * eps
        else:
            z = mu
        x_out = self.decoder(z, None if noise is None else noise['dec_fc_mask'])
        return ml, z, z, torch.zeros((z.shape[0], 1), device=x.device), x_out

    def latent_loss(self, z_0_mu_logvar, **kwargs):
        return gaussian_dkl(z_0_mu_logvar[:, 0, :], z_0_mu_logvar[:, 1, :], self.normalize_latent_loss)


class CustomRealNVP(nf.CompositeTransform):                    # flows.py:42-90
    def __init__(self, features, hidden_features, num_layers, num_blocks_per_layer, dropout_probability=0.0,
                 batch_norm_within_layers=False, batch_norm_between_layers=False):
        mask = torch.ones(features)
        mask[::2] = -1
        layers = []
        for layer in range(num_layers):
            p = dropout_probability if layer < num_layers - 2 else 0.0

            def create_resnet(n_in, n_out, p=p):
                return nf.ResidualNet(n_in, n_out, hidden_features=hidden_features, num_blocks=num_blocks_per_layer,
                                      dropout_probability=p, use_batch_norm=batch_norm_within_layers)
            layers.append(nf.AffineCouplingTransform(mask=mask, transform_net_create_fn=create_resnet))
            mask *= -1
            if batch_norm_between_layers and layer < num_layers - 2:
                layers.append(nf.BatchNorm(features=features))
        super().__init__(layers)


class _MaskedDropout(nn.Module):
    """Stand-in for nn.Dropout inside a ResidualBlock while explicit masks are injected."""

    def __init__(self, mask):
        super().__init__()
        self.mask = mask

    def forward(self, x):
        return x * self.mask.to(x.dtype) if self.training else x


class PresetActivation(nn.Module):                             # regression.py:20-53
    def __init__(self, idx_helper, cat_softmax_activation=False):
        super().__init__()
        self.cat_softmax_activation = cat_softmax_activation
        self.num_indexes = idx_helper.get_numerical_learnable_indexes()
        self.cat_indexes = idx_helper.get_categorical_learnable_indexes()

    def forward(self, x):
        if not self.cat_softmax_activation:
            return F.hardtanh(x, 0.0, 1.0)
        out = x.clone()
        out[:, self.num_indexes] = F.hardtanh(x[:, self.num_indexes], 0.0, 1.0)
        for cols in self.cat_indexes:
            out[:, cols] = torch.softmax(x[:, cols], dim=-1)
        return out


class FlowRegression(nn.Module):                               # regression.py:105-189
    def __init__(self, architecture, dim_z, idx_helper, dropout_p=0.0, fast_forward_flow=True,
                 cat_softmax_activation=False):
        super().__init__()
        self.dim_z = dim_z
        self._fast_forward_flow = fast_forward_flow
        n_layers, hidden = _parse_flow_arch(architecture)
        self._forward_flow_transform = CustomRealNVP(dim_z, hidden, n_layers, 2, dropout_probability=dropout_p,
                                                     batch_norm_between_layers=True, batch_norm_within_layers=True)
        self.activation_layer = PresetActivation(idx_helper, cat_softmax_activation)

    @property
    def flow_forward_function(self):
        t = self._forward_flow_transform
        return t.forward if self._fast_forward_flow else t.inverse

    @property
    def flow_inverse_function(self):
        t = self._forward_flow_transform
        return t.inverse if self._fast_forward_flow else t.forward

    def forward(self, zK, noise=None):
        if noise is None:
            v, _ = self.flow_forward_function(zK)
            return self.activation_layer(v)
        couplings = [t for t in self._forward_flow_transform._transforms if isinstance(t, nf.AffineCouplingTransform)]
        saved = []
        for layer, t in enumerate(couplings):
            for b, block in enumerate(t.transform_net.blocks):
                saved.append((block, block.dropout))
                block.dropout = _MaskedDropout(noise['reg_masks'][layer][b])
        try:
            v, _ = self.flow_forward_function(zK)
        finally:
            for block, drop in saved:
                block.dropout = drop
        return self.activation_layer(v)


class MLPRegression(nn.Module):                                # regression.py:61-102
    def __init__(self, architecture, dim_z, idx_helper, dropout_p=0.0, cat_softmax_activation=False):
        super().__init__()
        n_hidden, width = (int(v) for v in architecture.split('l'))
        self.reg_model = nn.Sequential()
        for l in range(n_hidden):
            self.reg_model.add_module('fc%d' % (l + 1), nn.Linear(dim_z if l == 0 else width, width))
            if l < n_hidden - 1:
                self.reg_model.add_module('bn%d' % (l + 1), nn.BatchNorm1d(width))
                self.reg_model.add_module('drp%d' % (l + 1), nn.Dropout(dropout_p))
            self.reg_model.add_module('act%d' % (l + 1), nn.ReLU())
        self.reg_model.add_module('fc%d' % (n_hidden + 1), nn.Linear(width, idx_helper.learnable_preset_size))
        self.reg_model.add_module('act', PresetActivation(idx_helper, cat_softmax_activation))

    def forward(self, zK, noise=None):
        return self.reg_model(zK)


class ExtendedAE(nn.Module):                                   # extendedAE.py:13-51
    def __init__(self, ae_model, reg_model, idx_helper, dropout_p=0.0):
        super().__init__()
        self.idx_helper = idx_helper
        self.ae_model = ae_model
        self.reg_model = reg_model

    def forward(self, x, sample_info=None, noise=None):
        return self.ae_model(x, sample_info, noise)

    def latent_loss(self, *args):
        return self.ae_model.latent_loss(*args)


def build_extended_ae_model(model_config, train_config, idx_helper):      # build.py:11-80
    deepest = getattr(model_config, 'stack_specs_deepest_features_mix', True)
    bigger = (len(model_config.midi_notes) > 1) and not model_config.stack_spectrograms
    enc_z = model_config.dim_z - 2 if model_config.concat_midi_to_z else model_config.dim_z
    enc = Encoder(model_config.encoder_architecture, enc_z, model_config.input_tensor_size, train_config.fc_dropout,
                  output_bn=(train_config.latent_flow_input_regularization.lower() == 'bn'),
                  deepest_features_mix=deepest, force_bigger_network=bigger)
    dec = Decoder(model_config.encoder_architecture, model_config.dim_z, model_config.input_tensor_size,
                  train_config.fc_dropout, force_bigger_network=bigger)
    assert model_config.latent_flow_arch is not None            # BasicVAE unreachable via ExtendedAE
    ae = FlowVAE(enc, model_config.dim_z, dec, train_config.normalize_losses, model_config.latent_flow_arch,
                 concat_midi_to_z0=model_config.concat_midi_to_z)
    softmax = getattr(model_config, 'params_reg_softmax', True)
    arch = model_config.params_regression_architecture
    if arch.startswith('mlp_'):
        reg = MLPRegression(arch.replace('mlp_', ''), model_config.dim_z, idx_helper, train_config.reg_fc_dropout,
                            cat_softmax_activation=softmax)
    elif arch.startswith('flow_'):
        assert model_config.learnable_params_tensor_length > 0
        reg = FlowRegression(arch.replace('flow_', ''), model_config.dim_z, idx_helper,
                             fast_forward_flow=model_config.forward_controls_loss,
                             dropout_p=train_config.reg_fc_dropout, cat_softmax_activation=softmax)
    else:
        raise NotImplementedError(arch)
    return enc, dec, ae, ExtendedAE(ae, reg, idx_helper, train_config.fc_dropout)
