"""VAE wrappers with the reference's interface (model/VAE.py:19-193): `BasicVAE` and `FlowVAE`.

`FlowVAE.forward(x, sample_info)` returns `(z_0_mu_logvar [B,2,D], z_0_sampled [B,D], z_K_sampled [B,D],
log_abs_det_jac [B], x_out [B,C,257,347])`; `latent_loss(...)` the normalised negative ELBO latent terms.
Reparameterisation and the latent loss are libpgv.so kernels; eps is drawn from torch's CUDA generator exactly where
the reference draws it (VAE.py:172-173) unless given explicitly (parity tests).
"""
import numpy as np
import torch
import torch.nn as nn

from . import loss, ops
from .flows import SimpleRealNVP
from .program import run_program


class _Reparam:
    def prog_fwd(self, inputs, training, extra):
        ml, eps = inputs[0].contiguous(), inputs[1]
        return ops.reparam_fwd(ml, eps), (ml, eps)

    def prog_bwd(self, dz, ctx, grads, needs):
        ml, eps = ctx
        return ops.reparam_bwd(dz, ml, eps), None


_REPARAM = _Reparam()


def reparametrize(z_mu_logvar, eps):
    """z = mu + exp(logvar/2) * eps; eps None => z = mu (eval mode, VAE.py:175-176)."""
    return run_program(_REPARAM, (z_mu_logvar, eps), [], True)


class _MidiConcat:
    """z_0_mu_logvar[:, :, 2:] = encoder output; dims 0-1 carry MIDI pitch / velocity (VAE.py:155-165)."""

    def prog_fwd(self, inputs, training, extra):
        enc, head = inputs
        return torch.cat([head, enc], dim=2).contiguous(), None

    def prog_bwd(self, d, ctx, grads, needs):
        return d[:, :, 2:].contiguous(), None


class BasicVAE(nn.Module):
    """Standard VAE without latent flow (VAE.py:19-66).  As in the reference, `forward(x)` takes one argument, so it
    cannot be driven through ExtendedAE.forward(x, sample_info); kept for API / isinstance parity."""

    def __init__(self, encoder, dim_z, decoder, normalize_latent_loss, latent_loss_type):
        super().__init__()
        self.encoder = encoder
        self.dim_z = dim_z
        self.decoder = decoder
        self.is_profiled = False
        if latent_loss_type.lower() == 'dkl':
            self.latent_criterion = loss.GaussianDkl(normalize=normalize_latent_loss)
        else:
            raise NotImplementedError("Latent loss '{}' unavailable".format(latent_loss_type))

    def forward(self, x, noise=None):
        z_mu_logvar = self.encoder(x, None if noise is None else noise.get('enc_fc_mask'))
        n = z_mu_logvar.shape[0]
        if self.training:
            eps = torch.randn(n, self.dim_z, device=x.device) if noise is None else noise['eps']
            z_sampled = reparametrize(z_mu_logvar, eps)
        else:
            z_sampled = reparametrize(z_mu_logvar, None)
        x_out = self.decoder(z_sampled, None if noise is None else noise.get('dec_fc_mask'))
        return z_mu_logvar, z_sampled, z_sampled, torch.zeros((n, 1), device=x.device), x_out

    def latent_loss(self, z_0_mu_logvar, **kwargs):
        return self.latent_criterion(z_0_mu_logvar[:, 0, :], z_0_mu_logvar[:, 1, :], packed=z_0_mu_logvar)


class FlowVAE(nn.Module):
    def __init__(self, encoder, dim_z, decoder, normalize_latent_loss: bool, flow_arch: str, concat_midi_to_z0=False):
        super().__init__()
        self.encoder = encoder
        self.dim_z = dim_z
        self.concat_midi_to_z0 = concat_midi_to_z0
        self.decoder = decoder
        self.is_profiled = False
        self.normalize_latent_loss = normalize_latent_loss
        flow_args = flow_arch.split('_')
        if len(flow_args) < 2:
            raise AssertionError("flow_arch argument must contains at least a flow type and layers description, "
                                 "e.g. 'realnvp_4l200'")
        elif len(flow_args) > 2:
            raise NotImplementedError("Optional flow arch argument not supported yet")
        self.flow_arch = flow_args[0]
        flow_layers_args = flow_args[1].split('l')
        self.flow_layers_count = int(flow_layers_args[0])
        self.flow_hidden_features = int(flow_layers_args[1])
        if self.flow_arch.lower() == 'realnvp':
            flow = SimpleRealNVP(features=self.dim_z, hidden_features=self.flow_hidden_features,
                                 num_layers=self.flow_layers_count, num_blocks_per_layer=2,
                                 batch_norm_within_layers=True, batch_norm_between_layers=False)
            self.flow_transform = flow._transform
        elif self.flow_arch.lower() == 'maf':
            raise NotImplementedError("'maf' latent flows are out of scope (reference: 'very unstable', regression.py:160-163)")
        else:
            raise NotImplementedError("Unavailable flow '{}'".format(self.flow_arch))

    @property
    def flow_forward_function(self):
        return self.flow_transform.forward

    @property
    def flow_inverse_function(self):
        return self.flow_transform.inverse

    def forward(self, x, sample_info=None, noise=None):
        """noise: optional dict with 'enc_fc_mask', 'eps', 'dec_fc_mask' (tests); None => drawn like the reference."""
        n_minibatch = x.size()[0]
        enc_out = self.encoder(x, None if noise is None else noise.get('enc_fc_mask'))
        if not self.concat_midi_to_z0:
            z_0_mu_logvar = enc_out
        else:
            head = torch.zeros((n_minibatch, 2, 2), device=x.device)
            if sample_info is not None:
                head[:, 0, :] = -1.0 + 2.0 * sample_info[:, [1, 2]].float() / 127.0
                head[:, 1, :] = float(np.log(4.0 / (127 ** 2)))
            z_0_mu_logvar = run_program(_MidiConcat(), (enc_out, head), [], True)
        if self.training:
            eps = torch.randn(n_minibatch, self.dim_z, device=x.device) if noise is None else noise['eps']
            z_0_sampled = reparametrize(z_0_mu_logvar, eps)
        else:
            z_0_sampled = reparametrize(z_0_mu_logvar, None)
        z_K_sampled, log_abs_det_jac = self.flow_transform(z_0_sampled)
        dec_mask = None if noise is None else noise.get('dec_fc_mask')
        side = getattr(self, 'decoder_stream', None)
        if side is None:
            x_out = self.decoder(z_K_sampled, dec_mask)
        else:
            # The decoder and whatever the caller does next with z_K (the regression flow, train.py:218) are independent:
            # enqueue the decoder on a side stream so both run concurrently.  The caller must make its stream wait for
            # `decoder_stream` before it reads x_out.  (autograd runs each backward node on its forward stream, so the two
            # backward branches overlap as well.)
            main = torch.cuda.current_stream(x.device)
            side.wait_stream(main)
            z_K_sampled.record_stream(side)
            with torch.cuda.stream(side):
                x_out = self.decoder(z_K_sampled, dec_mask)
            x_out.record_stream(main)
        return z_0_mu_logvar, z_0_sampled, z_K_sampled, log_abs_det_jac, x_out

    def latent_loss(self, z_0_mu_logvar, z_0_sampled, z_K_sampled, log_abs_det_jac):
        return loss.flow_latent_loss(z_0_mu_logvar, z_0_sampled, z_K_sampled, log_abs_det_jac, self.normalize_latent_loss)
