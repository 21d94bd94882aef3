"""Training criteria with the reference's interface (model/loss.py:15-183) plus the two criteria train.py builds from
torch / the models (nn.MSELoss at train.py:103-104, FlowVAE.latent_loss).  Each call is one autograd node; values and
gradients are computed by libpgv.so kernels and the scalar stays on the device (no host synchronisation).
"""
import torch

from . import ops
from .program import run_program


class _SqErr:
    def prog_fwd(self, inputs, training, scale):
        a, b = inputs[0].contiguous(), inputs[1].contiguous()
        return ops.sqerr_fwd(a, b, scale).view(()), (a, b, scale)

    def prog_bwd(self, dout, ctx, grads, needs):
        a, b, scale = ctx
        g = dout.reshape(1)
        da = ops.sqerr_bwd(a, b, scale, g) if needs[0] else None
        db = ops.sqerr_bwd(b, a, scale, g) if needs[1] else None
        return da, db


_SQERR = _SqErr()


class MSELoss:
    """nn.MSELoss(reduction='mean') replacement (train.py:103-104)."""

    def __init__(self, reduction='mean'):
        assert reduction == 'mean'

    def __call__(self, inferred, target):
        return run_program(_SQERR, (inferred, target), [], True, 1.0 / inferred.numel())


class L2Loss:
    """model/loss.py:15-43."""

    def __init__(self, contents_average=False, batch_average=True):
        self.contents_average = contents_average
        self.batch_average = batch_average

    def __call__(self, inferred, target):
        scale = 1.0
        if self.batch_average:
            scale /= inferred.shape[0]
        if self.contents_average:
            scale /= inferred[0, :].numel()
        return run_program(_SQERR, (inferred, target), [], True, scale)


class _Dkl:
    def prog_fwd(self, inputs, training, normalize):
        ml = inputs[0].contiguous()
        return ops.dkl_fwd(ml, normalize).view(()), (ml, normalize)

    def prog_bwd(self, dout, ctx, grads, needs):
        ml, normalize = ctx
        return ops.dkl_bwd(dout.reshape(1), ml, normalize)


_DKL = _Dkl()


class GaussianDkl:
    """model/loss.py:46-66.  The reference passes mu and logvar as two [B,D] slices of the packed [B,2,D] encoder
    output; pass `packed=` to avoid re-packing, otherwise the slices are stacked."""

    def __init__(self, normalize=True):
        self.normalize = normalize

    def __call__(self, mu1, logvar1, mu2=None, logvar2=None, packed=None):
        if mu2 is not None or logvar2 is not None:
            raise NotImplementedError("General Dkl not implemented yet...")
        if packed is None:
            packed = torch.stack([mu1, logvar1], dim=1)
        return run_program(_DKL, (packed,), [], True, self.normalize)


class _FlowLatent:
    def prog_fwd(self, inputs, training, normalize):
        ml, z0, zk, ld = (t.contiguous() for t in inputs)
        return ops.latent_loss_fwd(ml, z0, zk, ld, normalize).view(()), (ml, z0, zk, normalize)

    def prog_bwd(self, dout, ctx, grads, needs):
        ml, z0, zk, normalize = ctx
        return ops.latent_loss_bwd(dout.reshape(1), ml, z0, zk, normalize)


_FLOW_LATENT = _FlowLatent()


def flow_latent_loss(z_0_mu_logvar, z_0_sampled, z_K_sampled, log_abs_det_jac, normalize):
    """FlowVAE.latent_loss (VAE.py:183-193)."""
    return run_program(_FLOW_LATENT, (z_0_mu_logvar, z_0_sampled, z_K_sampled, log_abs_det_jac), [], True, normalize)


class _Synth:
    def prog_fwd(self, inputs, training, cfg):
        v_out, v_in = inputs[0].contiguous(), inputs[1].contiguous()
        tables, normalize, factor, cat_softmax, temp = cfg
        out, ws = ops.synth_loss_fwd(v_out, v_in, tables, normalize, factor, cat_softmax, temp)
        return out.view(()), (v_out, v_in, ws, cfg)

    def prog_bwd(self, dout, ctx, grads, needs):
        v_out, v_in, ws, (tables, normalize, factor, cat_softmax, temp) = ctx
        return ops.synth_loss_bwd(dout.reshape(1), v_out, v_in, tables, normalize, factor, cat_softmax, temp, ws), None


_SYNTH = _Synth()


class SynthParamsLoss:
    """model/loss.py:73-183.  The useless-parameter search (silent Dexed operators, data/preset.py:247-283) runs on the
    device from `u_in` instead of Python loops with `.item()` per row.  Unlike the reference (loss.py:134-135) the
    inputs are NOT modified in place."""

    def __init__(self, idx_helper, normalize_losses: bool, categorical_loss_factor=0.2, prevent_useless_params_loss=True,
                 cat_bce=True, cat_softmax=False, cat_softmax_t=0.1):
        self.idx_helper = idx_helper
        self.normalize_losses = normalize_losses
        if cat_bce and cat_softmax:
            raise ValueError("'cat_bce' (Binary Cross-Entropy) and 'cat_softmax' (implies Categorical Cross-Entropy)"
                             "cannot be both set to True")
        if cat_bce:
            raise NotImplementedError("the binary cross-entropy variant (train.params_cat_bceloss=True) is not implemented; "
                                      "the reference calls it 'very bad perfs' (loss.py:94-95) and defaults to CCE")
        if not prevent_useless_params_loss:
            raise NotImplementedError("prevent_useless_params_loss=False is never used by the reference")
        self.cat_bce = cat_bce
        self.cat_softmax = cat_softmax
        self.cat_softmax_t = cat_softmax_t
        self.cat_loss_factor = categorical_loss_factor
        self.prevent_useless_params_loss = prevent_useless_params_loss
        self.num_indexes = self.idx_helper.get_numerical_learnable_indexes()
        self.cat_indexes = self.idx_helper.get_categorical_learnable_indexes()
        self._tables = _tables_for(idx_helper)

    def __call__(self, u_out: torch.Tensor, u_in: torch.Tensor):
        cfg = (self._tables, self.normalize_losses, self.cat_loss_factor, self.cat_softmax, self.cat_softmax_t)
        return run_program(_SYNTH, (u_out, u_in), [], True, cfg)


def _tables_for(idx_helper):
    """Device tables from this package's PresetIndexesHelper, or derived by probing a reference helper."""
    if hasattr(idx_helper, 'device_tables'):
        return ops.DeviceTables(idx_helper)
    from ..data.preset import tables_from_foreign_helper
    return ops.DeviceTables(tables_from_foreign_helper(idx_helper))


class _Total:
    """recons + beta * latent + controls (train.py:228) with beta read from DEVICE memory, so that the beta warm-up
    schedule (train.py:122-124) reaches a captured CUDA graph."""

    def prog_fwd(self, inputs, training, extra):
        recons, lat, cont, beta = (t.reshape(1) for t in inputs)
        return ops.add(ops.add(recons, ops.mul(lat, beta)), cont).view(()), beta

    def prog_bwd(self, dout, ctx, grads, needs):
        g = dout.reshape(1)
        return g.view(()), ops.mul(g, ctx).view(()), g.view(()), None


_TOTAL = _Total()


def total_loss(recons, latent, controls, beta_dev):
    """beta_dev: 1-element float32 device tensor."""
    return run_program(_TOTAL, (recons, latent, controls, beta_dev), [], True)
