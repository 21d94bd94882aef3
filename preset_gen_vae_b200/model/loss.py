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
        tables, normalize, factor, cat_softmax, temp, counts = cfg
        out, ws = ops.synth_loss_fwd(v_out, v_in, tables, normalize, factor, cat_softmax, temp, counts)
        return out.view(()), (v_out, v_in, ws, cfg)

    def prog_bwd(self, dout, ctx, grads, needs):
        v_out, v_in, ws, (tables, normalize, factor, cat_softmax, temp, _) = ctx
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

    def __call__(self, u_out: torch.Tensor, u_in: torch.Tensor, group_counts=None):
        """group_counts (extension, fp64 [n_groups] device tensor): useful-row counts to normalise the categorical groups with instead
        of this batch's own - see TrainStep (data-parallel shards) and pgv.h."""
        cfg = (self._tables, self.normalize_losses, self.cat_loss_factor, self.cat_softmax, self.cat_softmax_t, group_counts)
        return run_program(_SYNTH, (u_out, u_in), [], True, cfg)

    def useful_counts(self, u_in):
        return ops.synth_useful_counts(u_in, self._tables)


class QuantizedNumericalParamsLoss:
    """model/loss.py:187-261: numerical VST parameters only, inferred values quantised like the synthesizer does (round to the
    parameter's cardinality; one-hot representations -> argmax / (cardinal - 1)); not differentiable.  One kernel pass shared with
    CategoricalParamsAccuracy (`PresetMetrics`); the result stays on the device (0-dim tensor), train.py:232 only logs it."""

    def __init__(self, idx_helper, numerical_loss=None, limited_vst_params_indexes=None):
        self.idx_helper = idx_helper
        name = type(numerical_loss).__name__ if numerical_loss is not None else 'MSELoss'
        if name not in ('MSELoss', 'L1Loss') or getattr(numerical_loss, 'reduction', 'mean') != 'mean':
            raise NotImplementedError("numerical_loss must be nn.MSELoss() or nn.L1Loss() with mean reduction (train.py:121-122, eval.py)")
        self.l1 = name == 'L1Loss'
        self.numerical_loss = numerical_loss
        self.limited_vst_params_indexes = limited_vst_params_indexes
        self.num_params_count = len(idx_helper.num_idx_learned_as_num) + len(idx_helper.num_idx_learned_as_cat)
        self._tables = ops.MetricTables(idx_helper, limited_vst_params_indexes)

    def __call__(self, u_out, u_in):
        out4, _ = ops.preset_metrics(u_out, u_in, self._tables, l1=self.l1)
        if self.limited_vst_params_indexes is None:
            return out4[0]
        # limited lists: the reference still averages over ALL pre-allocated numerical columns, the unused ones zero-filled (loss.py:222-226)
        return out4[0] * out4[2] / float(self.num_params_count)


class CategoricalParamsAccuracy:
    """model/loss.py:265-315: accuracy of the categorical VST parameters (one-hot: argmax match; numerical representation: match after
    rounding to the class index).  reduce=True returns a 0-dim DEVICE tensor (the reference returns a numpy scalar after one `.item()`
    per parameter); reduce=False returns {vst_idx: accuracy} like the reference (one device -> host copy)."""

    def __init__(self, idx_helper, reduce=True, percentage_output=True, limited_vst_params_indexes=None):
        self.idx_helper = idx_helper
        self.reduce = reduce
        self.percentage_output = percentage_output
        self.limited_vst_params_indexes = limited_vst_params_indexes
        self._tables = ops.MetricTables(idx_helper, limited_vst_params_indexes)

    def __call__(self, u_out, u_in):
        out4, acc = ops.preset_metrics(u_out, u_in, self._tables, acc_scale=100.0 if self.percentage_output else 1.0,
                                       per_param=not self.reduce)
        if self.reduce:
            return out4[1]
        acc = acc.cpu().tolist()
        # the reference fills its dict with the numerically-learned parameters first, then the one-hot ones (loss.py:289-307)
        keys = list(self.idx_helper.cat_idx_learned_as_num.keys()) + list(self.idx_helper.cat_idx_learned_as_cat.keys())
        return {k: acc[k] for k in keys if acc[k] >= 0.0}


class PresetMetrics:
    """Both monitoring metrics of train.py:232-233 from ONE kernel pass: returns a device tensor
    [QuantizedNumericalParamsLoss (MSE), CategoricalParamsAccuracy (%), #numerical, #categorical]."""

    def __init__(self, idx_helper):
        self._tables = ops.MetricTables(idx_helper)

    def __call__(self, u_out, u_in):
        return ops.preset_metrics(u_out, u_in, self._tables)[0]


class _FlowParams:
    def prog_fwd(self, inputs, training, divisor):
        ml, z0, ld_t, ld_u = (t.contiguous() for t in inputs)
        return ops.flow_params_loss_fwd(ml, z0, ld_t, ld_u, divisor).view(()), (ml, z0, divisor)

    def prog_bwd(self, dout, ctx, grads, needs):
        ml, z0, divisor = ctx
        dml, dz0, dld = ops.flow_params_loss_bwd(dout.reshape(1), ml, z0, divisor)
        return dml, dz0, dld, dld


_FLOW_PARAMS = _FlowParams()


class FlowParamsLoss:
    """model/loss.py:318-346: -mean(log q_Z0(z0) + log|det J_invT| + log|det J_invU|) / 1000 with z0 = invT(invU(v_target)); the two
    inverse-flow functions are the models' `flow_inverse_function`s (train.py:117-119) and are differentiated through."""

    def __init__(self, idx_helper, latent_flow_inverse_function, reg_flow_inverse_function):
        self.idx_helper = idx_helper
        self.latent_flow_inverse_function = latent_flow_inverse_function
        self.reg_flow_inverse_function = reg_flow_inverse_function

    def __call__(self, z_0_mu_logvar, v_target):
        z_K, ld_u = self.reg_flow_inverse_function(v_target)
        z_0, ld_t = self.latent_flow_inverse_function(z_K)
        return run_program(_FLOW_PARAMS, (z_0_mu_logvar, z_0, ld_t, ld_u), [], True, 1000.0)


def _tables_for(idx_helper):
    """Device tables from this package's PresetIndexesHelper, or derived by probing a reference helper."""
    if hasattr(idx_helper, 'device_tables'):
        return ops.DeviceTables(idx_helper)
    from ..data.preset import tables_from_foreign_helper
    return ops.DeviceTables(tables_from_foreign_helper(idx_helper))


class _Total:
    """recons + beta * latent + controls [+ 0.1 * beta * flow_input_dkl] (train.py:227, 236-246) with beta read from DEVICE memory, so
    that the beta warm-up schedule (train.py:122-124) reaches a captured CUDA graph."""

    def prog_fwd(self, inputs, training, factor):
        recons, lat, cont, beta = (t.reshape(1) for t in inputs[:4])
        total = ops.add(ops.add(recons, ops.mul(lat, beta)), cont)
        f = None
        if len(inputs) > 4:
            f = torch.full((1,), factor, dtype=torch.float32, device=beta.device)
            total = ops.add(total, ops.mul(ops.mul(inputs[4].reshape(1), f), beta))
        return total.view(()), (beta, f)

    def prog_bwd(self, dout, ctx, grads, needs):
        beta, f = ctx
        g = dout.reshape(1)
        gb = ops.mul(g, beta)
        out = (g.view(()), gb.view(()), g.view(()), None)
        if f is not None:
            out = out + (ops.mul(gb, f).view(()),)
        return out


_TOTAL = _Total()


def total_loss(recons, latent, controls, beta_dev, flow_input=None, flow_input_factor=0.1):
    """beta_dev: 1-element float32 device tensor.  flow_input: un-scaled GaussianDkl of the flow input (train.py:236-239), added as
    flow_input_factor * beta * flow_input."""
    if flow_input is None:
        return run_program(_TOTAL, (recons, latent, controls, beta_dev), [], True, 1.0)
    return run_program(_TOTAL, (recons, latent, controls, beta_dev, flow_input), [], True, float(flow_input_factor))
