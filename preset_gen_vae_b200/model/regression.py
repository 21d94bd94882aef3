"""Synth-parameter regression heads with the reference's interface (model/regression.py:20-189):
`PresetActivation`, `MLPRegression`, `FlowRegression` (a `CustomRealNVP` followed by Hardtanh(0,1)).
"""
import torch
import torch.nn as nn

from . import ops
from .flows import CustomRealNVP
from .program import run_program


class _Hardtanh:
    def prog_fwd(self, inputs, training, rng):
        x = inputs[0].contiguous()
        return ops.hardtanh_fwd(x, rng[0], rng[1]), (x, rng)

    def prog_bwd(self, dout, ctx, grads, needs):
        x, rng = ctx
        return ops.hardtanh_bwd(dout, x, rng[0], rng[1])


class _SoftmaxAct:
    def prog_fwd(self, inputs, training, tables):
        x = inputs[0].contiguous()
        y = ops.preset_act_softmax_fwd(x, tables)
        return y, (x, y, tables)

    def prog_bwd(self, dout, ctx, grads, needs):
        x, y, tables = ctx
        return ops.preset_act_softmax_bwd(dout, x, y, tables)


_HARDTANH, _SOFTMAX_ACT = _Hardtanh(), _SoftmaxAct()


class PresetActivation(nn.Module):
    """Hardtanh(0,1) on every output, or (cat_softmax_activation=True) Hardtanh on numerical outputs and a softmax per
    categorical group (regression.py:20-53)."""

    def __init__(self, idx_helper, numerical_activation=nn.Hardtanh(min_val=0.0, max_val=1.0), cat_softmax_activation=False):
        super().__init__()
        self.idx_helper = idx_helper
        if not isinstance(numerical_activation, nn.Hardtanh):
            raise NotImplementedError("only nn.Hardtanh numerical activations are implemented")
        self.numerical_act = numerical_activation
        self.cat_softmax_activation = cat_softmax_activation
        self._tables = None
        if self.cat_softmax_activation:
            self.categorical_act = nn.Softmax(dim=-1)
            self.num_indexes = self.idx_helper.get_numerical_learnable_indexes()
            self.cat_indexes = self.idx_helper.get_categorical_learnable_indexes()
            from .loss import _tables_for
            self._tables = _tables_for(idx_helper)

    def forward(self, x):
        if self.cat_softmax_activation:
            assert (self.numerical_act.min_val, self.numerical_act.max_val) == (0.0, 1.0)
            return run_program(_SOFTMAX_ACT, (x,), [], True, self._tables)
        return run_program(_HARDTANH, (x,), [], True, (self.numerical_act.min_val, self.numerical_act.max_val))


class _MLP:
    """Program of MLPRegression: [Linear -> BN1d -> Dropout -> ReLU] x (n-1) -> Linear -> ReLU -> Linear."""

    def __init__(self, owner):
        self.o = owner

    def prog_fwd(self, inputs, training, masks):
        o = self.o
        h = inputs[0].contiguous()
        ctxs = []
        for l, (fc, bn) in enumerate(o._layers):
            if bn is not None:
                u = ops.linear_fwd(h, fc.weight, fc.bias)
                if training:
                    # reference order is BN -> Dropout -> ReLU; dropout (scaling by >= 0) and ReLU commute
                    t, m, r = ops.bn1d_train_fwd(u, bn, relu=True, mask=None if masks is None else masks[l])
                    o._nbt_pending[l] += 1
                    ctxs.append((h, u, m, r, None if masks is None else masks[l]))
                else:
                    t = ops.bn1d_eval_fwd(u, bn, relu=True)
                h = t
            else:
                ctxs.append((h,))
                h = ops.linear_fwd(h, fc.weight, fc.bias, relu=True)
                ctxs[-1] = ctxs[-1] + (h,)
        fc = o._last
        ctxs.append((h,))
        return ops.linear_fwd(h, fc.weight, fc.bias), ctxs

    def prog_bwd(self, dout, ctxs, grads, needs):
        o = self.o
        fc = o._last
        (h,) = ctxs[-1]
        grads[id(fc.weight)], grads[id(fc.bias)] = ops.linear_wgrad(dout, h)
        d = ops.linear_dgrad(dout, fc.weight)
        for l in range(len(o._layers) - 1, -1, -1):
            fc, bn = o._layers[l]
            if bn is not None:
                h_in, u, m, r, mask = ctxs[l]
                du, grads[id(bn.weight)], grads[id(bn.bias)] = ops.bn1d_train_bwd(d, u, bn, m, r, relu=True, mask=mask)
            else:
                h_in, h_out = ctxs[l]
                du = ops.lrelu_bwd(d, h_out, 0.0)
            grads[id(fc.weight)], grads[id(fc.bias)] = ops.linear_wgrad(du, h_in)
            d = ops.linear_dgrad(du, fc.weight)
        return d


class MLPRegression(nn.Module):
    def __init__(self, architecture, dim_z, idx_helper, dropout_p=0.0, cat_softmax_activation=False):
        super().__init__()
        self.architecture = architecture.split('_')
        self.dim_z = dim_z
        self.idx_helper = idx_helper
        if len(self.architecture) == 1:
            num_hidden_layers, num_hidden_neurons = (int(v) for v in self.architecture[0].split('l'))
        else:
            raise NotImplementedError("Arch suffix arguments not implemented yet")
        self.reg_model = nn.Sequential()
        self._layers = []
        self.dropout_p = dropout_p
        for l in range(0, num_hidden_layers):
            fc = nn.Linear(dim_z if l == 0 else num_hidden_neurons, num_hidden_neurons)
            self.reg_model.add_module('fc{}'.format(l + 1), fc)
            bn = None
            if l < (num_hidden_layers - 1):
                bn = nn.BatchNorm1d(num_features=num_hidden_neurons)
                self.reg_model.add_module('bn{}'.format(l + 1), bn)
                self.reg_model.add_module('drp{}'.format(l + 1), nn.Dropout(dropout_p))
            self.reg_model.add_module('act{}'.format(l + 1), nn.ReLU())
            self._layers.append((fc, bn))
        self._last_name = 'fc{}'.format(num_hidden_layers + 1)
        self.reg_model.add_module(self._last_name, nn.Linear(num_hidden_neurons, self.idx_helper.learnable_preset_size))
        self.reg_model.add_module('act', PresetActivation(self.idx_helper, cat_softmax_activation=cat_softmax_activation))
        self._nbt_pending = [0] * num_hidden_layers
        self._prog = _MLP(self)

    @property
    def _last(self):            # not registered a second time: the state_dict must have the reference's keys only
        return getattr(self.reg_model, self._last_name)

    def state_dict(self, *args, **kwargs):
        for l, (fc, bn) in enumerate(self._layers):
            if bn is not None and self._nbt_pending[l]:
                bn.num_batches_tracked += self._nbt_pending[l]
            self._nbt_pending[l] = 0
        return super().state_dict(*args, **kwargs)

    def forward(self, z_K, dropout_masks=None):
        from .encoder import make_dropout_mask
        if self.training and dropout_masks is None and self.dropout_p > 0.0:
            dropout_masks = [make_dropout_mask((z_K.shape[0], fc.out_features), self.dropout_p, z_K.device) if bn is not None
                             else None for fc, bn in self._layers]
        params = [p for fc, bn in self._layers for p in ([fc.weight, fc.bias] + ([bn.weight, bn.bias] if bn is not None else []))]
        params += [self._last.weight, self._last.bias]
        v = run_program(self._prog, (z_K,), params, self.training, dropout_masks)
        return self.reg_model.act(v)


class FlowRegression(nn.Module):
    def __init__(self, architecture, dim_z, idx_helper, dropout_p=0.0, fast_forward_flow=True, cat_softmax_activation=False):
        super().__init__()
        self.dim_z = dim_z
        self.idx_helper = idx_helper
        self._fast_forward_flow = fast_forward_flow
        arch_args = architecture.split('_')
        if len(arch_args) < 2:
            raise AssertionError("Unvalid architecture string argument '{}' does not contain enough information"
                                 .format(architecture))
        elif len(arch_args) == 2:
            self.flow_type = arch_args[0]
            self.num_flow_layers, self.num_flow_hidden_features = (int(v) for v in arch_args[1].split('l'))
            self.bn_between_flows = True
            self.bn_within_flows = True
        else:
            raise NotImplementedError("Arch suffix arguments not implemented yet (too many arch args given in '{}')"
                                      .format(architecture))
        if self.flow_type.lower() in ('realnvp', 'rnvp'):
            self._forward_flow_transform = CustomRealNVP(self.dim_z, self.num_flow_hidden_features, self.num_flow_layers,
                                                         num_blocks_per_layer=2,
                                                         batch_norm_between_layers=self.bn_between_flows,
                                                         batch_norm_within_layers=self.bn_within_flows,
                                                         dropout_probability=dropout_p)
        else:
            raise NotImplementedError("'{}' regression flows are out of scope (only RealNVP; the reference calls its MAF "
                                      "branch 'very unstable', regression.py:160-163)".format(self.flow_type))
        self.activation_layer = PresetActivation(self.idx_helper, cat_softmax_activation=cat_softmax_activation)

    @property
    def is_flow_fast_forward(self):
        return self._fast_forward_flow

    @property
    def flow_forward_function(self):
        t = self._forward_flow_transform
        return t.forward if self._fast_forward_flow else t.inverse

    @property
    def flow_inverse_function(self):
        t = self._forward_flow_transform
        return t.inverse if self._fast_forward_flow else t.forward

    def forward(self, z_K, dropout_masks=None):
        if dropout_masks is not None:
            v_out, _ = self._forward_flow_transform.forward(z_K, dropout_masks=dropout_masks)
        else:
            v_out, _ = self.flow_forward_function(z_K)
        return self.activation_layer(v_out)
