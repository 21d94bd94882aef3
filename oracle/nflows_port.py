"""Restatement of the nflows ~=0.14 classes the reference hot path instantiates.  TEST INFRASTRUCTURE.

PARITY UNPINNED: nflows is an un-vendored dependency (reference requirements.txt:11), absent from
/root/reference and from this image.  These classes restate its published algorithm (RealNVP affine
coupling with residual-MLP conditioners, Dinh et al. 2017, as implemented by nflows 0.14) and keep
nflows' attribute names so that state_dict keys match what a reference checkpoint holds.

Reference call sites anchoring the behaviour:
  model/VAE.py:118-125       SimpleRealNVP(features, hidden_features, num_layers, num_blocks_per_layer=2,
                             batch_norm_within_layers=True, batch_norm_between_layers=False)._transform
  model/VAE.py:178           z_K, log_abs_det_jac = self.flow_transform(z_0)
  model/flows.py:42-90       CustomRealNVP(CompositeTransform) built from AffineCouplingTransform,
                             nets.ResidualNet and transforms.normalization.BatchNorm
  model/regression.py:186-189 v_out, _ = flow.forward(z_K)  /  .inverse for the other direction

`install_as_nflows()` registers this module tree under the name `nflows` in sys.modules so that the
reference's own model/*.py files can be imported unmodified by oracle/make_golden.py.
"""
import sys
import types

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F
from torch.nn import init


class Transform(nn.Module):
    def forward(self, inputs, context=None):
        raise NotImplementedError()

    def inverse(self, inputs, context=None):
        raise NotImplementedError()


class CompositeTransform(Transform):
    """Cascade of transforms; log|det J| values add up.  Inverse walks the list backwards."""

    def __init__(self, transforms):
        super().__init__()
        self._transforms = nn.ModuleList(transforms)

    @staticmethod
    def _cascade(inputs, funcs, context):
        batch_size = inputs.shape[0]
        outputs = inputs
        total_logabsdet = inputs.new_zeros(batch_size)
        for func in funcs:
            outputs, logabsdet = func(outputs, context)
            total_logabsdet = total_logabsdet + logabsdet
        return outputs, total_logabsdet

    def forward(self, inputs, context=None):
        return self._cascade(inputs, self._transforms, context)

    def inverse(self, inputs, context=None):
        funcs = (transform.inverse for transform in self._transforms[::-1])
        return self._cascade(inputs, funcs, context)


class ResidualBlock(nn.Module):
    """[BN1d(eps=1e-3)] -> act -> Linear -> [BN1d] -> act -> Dropout -> Linear(init U(-1e-3, 1e-3)); x + f(x)."""

    def __init__(self, features, context_features, activation=F.relu, dropout_probability=0.0,
                 use_batch_norm=False, zero_initialization=True):
        super().__init__()
        self.activation = activation
        self.use_batch_norm = use_batch_norm
        if use_batch_norm:
            self.batch_norm_layers = nn.ModuleList([nn.BatchNorm1d(features, eps=1e-3) for _ in range(2)])
        if context_features is not None:
            self.context_layer = nn.Linear(context_features, features)
        self.linear_layers = nn.ModuleList([nn.Linear(features, features) for _ in range(2)])
        self.dropout = nn.Dropout(p=dropout_probability)
        if zero_initialization:
            init.uniform_(self.linear_layers[-1].weight, -1e-3, 1e-3)
            init.uniform_(self.linear_layers[-1].bias, -1e-3, 1e-3)

    def forward(self, inputs, context=None):
        temps = inputs
        if self.use_batch_norm:
            temps = self.batch_norm_layers[0](temps)
        temps = self.activation(temps)
        temps = self.linear_layers[0](temps)
        if self.use_batch_norm:
            temps = self.batch_norm_layers[1](temps)
        temps = self.activation(temps)
        temps = self.dropout(temps)
        temps = self.linear_layers[1](temps)
        if context is not None:
            temps = F.glu(torch.cat((temps, self.context_layer(context)), dim=1), dim=1)
        return inputs + temps


class ResidualNet(nn.Module):
    """initial_layer Linear -> num_blocks x ResidualBlock -> final_layer Linear."""

    def __init__(self, in_features, out_features, hidden_features, context_features=None, num_blocks=2,
                 activation=F.relu, dropout_probability=0.0, use_batch_norm=False):
        super().__init__()
        self.hidden_features = hidden_features
        self.context_features = context_features
        if context_features is not None:
            self.initial_layer = nn.Linear(in_features + context_features, hidden_features)
        else:
            self.initial_layer = nn.Linear(in_features, hidden_features)
        self.blocks = nn.ModuleList([
            ResidualBlock(features=hidden_features, context_features=context_features, activation=activation,
                          dropout_probability=dropout_probability, use_batch_norm=use_batch_norm)
            for _ in range(num_blocks)])
        self.final_layer = nn.Linear(hidden_features, out_features)

    def forward(self, inputs, context=None):
        if context is None:
            temps = self.initial_layer(inputs)
        else:
            temps = self.initial_layer(torch.cat((inputs, context), dim=1))
        for block in self.blocks:
            temps = block(temps, context=context)
        return self.final_layer(temps)


class CouplingTransform(Transform):
    """Splits features by the sign of `mask` (<=0: identity / conditioner input, >0: transformed)."""

    def __init__(self, mask, transform_net_create_fn, unconditional_transform=None):
        mask = torch.as_tensor(mask)
        if mask.dim() != 1:
            raise ValueError("Mask must be a 1-dim tensor.")
        if mask.numel() <= 0:
            raise ValueError("Mask can't be empty.")
        super().__init__()
        self.features = len(mask)
        features_vector = torch.arange(self.features)
        self.register_buffer("identity_features", features_vector.masked_select(mask <= 0))
        self.register_buffer("transform_features", features_vector.masked_select(mask > 0))
        assert self.num_identity_features + self.num_transform_features == self.features
        self.transform_net = transform_net_create_fn(
            self.num_identity_features, self.num_transform_features * self._transform_dim_multiplier())
        assert unconditional_transform is None  # never used by the reference
        self.unconditional_transform = None

    @property
    def num_identity_features(self):
        return len(self.identity_features)

    @property
    def num_transform_features(self):
        return len(self.transform_features)

    def forward(self, inputs, context=None):
        identity_split = inputs[:, self.identity_features, ...]
        transform_split = inputs[:, self.transform_features, ...]
        transform_params = self.transform_net(identity_split, context)
        transform_split, logabsdet = self._coupling_transform_forward(transform_split, transform_params)
        outputs = torch.empty_like(inputs)
        outputs[:, self.identity_features, ...] = identity_split
        outputs[:, self.transform_features, ...] = transform_split
        return outputs, logabsdet

    def inverse(self, inputs, context=None):
        identity_split = inputs[:, self.identity_features, ...]
        transform_split = inputs[:, self.transform_features, ...]
        transform_params = self.transform_net(identity_split, context)
        transform_split, logabsdet = self._coupling_transform_inverse(transform_split, transform_params)
        outputs = torch.empty_like(inputs)
        outputs[:, self.identity_features] = identity_split
        outputs[:, self.transform_features] = transform_split
        return outputs, logabsdet


class AffineCouplingTransform(CouplingTransform):
    """y = x * s + t with s = sigmoid(u + 2) + 1e-3; params[:, :n] = shift t, params[:, n:] = u."""

    def _transform_dim_multiplier(self):
        return 2

    def _scale_and_shift(self, transform_params):
        unconstrained_scale = transform_params[:, self.num_transform_features:, ...]
        shift = transform_params[:, :self.num_transform_features, ...]
        scale = torch.sigmoid(unconstrained_scale + 2) + 1e-3
        return scale, shift

    def _coupling_transform_forward(self, inputs, transform_params):
        scale, shift = self._scale_and_shift(transform_params)
        log_scale = torch.log(scale)
        outputs = inputs * scale + shift
        return outputs, log_scale.reshape(log_scale.shape[0], -1).sum(dim=1)

    def _coupling_transform_inverse(self, inputs, transform_params):
        scale, shift = self._scale_and_shift(transform_params)
        log_scale = torch.log(scale)
        outputs = (inputs - shift) / scale
        return outputs, -log_scale.reshape(log_scale.shape[0], -1).sum(dim=1)


class AdditiveCouplingTransform(AffineCouplingTransform):
    """Volume-preserving variant (never selected by the reference's default config)."""

    def _transform_dim_multiplier(self):
        return 1

    def _scale_and_shift(self, transform_params):
        return torch.ones_like(transform_params), transform_params


class BatchNorm(Transform):
    """Invertible batch-norm transform (flows.py:87-88): batch mean / UNBIASED variance in training,
    weight = softplus(unconstrained_weight) + eps, running_var initialised to ZERO (nflows quirk)."""

    def __init__(self, features, eps=1e-5, momentum=0.1, affine=True):
        super().__init__()
        self.momentum = momentum
        self.eps = eps
        constant = np.log(np.exp(1 - eps) - 1)
        self.unconstrained_weight = nn.Parameter(constant * torch.ones(features))
        self.bias = nn.Parameter(torch.zeros(features))
        self.register_buffer("running_mean", torch.zeros(features))
        self.register_buffer("running_var", torch.zeros(features))

    @property
    def weight(self):
        return F.softplus(self.unconstrained_weight) + self.eps

    def forward(self, inputs, context=None):
        if inputs.dim() != 2:
            raise ValueError("Expected 2-dim inputs, got inputs of shape: {}".format(inputs.shape))
        if self.training:
            mean, var = inputs.mean(0), inputs.var(0)
            self.running_mean.mul_(1 - self.momentum).add_(mean.detach() * self.momentum)
            self.running_var.mul_(1 - self.momentum).add_(var.detach() * self.momentum)
        else:
            mean, var = self.running_mean, self.running_var
        outputs = self.weight * ((inputs - mean) / torch.sqrt((var + self.eps))) + self.bias
        logabsdet_ = torch.log(self.weight) - 0.5 * torch.log(var + self.eps)
        logabsdet = torch.sum(logabsdet_) * inputs.new_ones(inputs.shape[0])
        return outputs, logabsdet

    def inverse(self, inputs, context=None):
        if self.training:
            raise RuntimeError("Batch norm inverse is only available in eval mode, not in training mode.")
        if inputs.dim() != 2:
            raise ValueError("Expected 2-dim inputs, got inputs of shape: {}".format(inputs.shape))
        outputs = torch.sqrt(self.running_var + self.eps) * ((inputs - self.bias) / self.weight) + self.running_mean
        logabsdet_ = -torch.log(self.weight) + 0.5 * torch.log(self.running_var + self.eps)
        logabsdet = torch.sum(logabsdet_) * inputs.new_ones(inputs.shape[0])
        return outputs, logabsdet


class StandardNormal(nn.Module):
    def __init__(self, shape):
        super().__init__()
        self._shape = torch.Size(shape)
        self.register_buffer("_log_z", torch.tensor(0.5 * np.prod(shape) * np.log(2 * np.pi), dtype=torch.float64),
                             persistent=False)


class Flow(nn.Module):
    def __init__(self, transform, distribution, embedding_net=None):
        super().__init__()
        self._transform = transform
        self._distribution = distribution


class SimpleRealNVP(Flow):
    """Alternating-mask affine couplings; mask = ones, mask[::2] = -1, sign flipped after every layer."""

    def __init__(self, features, hidden_features, num_layers, num_blocks_per_layer, use_volume_preserving=False,
                 activation=F.relu, dropout_probability=0.0, batch_norm_within_layers=False,
                 batch_norm_between_layers=False):
        coupling_constructor = AdditiveCouplingTransform if use_volume_preserving else AffineCouplingTransform
        mask = torch.ones(features)
        mask[::2] = -1

        def create_resnet(in_features, out_features):
            return ResidualNet(in_features, out_features, hidden_features=hidden_features,
                               num_blocks=num_blocks_per_layer, activation=activation,
                               dropout_probability=dropout_probability, use_batch_norm=batch_norm_within_layers)

        layers = []
        for _ in range(num_layers):
            layers.append(coupling_constructor(mask=mask, transform_net_create_fn=create_resnet))
            mask *= -1
            if batch_norm_between_layers:
                layers.append(BatchNorm(features=features))
        super().__init__(transform=CompositeTransform(layers), distribution=StandardNormal([features]))


class _Unsupported(nn.Module):
    """Placeholder for nflows classes the reference imports but the in-scope path never builds (MAF branch)."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError("out of scope: only the RealNVP branch of the reference is restated")


def install_as_nflows():
    """Expose this module under the `nflows.*` names imported by the reference's model/*.py."""
    if "nflows" in sys.modules and not getattr(sys.modules["nflows"], "_is_oracle_port", False):
        return  # a real nflows is installed: use it (it would then PIN these classes)
    me = sys.modules[__name__]

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m._is_oracle_port = True
        sys.modules[name] = m
        return m

    nets = mod("nflows.nn.nets", ResidualNet=ResidualNet, ResidualBlock=ResidualBlock)
    nn_ = mod("nflows.nn", nets=nets)
    base = mod("nflows.transforms.base", CompositeTransform=CompositeTransform, Transform=Transform)
    coupling = mod("nflows.transforms.coupling", AffineCouplingTransform=AffineCouplingTransform,
                   AdditiveCouplingTransform=AdditiveCouplingTransform)
    autoreg = mod("nflows.transforms.autoregressive", MaskedAffineAutoregressiveTransform=_Unsupported)
    perm = mod("nflows.transforms.permutations", ReversePermutation=_Unsupported)
    norm = mod("nflows.transforms.normalization", BatchNorm=BatchNorm)
    transforms = mod("nflows.transforms", base=base, coupling=coupling, autoregressive=autoreg,
                     permutations=perm, normalization=norm)
    realnvp = mod("nflows.flows.realnvp", SimpleRealNVP=SimpleRealNVP)
    flows_base = mod("nflows.flows.base", Flow=Flow)
    flows = mod("nflows.flows", realnvp=realnvp, base=flows_base)
    dnormal = mod("nflows.distributions.normal", StandardNormal=StandardNormal)
    dists = mod("nflows.distributions", normal=dnormal)
    mod("nflows", nn=nn_, transforms=transforms, flows=flows, distributions=dists, _port=me)
