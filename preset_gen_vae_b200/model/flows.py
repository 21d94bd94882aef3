"""RealNVP flows with the nflows 0.14 module tree (attribute names = state_dict keys of a reference checkpoint):

    CompositeTransform._transforms[i]            AffineCouplingTransform | BatchNorm
    AffineCouplingTransform.identity_features / transform_features (int64 buffers), .transform_net = ResidualNet
    ResidualNet.initial_layer, .blocks[j].{batch_norm_layers[0..1], linear_layers[0..1], dropout}, .final_layer
    BatchNorm.unconstrained_weight, .bias, .running_mean, .running_var

Call sites in the reference: model/VAE.py:118-125 (`SimpleRealNVP(...)._transform`), model/flows.py:42-90
(`CustomRealNVP`), model/regression.py:142-148.  nflows itself is not installed here; its behaviour is restated in
oracle/nflows_port.py (PARITY UNPINNED) and these classes are checked against that restatement.

A whole CompositeTransform runs as ONE autograd node (`CompositeTransform.forward`); its forward and backward launch
libpgv.so kernels only (dense layers, BatchNorm1d+ReLU(+dropout), affine coupling with log|det J|, flow BatchNorm).
"""
import numpy as np
import torch
from torch import nn
from torch.nn import functional as F
from torch.nn import init

from . import ops
from .program import run_program


class ResidualBlock(nn.Module):
    """[BN1d(eps=1e-3)] -> relu -> Linear -> [BN1d] -> relu -> dropout -> Linear(init U(-1e-3,1e-3)); out = x + f(x)."""

    def __init__(self, features, context_features, activation=F.relu, dropout_probability=0.0, use_batch_norm=False,
                 zero_initialization=True):
        super().__init__()
        assert context_features is None and activation is F.relu
        self.use_batch_norm = use_batch_norm
        if use_batch_norm:
            self.batch_norm_layers = nn.ModuleList([nn.BatchNorm1d(features, eps=1e-3) for _ in range(2)])
        else:
            raise NotImplementedError("the reference always builds its flows with batch_norm_within_layers=True")
        self.linear_layers = nn.ModuleList([nn.Linear(features, features) for _ in range(2)])
        self.dropout = nn.Dropout(p=dropout_probability)
        if zero_initialization:
            init.uniform_(self.linear_layers[-1].weight, -1e-3, 1e-3)
            init.uniform_(self.linear_layers[-1].bias, -1e-3, 1e-3)
        self._nbt_pending = 0

    def params(self):
        p = []
        for bn in self.batch_norm_layers:
            p += [bn.weight, bn.bias]
        for lin in self.linear_layers:
            p += [lin.weight, lin.bias]
        return p

    def fwd_fused(self, x, t0, m0, r0, mask, next_bn):
        """Training forward with the BatchNorms fused into the Linear layers (column-slice kernels).  (t0, m0, r0) =
        relu(bn0(x)) and its statistics come from the kernel that produced x; if `next_bn` is given, the second Linear also
        applies the following block's first BatchNorm and hands (t0', m0', r0') on."""
        bn0, bn1 = self.batch_norm_layers
        l0, l1 = self.linear_layers
        u, t1, m1, r1 = ops.linear_bn_fwd(t0, l0.weight, l0.bias, bn1, mask=mask)
        self._nbt_pending += 1
        ctx = (x, m0, r0, t0, u, m1, r1, t1, mask)
        if next_bn is None:
            return ops.linear_fwd(t1, l1.weight, l1.bias, residual=x), None, ctx
        y, nt0, nm0, nr0 = ops.linear_bn_fwd(t1, l1.weight, l1.bias, next_bn, residual=x)
        return y, (nt0, nm0, nr0), ctx

    def bwd_fused(self, dy, ctx, grads):
        x, m0, r0, t0, u, m1, r1, t1, mask = ctx
        bn0, bn1 = self.batch_norm_layers
        l0, l1 = self.linear_layers
        with ops.forked(dy, dy, t1):
            grads[id(l1.weight)], grads[id(l1.bias)] = ops.linear_wgrad(dy, t1)
        du, grads[id(bn1.weight)], grads[id(bn1.bias)] = ops.linear_dgrad_bn_bwd(dy, l1.weight, u, bn1, m1, r1, mask=mask)
        with ops.forked(du, du, t0):
            grads[id(l0.weight)], grads[id(l0.bias)] = ops.linear_wgrad(du, t0)
        dx, grads[id(bn0.weight)], grads[id(bn0.bias)] = ops.linear_dgrad_bn_bwd(du, l0.weight, x, bn0, m0, r0, add_post=dy)   # + residual path
        return dx

    def fwd(self, x, training, mask):
        bn0, bn1 = self.batch_norm_layers
        l0, l1 = self.linear_layers
        if training:
            t0, m0, r0 = ops.bn1d_train_fwd(x, bn0, relu=True)
            u = ops.linear_fwd(t0, l0.weight, l0.bias)
            t1, m1, r1 = ops.bn1d_train_fwd(u, bn1, relu=True, mask=mask)
            self._nbt_pending += 1
            y = ops.linear_fwd(t1, l1.weight, l1.bias, residual=x)
            return y, (x, m0, r0, t0, u, m1, r1, t1, mask)
        t0 = ops.bn1d_eval_fwd(x, bn0, relu=True)
        u = ops.linear_fwd(t0, l0.weight, l0.bias)
        t1 = ops.bn1d_eval_fwd(u, bn1, relu=True)
        return ops.linear_fwd(t1, l1.weight, l1.bias, residual=x), None

    def bwd(self, dy, ctx, grads):
        x, m0, r0, t0, u, m1, r1, t1, mask = ctx
        bn0, bn1 = self.batch_norm_layers
        l0, l1 = self.linear_layers
        with ops.forked(dy, dy, t1):
            grads[id(l1.weight)], grads[id(l1.bias)] = ops.linear_wgrad(dy, t1)
        dt1 = ops.linear_dgrad(dy, l1.weight)
        du, grads[id(bn1.weight)], grads[id(bn1.bias)] = ops.bn1d_train_bwd(dt1, u, bn1, m1, r1, relu=True, mask=mask)
        with ops.forked(du, du, t0):
            grads[id(l0.weight)], grads[id(l0.bias)] = ops.linear_wgrad(du, t0)
        dt0 = ops.linear_dgrad(du, l0.weight)
        dx, grads[id(bn0.weight)], grads[id(bn0.bias)] = ops.bn1d_train_bwd(dt0, x, bn0, m0, r0, relu=True)
        return ops.add(dx, dy)                           # residual path

    def flush_counters(self):
        if self._nbt_pending:
            for bn in self.batch_norm_layers:
                bn.num_batches_tracked += self._nbt_pending
        self._nbt_pending = 0


class ResidualNet(nn.Module):
    def __init__(self, in_features, out_features, hidden_features, context_features=None, num_blocks=2, activation=F.relu,
                 dropout_probability=0.0, use_batch_norm=False):
        super().__init__()
        assert context_features is None
        self.hidden_features = hidden_features
        self.context_features = context_features
        self.initial_layer = nn.Linear(in_features, hidden_features)
        self.blocks = nn.ModuleList([ResidualBlock(features=hidden_features, context_features=None, activation=activation,
                                                   dropout_probability=dropout_probability, use_batch_norm=use_batch_norm)
                                     for _ in range(num_blocks)])
        self.final_layer = nn.Linear(hidden_features, out_features)

    def params(self):
        p = [self.initial_layer.weight, self.initial_layer.bias]
        for b in self.blocks:
            p += b.params()
        return p + [self.final_layer.weight, self.final_layer.bias]

    def fwd(self, x, training, masks):
        blocks = list(self.blocks)
        if training and blocks and ops.colslice_ok(x.shape[0]):
            # every Linear that feeds a BatchNorm1d computes that BatchNorm (+ReLU, +Dropout mask) in its epilogue
            h, t0, m0, r0 = ops.linear_bn_fwd(x, self.initial_layer.weight, self.initial_layer.bias, blocks[0].batch_norm_layers[0])
            ctxs, handoff = [], (t0, m0, r0)
            for j, blk in enumerate(blocks):
                nxt = blocks[j + 1].batch_norm_layers[0] if j + 1 < len(blocks) else None
                h, handoff, c = blk.fwd_fused(h, *handoff, None if masks is None else masks[j], nxt)
                ctxs.append(c)
            out = ops.linear_fwd(h, self.final_layer.weight, self.final_layer.bias)
            return out, (x, ctxs, h, True)
        h = ops.linear_fwd(x, self.initial_layer.weight, self.initial_layer.bias)
        ctxs = []
        for j, blk in enumerate(self.blocks):
            h_in = h
            h, c = blk.fwd(h_in, training, None if masks is None else masks[j])
            ctxs.append(c)
        out = ops.linear_fwd(h, self.final_layer.weight, self.final_layer.bias)
        return out, (x, ctxs, h, False)

    def bwd(self, dout, ctx, grads):
        x, ctxs, h_last, fused = ctx
        with ops.forked(dout, dout, h_last):
            grads[id(self.final_layer.weight)], grads[id(self.final_layer.bias)] = ops.linear_wgrad(dout, h_last)
        d = ops.linear_dgrad(dout, self.final_layer.weight)
        for blk, c in zip(reversed(list(self.blocks)), reversed(ctxs)):
            d = blk.bwd_fused(d, c, grads) if fused else blk.bwd(d, c, grads)
        with ops.forked(d, d, x):
            grads[id(self.initial_layer.weight)], grads[id(self.initial_layer.bias)] = ops.linear_wgrad(d, x)
        return ops.linear_dgrad(d, self.initial_layer.weight)


class Transform(nn.Module):
    pass


class AffineCouplingTransform(Transform):
    def __init__(self, mask, transform_net_create_fn, unconditional_transform=None):
        mask = torch.as_tensor(mask)
        if mask.dim() != 1:
            raise ValueError("Mask must be a 1-dim tensor.")
        if mask.numel() <= 0:
            raise ValueError("Mask can't be empty.")
        assert unconditional_transform is None
        super().__init__()
        self.features = len(mask)
        features_vector = torch.arange(self.features)
        self.register_buffer("identity_features", features_vector.masked_select(mask <= 0))
        self.register_buffer("transform_features", features_vector.masked_select(mask > 0))
        self.transform_net = transform_net_create_fn(self.num_identity_features, self.num_transform_features * 2)
        self.unconditional_transform = None
        self._idx32 = {}

    num_identity_features = property(lambda self: len(self.identity_features))
    num_transform_features = property(lambda self: len(self.transform_features))

    def _idx(self, device):
        key = (device.type, device.index)
        if key not in self._idx32:
            self._idx32[key] = (self.identity_features.to(device=device, dtype=torch.int32).contiguous(),
                                self.transform_features.to(device=device, dtype=torch.int32).contiguous())
        return self._idx32[key]

    def params(self):
        return self.transform_net.params()

    def fwd(self, x, logdet, training, masks):
        id_idx, tr_idx = self._idx(x.device)
        ident = ops.gather_cols(x, id_idx)
        prm, net_ctx = self.transform_net.fwd(ident, training, masks)
        y, ld = ops.coupling_fwd(x, prm, id_idx, tr_idx, logdet)
        return y, ld, (x, prm, net_ctx)

    def bwd(self, dy, dld, ctx, grads):
        x, prm, net_ctx = ctx
        id_idx, tr_idx = self._idx(x.device)
        dx, dprm = ops.coupling_bwd(dy, dld, x, prm, id_idx, tr_idx)
        dident = self.transform_net.bwd(dprm, net_ctx, grads)
        return ops.scatter_add_cols_(dx, id_idx, dident), dld

    def inv(self, y, logdet):
        id_idx, tr_idx = self._idx(y.device)
        prm, _ = self.transform_net.fwd(ops.gather_cols(y, id_idx), False, None)
        return ops.coupling_fwd(y, prm, id_idx, tr_idx, logdet, inverse=True)

    def inv_fwd(self, y, logdet, training, masks):
        """Inverse direction with the conditioner in its current mode, keeping what `inv_bwd` needs."""
        id_idx, tr_idx = self._idx(y.device)
        prm, net_ctx = self.transform_net.fwd(ops.gather_cols(y, id_idx), training, masks)
        x, ld = ops.coupling_fwd(y, prm, id_idx, tr_idx, logdet, inverse=True)
        return x, ld, (x, prm, net_ctx)

    def inv_bwd(self, dx, dld, ctx, grads):
        x, prm, net_ctx = ctx
        id_idx, tr_idx = self._idx(x.device)
        dy, dprm = ops.coupling_inv_bwd(dx, dld, x, prm, id_idx, tr_idx)
        dident = self.transform_net.bwd(dprm, net_ctx, grads)
        return ops.scatter_add_cols_(dy, id_idx, dident), dld


class BatchNorm(Transform):
    """Invertible batch-norm transform placed between regression-flow couplings (flows.py:87-88)."""

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

    def params(self):
        return [self.unconstrained_weight, self.bias]


def _add_scalar_to_rows(logdet, scalar, B):
    """logdet[b] + scalar (a 1-element device tensor) for every row; logdet None counts as zeros."""
    return ops.add_scalar(logdet, scalar, B, scalar)


class CompositeTransform(Transform):
    """Cascade of transforms; log|det J| accumulates (nflows transforms.base.CompositeTransform)."""

    def __init__(self, transforms):
        super().__init__()
        self._transforms = nn.ModuleList(transforms)

    def program_params(self):
        p = []
        for t in self._transforms:
            p += t.params()
        return p

    def flush_counters(self):
        for m in self.modules():
            if isinstance(m, ResidualBlock):
                m.flush_counters()

    def state_dict(self, *args, **kwargs):
        self.flush_counters()
        return super().state_dict(*args, **kwargs)

    def prog_fwd(self, inputs, training, extra):
        """extra: None, or a list (one entry per coupling) of per-block dropout masks."""
        x = inputs[0].contiguous()
        B = x.shape[0]
        rng = getattr(self, '_pgv_param_range', None)    # TrainStep: this flow's parameters as one contiguous slice of the flat buffer
        if rng is not None and training:
            ops.l2_prefetch(rng)
        if training:
            ops.mega_begin(x, B)                         # the whole chain as one persistent launch (ops.py, pgv_flow_program)
        try:
            return self._prog_fwd_body(x, B, training, extra, rng)
        finally:
            ops.mega_end()

    def _prog_fwd_body(self, x, B, training, extra, rng):
        logdet = None                                    # the first coupling starts the sum (NULL logdet_in)
        ctxs, c_i = [], 0
        for t in self._transforms:
            if isinstance(t, AffineCouplingTransform):
                masks = None if extra is None else extra[c_i]
                c_i += 1
                x, logdet, c = t.fwd(x, logdet, training, masks)
                ctxs.append(c)
            else:
                if training:
                    x_in = x
                    x, mean, var, ld = ops.flowbn_train_fwd(x_in, t)
                    ctxs.append((x_in, mean, var))
                else:
                    x, ld = ops.flowbn_eval(x, t)
                    ctxs.append(None)
                logdet = _add_scalar_to_rows(logdet, ld, B)
        if rng is not None and training:
            ops.join_forks(x)
        return (x, logdet), ctxs

    def prog_bwd(self, douts, ctxs, grads, needs):
        dy, dld = douts
        B = ctxs[0][0].shape[0]
        hook = getattr(self, 'on_backward_start', None)      # TrainStep: independent work that fills this latency-bound stretch
        if hook is not None:
            hook()
        rng = getattr(self, '_pgv_param_range', None)
        if rng is not None:
            ops.l2_prefetch(rng)
        if dy is None:
            dy = torch.zeros_like(ctxs[0][0])
        if dld is None:
            dld = torch.zeros(B, device=dy.device)
        ops.mega_begin(dy, B)
        try:
            return self._prog_bwd_body(dy, dld, B, ctxs, grads)
        finally:
            ops.mega_end()

    def _prog_bwd_body(self, dy, dld, B, ctxs, grads):
        for t, c in zip(reversed(list(self._transforms)), reversed(ctxs)):
            if isinstance(t, AffineCouplingTransform):
                dy, dld = t.bwd(dy, dld, c, grads)
            else:
                x_in, mean, var = c
                g_sum = ops.colsum(dld.view(B, 1))
                dy, du, db = ops.flowbn_train_bwd(dy, x_in, t, mean, var, g_sum)
                grads[id(t.unconstrained_weight)], grads[id(t.bias)] = du, db
        ops.join_forks(dy)                                 # the conditioners' weight gradients (ops.forked)
        return dy

    def forward(self, inputs, context=None, dropout_masks=None):
        """Returns (outputs, logabsdet).  dropout_masks: optional explicit masks (parity tests); when None and the
        module is training, masks are drawn for the blocks whose nn.Dropout has p > 0."""
        assert context is None
        if self.training and dropout_masks is None:
            dropout_masks = self._draw_masks(inputs)
        return run_program(self, (inputs,), self.program_params(), self.training, dropout_masks)

    def _draw_masks(self, inputs):
        from .encoder import make_dropout_mask
        all_masks, any_mask = [], False
        for t in self._transforms:
            if isinstance(t, AffineCouplingTransform):
                ms = []
                for blk in t.transform_net.blocks:
                    p = blk.dropout.p
                    if p > 0.0:
                        ms.append(make_dropout_mask((inputs.shape[0], t.transform_net.hidden_features), p, inputs.device))
                        any_mask = True
                    else:
                        ms.append(None)
                all_masks.append(ms)
        return all_masks if any_mask else None

    def inverse(self, inputs, context=None, dropout_masks=None):
        """Inverse pass.  Under torch.no_grad() / in eval mode: kernels only, no autograd.  In training mode with autograd enabled it
        is ONE differentiable autograd node (FlowParamsLoss back-propagates through the inverse latent flow, model/loss.py:318-346);
        like nflows, a cascade that contains BatchNorm transforms has no training-mode inverse."""
        assert context is None
        if self.training and any(isinstance(t, BatchNorm) for t in self._transforms):
            raise RuntimeError("Batch norm inverse is only available in eval mode, not in training mode.")
        if self.training and torch.is_grad_enabled():
            if dropout_masks is None:
                dropout_masks = self._draw_masks(inputs)
            return run_program(_InverseProgram(self), (inputs,), self.program_params(), True, dropout_masks)
        with torch.no_grad():
            return self._inverse_eval(inputs)

    def _inverse_eval(self, inputs):
        x = inputs.contiguous()
        B = x.shape[0]
        logdet = None
        for t in reversed(list(self._transforms)):
            if isinstance(t, AffineCouplingTransform):
                x, logdet = t.inv(x, logdet)
            else:
                x, ld = ops.flowbn_eval(x, t, inverse=True)
                logdet = _add_scalar_to_rows(logdet, ld, B)
        return x, logdet


class _InverseProgram:
    """Differentiable inverse of a CompositeTransform made of couplings only (training mode)."""

    def __init__(self, composite):
        self.c = composite

    def prog_fwd(self, inputs, training, extra):
        x = inputs[0].contiguous()
        couplings = list(self.c._transforms)
        logdet, ctxs = None, []
        for i in range(len(couplings) - 1, -1, -1):
            masks = None if extra is None else extra[i]
            x, logdet, c = couplings[i].inv_fwd(x, logdet, training, masks)
            ctxs.append(c)
        return (x, logdet), ctxs

    def prog_bwd(self, douts, ctxs, grads, needs):
        dx, dld = douts
        B = ctxs[0][0].shape[0]
        if dx is None:
            dx = torch.zeros_like(ctxs[0][0])
        if dld is None:
            dld = torch.zeros(B, device=dx.device)
        couplings = list(self.c._transforms)
        for t, c in zip(couplings, reversed(ctxs)):          # forward ran couplings K-1 .. 0: unwind 0 .. K-1
            dx, dld = t.inv_bwd(dx, dld, c, grads)
        ops.join_forks(dx)
        return dx


def _resnet_factory(hidden_features, num_blocks, dropout_probability, use_batch_norm):
    def create_resnet(in_features, out_features):
        return ResidualNet(in_features, out_features, hidden_features=hidden_features, num_blocks=num_blocks,
                           activation=F.relu, dropout_probability=dropout_probability, use_batch_norm=use_batch_norm)
    return create_resnet


class SimpleRealNVP(nn.Module):
    """nflows.flows.realnvp.SimpleRealNVP as used at VAE.py:118-125: only `._transform` is kept by the caller."""

    def __init__(self, features, hidden_features, num_layers, num_blocks_per_layer, use_volume_preserving=False,
                 activation=F.relu, dropout_probability=0.0, batch_norm_within_layers=False, batch_norm_between_layers=False):
        super().__init__()
        assert not use_volume_preserving
        mask = torch.ones(features)
        mask[::2] = -1
        layers = []
        for _ in range(num_layers):
            layers.append(AffineCouplingTransform(mask=mask, transform_net_create_fn=_resnet_factory(
                hidden_features, num_blocks_per_layer, dropout_probability, batch_norm_within_layers)))
            mask *= -1
            if batch_norm_between_layers:
                layers.append(BatchNorm(features=features))
        self._transform = CompositeTransform(layers)


class CustomRealNVP(CompositeTransform):
    """model/flows.py:42-90: no dropout in the last two couplings, BatchNorm transform after all but the last two."""

    def __init__(self, features, hidden_features, num_layers, num_blocks_per_layer, use_volume_preserving=False,
                 activation=F.relu, dropout_probability=0.0, batch_norm_within_layers=False, batch_norm_between_layers=False):
        assert not use_volume_preserving
        mask = torch.ones(features)
        mask[::2] = -1
        layers = []
        for l in range(num_layers):
            p = dropout_probability if l < (num_layers - 2) else 0.0
            layers.append(AffineCouplingTransform(mask=mask, transform_net_create_fn=_resnet_factory(
                hidden_features, num_blocks_per_layer, p, batch_norm_within_layers)))
            mask *= -1
            if batch_norm_between_layers and l < (num_layers - 2):
                layers.append(BatchNorm(features=features))
        super().__init__(layers)


class InverseFlow(nn.Module):
    """Dead in the reference as well (flows.py:28-30)."""

    def __init__(self, flow):
        super().__init__()
        raise AssertionError("This class messes autograd graphs (or only pytorch summaries???) and will be removed")
