"""Spectrogram encoder with the reference's interface and state_dict layout (model/encoder.py:23-108, 111-306).

`SpectrogramEncoder(architecture, dim_z, input_tensor_size, fc_dropout, output_bn, deepest_features_mix,
force_bigger_network)(x)` maps `[B, C, 257, 347]` spectrograms to `[B, 2, dim_z]` (mu, log-variance).  Only the
'speccnn8l1_bn' architecture is implemented, the one the reference fully supports (encoder.py:53).

The whole forward (shared per-channel CNN, features mixer, dropout, Linear, BatchNorm1d) is one autograd node whose
forward and backward launch libpgv.so kernels.
"""
import torch
import torch.nn as nn

from . import layer, ops
from .program import run_program


class SpectrogramCNN(nn.Module):
    """Per-channel CNN (encoder.py:111-306, 'speccnn8l1_bn' branch at 233-259)."""

    def __init__(self, architecture, last_layers_to_remove=0):
        super().__init__()
        self.architecture = architecture
        if architecture != 'speccnn8l1_bn':
            raise NotImplementedError("Architecture '{}' not available (only 'speccnn8l1_bn' is fully supported by the "
                                      "reference, encoder.py:53)".format(architecture))
        act, act_p = nn.LeakyReLU, 0.1
        chans = (1, 8, 16, 32, 64, 128, 256)
        blocks = [layer.Conv2D(1, 8, [5, 5], [2, 2], 2, [1, 1], batch_norm=None, activation=act(act_p), name_prefix='enc1')]
        for i in range(1, 6):
            blocks.append(layer.Conv2D(chans[i], chans[i + 1], [4, 4], [2, 2], 2, [1, 1], activation=act(act_p),
                                       name_prefix='enc%d' % (i + 1)))
        self.enc_nn = nn.Sequential(*blocks)
        if last_layers_to_remove <= 1:
            self.enc_nn.add_module('4x4conv', layer.Conv2D(256, 512, [4, 4], [2, 2], 2, [1, 1], activation=act(act_p),
                                                           name_prefix='enc7'))
        if last_layers_to_remove == 0:
            self.enc_nn.add_module('1x1conv', layer.Conv2D(512, 1024, [1, 1], [1, 1], 0, [1, 1], batch_norm=None,
                                                           activation=act(act_p), name_prefix='enc8'))

    def blocks(self):
        return list(self.enc_nn.children())

    def forward(self, x_spectrogram):
        for b in self.blocks():
            x_spectrogram = b(x_spectrogram)
        return x_spectrogram


def _constructor_side_effect(single_blocks, mixer_blocks, input_tensor_size):
    """The reference infers the CNN output size with a dummy forward of zeros in TRAINING mode (encoder.py:73-78),
    which as a side effect updates every BatchNorm's running statistics (once per channel for the shared CNN).  The
    same dummy forward is run here with the libpgv.so kernels on device copies of the freshly initialised weights,
    and the resulting running statistics are copied back.  Without a GPU there is nothing to run it on (no CPU
    path): the statistics keep torch's defaults and False is returned; loading a checkpoint overrides them anyway."""
    if not torch.cuda.is_available():
        return False
    from types import SimpleNamespace
    dev = torch.device('cuda', torch.cuda.current_device())
    C = input_tensor_size[1]

    def run(blk, h):
        c = blk.conv
        a = ops.conv2d_fwd(h, c.weight.detach().to(dev), c.bias.detach().to(dev), c.stride[0], c.padding[0], blk.slope)
        bn = blk.bn
        if bn is None:
            return a
        shadow = SimpleNamespace(weight=bn.weight.detach().to(dev), bias=bn.bias.detach().to(dev),
                                 running_mean=bn.running_mean.to(dev), running_var=bn.running_var.to(dev),
                                 momentum=bn.momentum, eps=bn.eps)
        y, _, _ = ops.bn2d_train_fwd(a, shadow)
        bn.running_mean.copy_(shadow.running_mean)
        bn.running_var.copy_(shadow.running_var)
        bn.num_batches_tracked += 1
        return y
    precision = ops.get_precision()
    ops.set_precision('fp32')            # one tiny forward: exact arithmetic, so the statistics match the reference's to fp32 rounding
    with torch.no_grad():
        x = torch.zeros(1, C, input_tensor_size[2], input_tensor_size[3], device=dev)
        feats = []
        for ch in range(C):
            h = x[:, ch:ch + 1].contiguous()
            for blk in single_blocks:
                h = run(blk, h)
            feats.append(h)
        h = feats[0] if C == 1 else ops.cat_channels(feats)
        for blk in mixer_blocks:
            h = run(blk, h)
    ops.set_precision(precision)
    return True


class SpectrogramEncoder(nn.Module):
    def __init__(self, architecture, dim_z, input_tensor_size, fc_dropout, output_bn=False, deepest_features_mix=True,
                 force_bigger_network=False):
        super().__init__()
        self.dim_z = dim_z
        self.spectrogram_channels = input_tensor_size[1]
        self.architecture = architecture
        self.deepest_features_mix = deepest_features_mix
        self.mixer_1x1conv_ch = 1024 if (self.spectrogram_channels > 1) else 2048
        self.fc_dropout = fc_dropout
        self.single_ch_cnn = SpectrogramCNN(self.architecture, last_layers_to_remove=(1 if deepest_features_mix else 2))
        assert self.architecture == 'speccnn8l1_bn'
        C = self.spectrogram_channels
        if self.deepest_features_mix:
            self.features_mixer_cnn = layer.Conv2D(512 * C, self.mixer_1x1conv_ch, [1, 1], [1, 1], 0, [1, 1],
                                                   activation=nn.LeakyReLU(0.1), name_prefix='enc8', batch_norm=None)
        else:
            n_4x4_ch = 1800 if force_bigger_network else (512 if C == 1 else 768)
            self.features_mixer_cnn = nn.Sequential(
                layer.Conv2D(256 * C, n_4x4_ch, [4, 4], [2, 2], 2, [1, 1], activation=nn.LeakyReLU(0.1), name_prefix='enc7'),
                layer.Conv2D(n_4x4_ch, self.mixer_1x1conv_ch, [1, 1], [1, 1], 0, [1, 1], activation=nn.LeakyReLU(0.1),
                             name_prefix='enc8', batch_norm=None))
        self.constructor_bn_side_effect_applied = _constructor_side_effect(self.single_ch_cnn.blocks(), self._mixer_blocks(),
                                                                           input_tensor_size)
        h, w = input_tensor_size[2], input_tensor_size[3]
        for blk in self.single_ch_cnn.blocks() + self._mixer_blocks():
            h, w = blk.out_hw(h, w)
        self.cnn_out_size = torch.Size((1, self.mixer_1x1conv_ch, h, w))
        cnn_out_items = self.mixer_1x1conv_ch * h * w
        self.mlp = nn.Sequential(nn.Dropout(self.fc_dropout), nn.Linear(cnn_out_items, 2 * self.dim_z))
        if output_bn:
            self.mlp.add_module('lat_in_regularization', nn.BatchNorm1d(2 * self.dim_z))
        self._nbt_pending = 0

    def _mixer_blocks(self):
        m = self.features_mixer_cnn
        return [m] if isinstance(m, layer.Conv2D) else list(m.children())

    @property
    def out_bn(self):
        return getattr(self.mlp, 'lat_in_regularization', None)

    def flush_counters(self):
        for b in self.single_ch_cnn.blocks() + self._mixer_blocks():
            b.flush_counters()
        if self._nbt_pending and self.out_bn is not None:
            self.out_bn.num_batches_tracked += self._nbt_pending
        self._nbt_pending = 0

    def state_dict(self, *args, **kwargs):
        self.flush_counters()
        return super().state_dict(*args, **kwargs)

    def program_params(self):
        p = []
        for b in self.single_ch_cnn.blocks() + self._mixer_blocks():
            p += b.block_params()
        p += [self.mlp[1].weight, self.mlp[1].bias]
        if self.out_bn is not None:
            p += [self.out_bn.weight, self.out_bn.bias]
        return p

    # ---- program ----
    def prog_fwd(self, inputs, training, extra):
        x = inputs[0].contiguous()
        drop_mask = inputs[1] if len(inputs) > 1 else None
        B, C = x.shape[0], self.spectrogram_channels
        ch_ctx, feats = [], []
        for ch in range(C):                              # shared CNN, once per channel (encoder.py:97-98)
            h = x if C == 1 else x[:, ch:ch + 1].contiguous()
            ctxs = []
            for blk in self.single_ch_cnn.blocks():
                h, c = blk.fwd(h, training)
                ctxs.append(c)
            ch_ctx.append(ctxs)
            feats.append(h)
        h = feats[0] if C == 1 else ops.cat_channels(feats)
        mix_ctx = []
        for blk in self._mixer_blocks():
            h, c = blk.fwd(h, training)
            mix_ctx.append(c)
        cnn_shape = h.shape
        flat = ops.to_nchw(h).reshape(B, -1)          # nn.Linear expects the (c, h, w) flattening order
        fc_in = ops.mul(flat, drop_mask) if (training and drop_mask is not None) else flat
        lin = self.mlp[1]
        y, fc_ctx = ops.fc_fwd(fc_in, lin.weight, lin.bias, training)
        bn_ctx = None
        if self.out_bn is not None:
            if training:
                y_pre = y
                y, mean, rstd = ops.bn1d_train_fwd(y_pre, self.out_bn)
                self._nbt_pending += 1
                bn_ctx = (y_pre, mean, rstd)
            else:
                y = ops.bn1d_eval_fwd(y, self.out_bn)
        return y.view(B, 2, self.dim_z), (ch_ctx, mix_ctx, cnn_shape, fc_ctx, drop_mask, bn_ctx)

    def prog_bwd(self, dout, ctx, grads, needs):
        ch_ctx, mix_ctx, cnn_shape, fc_ctx, drop_mask, bn_ctx = ctx
        B, C = dout.shape[0], self.spectrogram_channels
        dy = dout.reshape(B, -1)
        if bn_ctx is not None:
            y_pre, mean, rstd = bn_ctx
            dy, dg, db = ops.bn1d_train_bwd(dy, y_pre, self.out_bn, mean, rstd)
            grads[id(self.out_bn.weight)], grads[id(self.out_bn.bias)] = dg, db
        lin = self.mlp[1]
        # fc_weight_grad_out (set by TrainStep): the 30 M-element weight gradient is written straight into the flat gradient
        # buffer instead of a temporary that would be copied there; autograd then gets no tensor for it
        direct = getattr(self, 'fc_weight_grad_out', None)
        dflat, dw, db = ops.fc_bwd(dy, fc_ctx, lin.weight, True, out=direct)
        # TrainStep: both FC weight gradients (the decoder's backward has already joined) are final here, and so is every gradient
        # outside the encoder: a hook packs / updates them under the rest of this backward, an event releases their exchange
        hook = getattr(self, 'before_fc_grads_ready', None)
        if hook is not None:
            hook()
        ready = getattr(self, 'fc_grads_ready_event', None)
        if ready is not None:
            ready.record()
        grads[id(lin.weight)], grads[id(lin.bias)] = (None if direct is not None else dw), db
        if drop_mask is not None:
            dflat = ops.mul(dflat, drop_mask)
        dh = dflat.view(cnn_shape)
        dh = layer.chain_bwd(self._mixer_blocks(), mix_ctx, dh, grads, True)
        need_dx = bool(needs[0])
        dxs = []
        per = dh.shape[1] // C
        blocks = self.single_ch_cnn.blocks()
        for ch in range(C):
            d = dh if C == 1 else ops.slice_channels(dh, ch * per, (ch + 1) * per)
            local = {}
            d = layer.chain_bwd(blocks, ch_ctx[ch], d, local, need_dx)
            if C > 1:
                ops.join_forks(dout)                       # the sums below read gradients produced on the child stream
            for k, v in local.items():                   # the CNN is shared: sum parameter gradients over channels
                grads[k] = v if k not in grads else ops.add(grads[k], v)
            dxs.append(d)
        dx = None
        if need_dx:
            dx = dxs[0] if C == 1 else torch.cat(dxs, dim=1)
        ops.join_forks(dout)                               # the weight gradients enqueued on the child stream (ops.forked)
        return (dx, None)[:len(needs)] if len(needs) > 1 else dx

    def _forward_cnns(self, x_spectrograms):
        raise NotImplementedError("use forward(); the CNN and the MLP head run as one fused program")

    def forward(self, x_spectrograms, dropout_mask=None):
        """dropout_mask: optional pre-scaled keep mask [B, cnn_out_items] (parity tests share it with the oracle);
        when None and training, it is drawn from torch's CUDA generator like nn.Dropout would."""
        if self.training and dropout_mask is None and self.fc_dropout > 0.0:
            dropout_mask = make_dropout_mask((x_spectrograms.shape[0], self.mlp[1].in_features), self.fc_dropout,
                                             x_spectrograms.device)
        inputs = (x_spectrograms,) if dropout_mask is None else (x_spectrograms, dropout_mask)
        return run_program(self, inputs, self.program_params(), self.training)


def make_dropout_mask(shape, p, device):
    """Pre-scaled keep mask, drawn from torch's generator for `device` (RNG is plumbing; the multiply is a pgv kernel)."""
    return torch.empty(shape, device=device).bernoulli_(1.0 - p).div_(1.0 - p)
