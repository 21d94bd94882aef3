"""Spectrogram decoder with the reference's interface and state_dict layout (model/decoder.py:9-92, 95-274).

`SpectrogramDecoder(architecture, dim_z, output_tensor_size, fc_dropout, force_bigger_network)(z)` maps `[B, dim_z]`
latent vectors to `[B, C, 257, 347]` spectrograms in [-1, 1]: Linear -> Dropout -> 1x1 transposed conv "un-mixer" ->
shared per-channel transposed-conv CNN -> Hardtanh.  One autograd node, all arithmetic in libpgv.so.
"""
import numpy as np
import torch
import torch.nn as nn

from . import layer, ops
from .encoder import make_dropout_mask
from .program import run_program


class SpectrogramCNN(nn.Module):
    """Per-channel decoder CNN (decoder.py:95-274, 'speccnn8l1(_bn)' branch at 199-220)."""

    def __init__(self, architecture, spectrogram_input_size, output_activation=nn.Hardtanh(), append_1x1_conv=True,
                 force_bigger_network=False):
        super().__init__()
        self.architecture = architecture
        if not append_1x1_conv:
            assert self.architecture == 'speccnn8l1_bn'
        self.spectrogram_input_size = spectrogram_input_size
        assert self.spectrogram_input_size[1] == 1
        if architecture not in ('speccnn8l1', 'speccnn8l1_bn'):
            raise NotImplementedError("Architecture '{}' not available".format(architecture))
        assert not append_1x1_conv, "the reference asserts False on this branch (decoder.py:221)"
        act, act_p = nn.LeakyReLU, 0.1
        spec = ((512 if not force_bigger_network else 1800, 256, [1, 1]), (256, 128, [1, 0]), (128, 64, [1, 1]),
                (64, 32, [1, 1]), (32, 16, [1, 0]), (16, 8, [1, 0]))
        mods = [layer.TConv2D(cin, cout, [4, 4], [2, 2], 2, output_padding=op, activation=act(act_p),
                              name_prefix='dec%d' % (i + 2)) for i, (cin, cout, op) in enumerate(spec)]
        assert isinstance(output_activation, nn.Hardtanh)
        self.dec_nn = nn.Sequential(*mods, nn.ConvTranspose2d(8, 1, [5, 5], [2, 2], 2), output_activation)

    def blocks(self):
        return list(self.dec_nn.children())[:-2]

    @property
    def last_tconv(self):
        return self.dec_nn[-2]

    @property
    def out_act(self):
        return self.dec_nn[-1]

    def fwd(self, h, training):
        ctxs = []
        for blk in self.blocks():
            h, c = blk.fwd(h, training)
            ctxs.append(c)
        lo, hi = self.out_act.min_val, self.out_act.max_val
        if layer.tconv_clamp_fusable(h, self.last_tconv):
            # Hardtanh fused into the thin transposed-conv kernel; y in (lo, hi) <=> the pre-activation is, so the
            # backward mask can be taken from y itself
            y = layer.tconv_fwd(h, self.last_tconv, clamp=(lo, hi))
            return y, (ctxs, h, y)
        pre = layer.tconv_fwd(h, self.last_tconv)
        y = ops.hardtanh_fwd(pre, lo, hi)
        return y, (ctxs, h, pre)

    def bwd(self, dy, ctx, grads):
        ctxs, h_last, pre = ctx
        d = ops.hardtanh_bwd(dy, pre, self.out_act.min_val, self.out_act.max_val)
        d = layer.tconv_bwd(d, h_last, self.last_tconv, grads, True)
        return layer.chain_bwd(self.blocks(), ctxs, d, grads, True)

    def forward(self, x_spectrogram):
        raise NotImplementedError("the per-channel CNN runs inside SpectrogramDecoder's fused program")


class SpectrogramDecoder(nn.Module):
    def __init__(self, architecture, dim_z, output_tensor_size, fc_dropout, force_bigger_network=False):
        super().__init__()
        self.output_tensor_size = output_tensor_size
        self.spectrogram_input_size = (self.output_tensor_size[2], self.output_tensor_size[3])
        self.spectrogram_channels = output_tensor_size[1]
        self.dim_z = dim_z
        self.architecture = architecture
        self.mixer_1x1conv_ch = 2048
        self.last_4x4conv_ch = (512 if not force_bigger_network else 1800)
        self.fc_dropout = fc_dropout
        if architecture != 'speccnn8l1_bn':
            raise NotImplementedError("Only speccnn8l1_bn is available (stacked multi-note spectrograms compatibility, "
                                      "decoder.py:35-37, 104)")
        if self.spectrogram_input_size != (257, 347):
            raise NotImplementedError("only the (257, 347) spectrogram size is defined for this decoder (decoder.py:58-68)")
        self.cnn_input_shape = (self.mixer_1x1conv_ch, 3, 4)
        self.mlp = nn.Sequential(nn.Linear(self.dim_z, int(np.prod(self.cnn_input_shape))), nn.Dropout(self.fc_dropout))
        self.features_unmixer_cnn = layer.TConv2D(self.mixer_1x1conv_ch, self.spectrogram_channels * self.last_4x4conv_ch,
                                                  [1, 1], [1, 1], 0, activation=nn.LeakyReLU(0.1), name_prefix='dec1')
        single_spec_size = list(self.spectrogram_input_size)
        single_spec_size[1] = 1
        self.single_ch_cnn = SpectrogramCNN(self.architecture, single_spec_size, append_1x1_conv=False,
                                            force_bigger_network=force_bigger_network)

    def flush_counters(self):
        for b in [self.features_unmixer_cnn] + self.single_ch_cnn.blocks():
            b.flush_counters()

    def state_dict(self, *args, **kwargs):
        self.flush_counters()
        return super().state_dict(*args, **kwargs)

    def program_params(self):
        p = [self.mlp[0].weight, self.mlp[0].bias] + self.features_unmixer_cnn.block_params()
        for b in self.single_ch_cnn.blocks():
            p += b.block_params()
        return p + [self.single_ch_cnn.last_tconv.weight, self.single_ch_cnn.last_tconv.bias]

    def prog_fwd(self, inputs, training, extra):
        z = inputs[0].contiguous()
        drop_mask = inputs[1] if len(inputs) > 1 else None
        lin = self.mlp[0]
        h, fc_ctx = ops.fc_fwd(z, lin.weight, lin.bias, training)
        if training and drop_mask is not None:
            h = ops.mul(h, drop_mask)
        h = h.view(-1, *self.cnn_input_shape)
        h, un_ctx = self.features_unmixer_cnn.fwd(h, training)
        C = self.spectrogram_channels
        outs, ctxs = [], []
        for ch in range(C):                              # shared CNN, once per 512-channel slice (decoder.py:89-91)
            part = h if C == 1 else ops.slice_channels(h, ch * self.last_4x4conv_ch, (ch + 1) * self.last_4x4conv_ch)
            y, c = self.single_ch_cnn.fwd(part, training)
            outs.append(y)
            ctxs.append(c)
        x_out = outs[0] if C == 1 else torch.cat(outs, dim=1)
        return x_out, (fc_ctx, drop_mask, un_ctx, ctxs)

    def prog_bwd(self, dout, ctx, grads, needs):
        fc_ctx, drop_mask, un_ctx, ctxs = ctx
        C = self.spectrogram_channels
        dparts = []
        for ch in range(C):
            d = dout if C == 1 else dout[:, ch:ch + 1].contiguous()
            local = {}
            dparts.append(self.single_ch_cnn.bwd(d, ctxs[ch], local))
            if C > 1:
                ops.join_forks(dout)                       # the sums below read gradients produced on the child stream
            for k, v in local.items():
                grads[k] = v if k not in grads else ops.add(grads[k], v)
        dh = dparts[0] if C == 1 else ops.cat_channels(dparts)
        dh = self.features_unmixer_cnn.bwd(dh, un_ctx, grads, True)
        dflat = ops.to_nchw(dh).reshape(dh.shape[0], -1)
        if drop_mask is not None:
            dflat = ops.mul(dflat, drop_mask)
        lin = self.mlp[0]
        # fc_weight_grad_out (set by TrainStep): the 30 M-element weight gradient is written straight into the flat gradient
        # buffer instead of a temporary that would be copied there; autograd then gets no tensor for it
        direct = getattr(self, 'fc_weight_grad_out', None)
        dz, dw, db = ops.fc_bwd(dflat, fc_ctx, lin.weight, bool(needs[0]), out=direct)
        grads[id(lin.weight)], grads[id(lin.bias)] = (None if direct is not None else dw), db
        ops.join_forks(dout)                               # the weight gradients enqueued on the child stream (ops.forked)
        return (dz, None)[:len(needs)] if len(needs) > 1 else dz

    def forward(self, z_sampled, dropout_mask=None):
        if self.training and dropout_mask is None and self.fc_dropout > 0.0:
            dropout_mask = make_dropout_mask((z_sampled.shape[0], self.mlp[0].out_features), self.fc_dropout, z_sampled.device)
        inputs = (z_sampled,) if dropout_mask is None else (z_sampled, dropout_mask)
        return run_program(self, inputs, self.program_params(), self.training)
