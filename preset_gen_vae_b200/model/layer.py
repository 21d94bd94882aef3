"""Conv2D / TConv2D blocks with the reference's interface and state_dict layout (model/layer.py:10-46):
`<prefix>conv` (or `<prefix>tconv`) -> `<prefix>act` -> `<prefix>bn`, i.e. convolution, LeakyReLU, BatchNorm2d.

The child modules are torch's own (nn.Conv2d, nn.ConvTranspose2d, nn.BatchNorm2d): they are used ONLY as parameter /
buffer containers, so that initialisation, `state_dict()` keys and tensor layouts are exactly the reference's and a
reference checkpoint loads unchanged.  Their `forward` is never called; the arithmetic is in libpgv.so:
   conv (+bias +LeakyReLU fused)  ->  BatchNorm2d (batch statistics, running-stat update)
and the backward runs BatchNorm-backward fused with LeakyReLU-backward, then weight- and data-gradient kernels.
"""
import torch
import torch.nn as nn

from . import ops


def _slope_of(activation):
    if isinstance(activation, nn.LeakyReLU):
        return float(activation.negative_slope)
    if isinstance(activation, nn.ReLU) or activation is nn.ReLU:
        return 0.0
    raise NotImplementedError("only (Leaky)ReLU activations are fused into the conv blocks")


class _BlockBase(nn.Sequential):
    transposed = False

    def _setup(self, conv, activation, name_prefix, batch_norm, out_ch):
        if batch_norm == 'before':
            raise NotImplementedError("batch_norm='before' is never used on the reference path (layer.py:22)")
        self._conv_name = name_prefix + ('tconv' if self.transposed else 'conv')
        self._bn_name = name_prefix + 'bn' if batch_norm == 'after' else None
        self.slope = _slope_of(activation)
        self.add_module(self._conv_name, conv)
        bn = nn.BatchNorm2d(out_ch)     # constructed unconditionally, like the reference (no RNG use)
        self.add_module(name_prefix + 'act', activation if isinstance(activation, nn.Module) else activation())
        if self._bn_name is not None:
            self.add_module(self._bn_name, bn)
        self._nbt_pending = 0

    @property
    def conv(self):
        return getattr(self, self._conv_name)

    @property
    def bn(self):
        return getattr(self, self._bn_name) if self._bn_name is not None else None

    def block_params(self):
        p = [self.conv.weight, self.conv.bias]
        if self.bn is not None:
            p += [self.bn.weight, self.bn.bias]
        return p

    def flush_counters(self):
        """num_batches_tracked is bookkeeping only (momentum is fixed): it is advanced lazily, not once per step."""
        if self._nbt_pending and self.bn is not None:
            self.bn.num_batches_tracked += self._nbt_pending
        self._nbt_pending = 0

    def _post(self, a, training, sums=None):
        if self.bn is None:
            return a, None, None
        if training:
            y, mean, rstd = ops.bn2d_train_fwd(a, self.bn, sums)
            self._nbt_pending += 1
            return y, mean, rstd
        return ops.bn2d_eval_fwd(a, self.bn), None, None

    def _pre_bwd(self, dy, a, mean, rstd, grads, raw_sums=None):
        """Returns (dz, bias gradient of the convolution or None if it still has to be computed from dz).  raw_sums: the BatchNorm
        backward's batch sums if the kernel that produced dy accumulated them (chain_bwd)."""
        if self.bn is None:
            return ops.lrelu_bwd(dy, a, self.slope), None
        dz, dg, db, dbias = ops.bn2d_train_bwd(dy, a, self.bn.weight, mean, rstd, self.slope, want_colsum=True, raw_sums=raw_sums)
        grads[id(self.bn.weight)] = dg
        grads[id(self.bn.bias)] = db
        return dz, dbias

    def forward(self, x):
        from .program import run_program
        return run_program(_SingleBlock(self), (x,), self.block_params(), self.training)


class _SingleBlock:
    """Program wrapper so that a lone block can be called like a module (tests, torchinfo-style summaries)."""

    def __init__(self, block):
        self.block = block

    def prog_fwd(self, inputs, training, extra):
        return self.block.fwd(inputs[0].contiguous(), training)

    def prog_bwd(self, dout, ctx, grads, needs):
        dx = self.block.bwd(dout, ctx, grads, needs[0])
        ops.join_forks(dout)
        return dx


class Conv2D(_BlockBase):
    """nn.Conv2d -> activation -> BatchNorm2d ('after'), or no BN when batch_norm is None (layer.py:10-26)."""

    def __init__(self, in_ch, out_ch, kernel_size, stride, padding, dilation, padding_mode='zeros', activation=nn.ReLU,
                 name_prefix='', batch_norm='after'):
        super().__init__()
        conv = nn.Conv2d(in_ch, out_ch, kernel_size, stride, padding, dilation, padding_mode=padding_mode)
        assert conv.dilation == (1, 1) and padding_mode == 'zeros' and conv.stride[0] == conv.stride[1] \
            and conv.padding[0] == conv.padding[1], "unsupported convolution geometry"
        self._setup(conv, activation, name_prefix, batch_norm, out_ch)

    def out_hw(self, h, w):
        k, s, p = self.conv.kernel_size, self.conv.stride[0], self.conv.padding[0]
        return ops.conv_out_size(h, k[0], s, p), ops.conv_out_size(w, k[1], s, p)

    def _route(self, x):
        c = self.conv
        Ho, Wo = self.out_hw(x.shape[2], x.shape[3])
        return ops.conv_route(c.in_channels, c.out_channels, c.kernel_size[0], c.kernel_size[1], c.stride[0], c.padding[0],
                              x.shape[2], x.shape[3], Ho, Wo)

    def fwd(self, x, training):
        c = self.conv
        wf = wq = None
        if self._route(x) == 'cl':         # channels-last tensor-core route: TF32-rounded input and re-packed weights
            x = ops.to_cl(x, round_out=True)
            wf, wq = ops.prep_conv_weights(c.weight, c.stride[0], c.padding[0], dgrad=training)
        # without a BatchNorm behind it, the activation itself is the next tensor-core operand
        want_sums = training and self.bn is not None       # the conv epilogue also accumulates the BatchNorm statistics
        a = ops.conv2d_fwd(x, c.weight, c.bias, c.stride[0], c.padding[0], self.slope, wf=wf, round_out=self.bn is None, bn_sums=want_sums)
        a, sums = a if want_sums else (a, None)
        y, mean, rstd = self._post(a, training, sums)
        return y, (x, a, mean, rstd, wq)

    def bwd(self, dy, ctx, grads, need_dx=True, sums=None, prev_a=None):
        """sums / prev_a: see chain_bwd; with prev_a the result is (dx, sums of the previous block's BatchNorm backward or None)."""
        x, a, mean, rstd, wq = ctx
        c = self.conv
        dz, dbias = self._pre_bwd(dy, a, mean, rstd, grads, sums)
        direct = ops.grad_out_of(c.weight)          # written in place into the flat gradient buffer: autograd gets no tensor for it
        with ops.forked(x, x, dz):                   # joined by the owning module at the end of its backward
            dw, db = ops.conv2d_wgrad(x, dz, c.weight.shape, c.stride[0], c.padding[0], want_bias=True, db=dbias, out=direct)
        grads[id(c.weight)], grads[id(c.bias)] = (None if direct is not None else dw), db
        if not need_dx:
            return None
        return ops.conv2d_dgrad(dz, c.weight, x.shape[2:], c.stride[0], c.padding[0], wq=wq, bn_bwd_x=prev_a)


class TConv2D(_BlockBase):
    """nn.ConvTranspose2d -> activation -> BatchNorm2d (layer.py:29-46).  The transposed convolution is evaluated as
    the data-gradient of the convolution that has the same weight tensor [Cin, Cout, kh, kw]."""
    transposed = True

    def __init__(self, in_ch, out_ch, kernel_size, stride, padding, output_padding=0, dilation=1, padding_mode='zeros',
                 activation=nn.ReLU, name_prefix='', batch_norm='after'):
        super().__init__()
        conv = nn.ConvTranspose2d(in_ch, out_ch, kernel_size, stride, padding, output_padding, dilation=dilation,
                                  padding_mode=padding_mode)
        assert conv.dilation == (1, 1) and conv.stride[0] == conv.stride[1] and conv.padding[0] == conv.padding[1]
        self._setup(conv, activation, name_prefix, batch_norm, out_ch)

    def out_hw(self, h, w):
        return tconv_out_hw(self.conv, h, w)

    def fwd(self, x, training):
        c = self.conv
        wf = wq = None
        if tconv_route(x, c) == 'cl':
            x = ops.to_cl(x, round_out=True)
            wf, wq = ops.prep_conv_weights(c.weight, c.stride[0], c.padding[0], fwd=training)
        want_sums = training and self.bn is not None
        a = tconv_fwd(x, c, self.slope, wq=wq, round_out=self.bn is None, bn_sums=want_sums)
        a, sums = a if want_sums else (a, None)
        y, mean, rstd = self._post(a, training, sums)
        return y, (x, a, mean, rstd, wf)

    def bwd(self, dy, ctx, grads, need_dx=True, sums=None, prev_a=None):
        x, a, mean, rstd, wf = ctx
        dz, dbias = self._pre_bwd(dy, a, mean, rstd, grads, sums)
        return tconv_bwd(dz, x, self.conv, grads, need_dx, wf=wf, dbias=dbias, bn_bwd_x=prev_a)


def tconv_out_hw(conv, h, w):
    k, s, p, op = conv.kernel_size, conv.stride[0], conv.padding[0], conv.output_padding
    return (h - 1) * s - 2 * p + k[0] + op[0], (w - 1) * s - 2 * p + k[1] + op[1]


def tconv_route(x, conv):
    """Kernel family of the transposed convolution = that of the convolution with the same weight tensor [Cin_t, Cout_t, kh, kw]."""
    cin_t, cout_t, kh, kw = conv.weight.shape
    H, W = tconv_out_hw(conv, x.shape[2], x.shape[3])
    return ops.conv_route(cout_t, cin_t, kh, kw, conv.stride[0], conv.padding[0], H, W, x.shape[2], x.shape[3])


def tconv_fwd(x, conv, slope=-1.0, clamp=None, wq=None, round_out=False, bn_sums=False):
    """ConvTranspose2d forward (+bias, optional fused LeakyReLU or Hardtanh clamp) = conv data-gradient."""
    return ops.conv2d_dgrad(x, conv.weight, tconv_out_hw(conv, x.shape[2], x.shape[3]), conv.stride[0], conv.padding[0],
                            bias=conv.bias, slope=slope, clamp=clamp, wq=wq, round_out=round_out, bn_sums=bn_sums)


def tconv_clamp_fusable(x, conv):
    """True when the transposed convolution is the thin 5x5 / one-output-channel layer whose kernel fuses the Hardtanh."""
    cin_t, cout_t, kh, kw = conv.weight.shape
    H, W = tconv_out_hw(conv, x.shape[2], x.shape[3])
    return ops.use_thin and ops._thin(cout_t, cin_t, kh, kw, conv.stride[0], conv.padding[0], H, W, x.shape[2], x.shape[3])


def tconv_bwd(dz, x, conv, grads, need_dx=True, wf=None, dbias=None, bn_bwd_x=None):
    """dz: gradient w.r.t. the transposed convolution's (pre-activation) output; dbias: its per-channel sums if already known.
    bn_bwd_x: as in ops.conv2d_fwd (the result is then (dx, sums))."""
    direct = ops.grad_out_of(conv.weight)
    with ops.forked(x, x, dz):
        dw, _ = ops.conv2d_wgrad(dz, x, conv.weight.shape, conv.stride[0], conv.padding[0], want_bias=False, out=direct)
        grads[id(conv.bias)] = dbias if dbias is not None else ops.channel_sum(dz)
    grads[id(conv.weight)] = None if direct is not None else dw
    if not need_dx:
        return None
    return ops.conv2d_fwd(dz, conv.weight, None, conv.stride[0], conv.padding[0], -1.0, out_hw=x.shape[2:], wf=wf, bn_bwd_x=bn_bwd_x)


def chain_bwd(blocks, ctxs, d, grads, need_first_dx=True):
    """Backward through a sequential chain of Conv2D / TConv2D blocks (block i's input is block i - 1's output).  The gradient that
    block i's data-gradient kernel writes is the gradient flowing into block i - 1's BatchNorm2d, so that kernel's epilogue also
    accumulates the two batch sums that BatchNorm backward needs (ops.conv2d_fwd / conv2d_dgrad, bn_bwd_x) and block i - 1 skips its
    reduction pass over two activation-sized tensors.  Launches that cannot (several N tiles, other kernel families) return no sums
    and the block falls back to its own reduction."""
    sums = None
    for i in range(len(blocks) - 1, -1, -1):
        prev_a = ctxs[i - 1][1] if (i > 0 and ops.fuse_bn_bwd and blocks[i - 1].bn is not None and ctxs[i - 1][2] is not None) else None
        out = blocks[i].bwd(d, ctxs[i], grads, need_first_dx or i > 0, sums=sums, prev_a=prev_a)
        d, sums = out if prev_a is not None else (out, None)
    return d
