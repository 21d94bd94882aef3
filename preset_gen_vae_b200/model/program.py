"""One autograd node per top-level module.

Every compute module of this package implements a *program*: `prog_fwd(inputs, training, extra) -> (outputs, ctx)`
launches the forward kernels and returns whatever the backward needs in `ctx`; `prog_bwd(douts, ctx, grads, needs)`
launches the backward kernels, stores parameter gradients in `grads[id(param)]` and returns the input gradients.
`run_program` wraps one such pair in a single torch.autograd.Function, so autograd sees the encoder, each flow, the
decoder and each loss as ONE node whose backward is hand-written CUDA, instead of hundreds of eager ops.
"""
import torch


class _ProgramFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prog, training, extra, n_in, *tensors):
        inputs, params = tensors[:n_in], tensors[n_in:]
        outs, pctx = prog.prog_fwd(inputs, training, extra)
        ctx.prog, ctx.pctx, ctx.n_in, ctx.params, ctx.training = prog, pctx, n_in, params, training
        ctx.single = not isinstance(outs, tuple)
        return outs

    @staticmethod
    def backward(ctx, *douts):
        if not ctx.training:
            raise NotImplementedError("backward through %s in eval mode is not implemented (the reference only "
                                      "evaluates under torch.no_grad, train.py:263)" % type(ctx.prog).__name__)
        grads = {}
        douts = tuple(None if d is None else d.contiguous() for d in douts)
        dins = ctx.prog.prog_bwd(douts[0] if ctx.single else douts, ctx.pctx, grads, ctx.needs_input_grad[4:4 + ctx.n_in])
        if not isinstance(dins, tuple):
            dins = (dins,)
        ctx.pctx = None
        return (None, None, None, None, *dins, *[grads.get(id(p)) for p in ctx.params])


def run_program(prog, inputs, params, training, extra=None):
    """inputs: tuple of tensors (None allowed); params: list of nn.Parameters the program reads."""
    return _ProgramFn.apply(prog, training, extra, len(inputs), *inputs, *params)
