"""Thin Python wrappers over the C ABI (include/pgv.h) for the model path.

Each function allocates its outputs with torch (PyTorch owns all device memory), passes raw pointers, sizes and the
current CUDA stream to ONE libpgv.so entry point and returns the tensors.  No arithmetic happens in Python and there
is no fallback: a non-CUDA tensor raises.

`set_precision('tf32' | 'fp32')` selects how GEMM-shaped layers multiply: 'tf32' (default) uses the tcgen05 tensor-core
kernels (TF32 products, fp32 accumulation, what north_star asks for); 'fp32' routes the same layers to the exact-fp32
CUDA-core kernels and is what the tight parity tests use.
"""
import ctypes

import numpy as np
import torch

from .. import _lib

LRELU_SLOPE = 0.1
use_thin = True       # route the 5x5 / one-channel layers (enc1, dec8) to the direct streaming kernels (exact fp32)
use_cl = True         # 'tf32' precision: run the 4x4 / 1x1 conv blocks channels-last on the cp.async-fed tcgen05 kernels
_precision = 'tf32'
launches = 0          # kernels launched through this module since the last reset (bench.py's gpu_launches)


def set_precision(p):
    global _precision
    assert p in ('tf32', 'fp32')
    _precision = p


def get_precision():
    return _precision


def _f(t):
    if t is None:
        return ctypes.c_void_p(0)
    if _mega is not None:
        _mega_keep.append(t)          # recorded ops run later: nothing they touch may be recycled by the allocator before the flush
    if not t.is_cuda:
        raise _lib.PgvError("pgv kernels need CUDA tensors (got %s): there is no CPU path" % t.device)
    assert t.dtype in (torch.float32, torch.int32, torch.float64, torch.uint8), t.dtype
    assert t.is_contiguous() or is_cl(t), "pgv kernels take dense NCHW or channels-last tensors"
    return ctypes.c_void_p(t.data_ptr())


# ------------------------------------------------------------------------------------------------ layouts
# Channels-last activations are ordinary torch tensors of logical shape [B, C, H, W] with torch.channels_last strides, so
# shapes read the same everywhere in Python and only the kernels see the physical [B, H, W, C] order.
def cl_mode():
    return _precision == 'tf32' and use_cl


def is_cl(t):
    return t.dim() == 4 and not t.is_contiguous() and t.is_contiguous(memory_format=torch.channels_last)


def _empty_cl(ref, B, C, H, W):
    return torch.empty((B, C, H, W), dtype=torch.float32, device=ref.device, memory_format=torch.channels_last)


def to_cl(x, round_out=False):
    """[B, C, H, W] (NCHW memory) -> the same logical tensor in channels-last memory, optionally rounded to TF32."""
    if is_cl(x):
        return x
    x = x.contiguous()
    B, C, H, W = x.shape
    y = _empty_cl(x, B, C, H, W)
    _call('pgv_transpose_inner', _f(x), _f(y), B, C, H * W, int(round_out), _s(x), nbytes=8 * x.numel())
    return y


def to_nchw(x):
    """Channels-last memory -> NCHW memory (no-op for tensors that already are)."""
    if x.is_contiguous():
        return x
    if not is_cl(x):
        return x.contiguous()
    B, C, H, W = x.shape
    y = _empty(x, B, C, H, W)
    _call('pgv_transpose_inner', _f(x), _f(y), B, H * W, C, 0, _s(x), nbytes=8 * x.numel())
    return y


def slice_channels(x, lo, hi):
    """x[:, lo:hi] as a dense tensor in x's own memory layout."""
    part = x[:, lo:hi]
    return part.contiguous(memory_format=torch.channels_last) if is_cl(x) else part.contiguous()


def cat_channels(parts):
    """Concatenation along the channel dimension, keeping a channels-last layout when the parts have one."""
    out = torch.cat(parts, dim=1)
    return to_cl(out) if (is_cl(parts[0]) and not is_cl(out)) else out


def same_layout(t, like):
    """`t` in the memory layout of `like` (both logically [B, C, H, W])."""
    return to_cl(t) if is_cl(like) else to_nchw(t)


profile = None        # None, or a dict filled by _call: name -> [calls, flops, bytes, [(start_event, end_event), ...]]


def start_profile():
    """Per-entry-point accounting for bench.py / tools: CUDA events around every C call (eager mode only)."""
    global profile
    profile = {}


def stop_profile():
    """Returns {name: dict(calls, ms, flops, bytes)} and disables profiling.  Synchronises the device."""
    global profile
    torch.cuda.synchronize()
    out = {}
    for name, (calls, flops, nbytes, events) in (profile or {}).items():
        out[name] = dict(calls=calls, ms=sum(a.elapsed_time(b) for a, b in events), flops=flops, bytes=nbytes)
    profile = None
    return out


def _call(name, *args, n=1, flops=0, nbytes=0):
    global launches
    if _mega is not None:
        if name in _MEGA_BUILDERS and len(_mega) < _mega_max_ops():
            _mega.append(_MEGA_BUILDERS[name](*args))
            return
        mega_flush()                  # an op the program kernel does not know: run what is pending first, then this one on its own
    launches += n
    if profile is None:
        _lib.check(getattr(_lib.lib(), name)(*args), name)
        return
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _lib.check(getattr(_lib.lib(), name)(*args), name)
    b.record()
    rec = profile.setdefault(name, [0, 0, 0, []])
    rec[0] += 1
    rec[1] += flops
    rec[2] += nbytes
    rec[3].append((a, b))


def _s(t):
    return _lib.stream_ptr(t.device)


# ------------------------------------------------------------------------------------------------ flow program (one launch per chain)
# A RealNVP flow is a chain of ~50 (forward) / ~90 (backward) latency-bound kernels; between mega_begin() and mega_end() the calls
# that the program kernel knows are RECORDED instead of launched and then run as one persistent launch with grid barriers between
# dependent ops (csrc/pgv_flow_fused.cu, pgv_flow_program).  Everything else flushes the pending program first, so order is preserved.
use_flow_program = True
_mega = None
_mega_keep = []
_mega_ref = None


class _CsParams(ctypes.Structure):
    _fields_ = [('a', ctypes.c_void_p), ('lda', ctypes.c_int), ('b', ctypes.c_void_p), ('ldb', ctypes.c_int), ('M', ctypes.c_int),
                ('N', ctypes.c_int), ('Kd', ctypes.c_int), ('bias', ctypes.c_void_p), ('add_pre', ctypes.c_void_p), ('out_pre', ctypes.c_void_p),
                ('out', ctypes.c_void_p), ('gamma', ctypes.c_void_p), ('beta', ctypes.c_void_p), ('mask', ctypes.c_void_p),
                ('save_mean', ctypes.c_void_p), ('save_rstd', ctypes.c_void_p), ('running_mean', ctypes.c_void_p),
                ('running_var', ctypes.c_void_p), ('momentum', ctypes.c_float), ('eps', ctypes.c_float), ('bn_x', ctypes.c_void_p),
                ('mean', ctypes.c_void_p), ('rstd', ctypes.c_void_p), ('add_post', ctypes.c_void_p), ('dgamma', ctypes.c_void_p),
                ('dbeta', ctypes.c_void_p), ('relu', ctypes.c_int), ('a_vec', ctypes.c_int), ('b_vec', ctypes.c_int)]


class _MegaOp(ctypes.Structure):
    _fields_ = [('kind', ctypes.c_int), ('barrier_after', ctypes.c_int), ('cs', _CsParams)]


(_MOP_GATHER, _MOP_CS_FWD, _MOP_CS_BN_FWD, _MOP_CS_DGRAD, _MOP_CS_BN_BWD, _MOP_COUPLING_FWD, _MOP_COUPLING_BWD, _MOP_SCATTER_ADD,
 _MOP_WGRAD) = range(9)


def _v(p):
    return p.value if isinstance(p, ctypes.c_void_p) else p


def _mop(kind, barrier=1, **kw):
    op = _MegaOp()
    op.kind, op.barrier_after = kind, barrier
    for k, val in kw.items():
        setattr(op.cs, k, _v(val))
    return op


# argument order = the C signatures in include/pgv.h
_MEGA_BUILDERS = {
    'pgv_gather_cols': lambda x, idx, out, B, D, n, st: _mop(_MOP_GATHER, a=x, b=idx, out=out, M=B, N=D, Kd=n),
    'pgv_scatter_add_cols': lambda dst, idx, src, B, D, n, st: _mop(_MOP_SCATTER_ADD, out=dst, b=idx, a=src, M=B, N=D, Kd=n),
    'pgv_coupling_fwd': lambda x, prm, ii, ti, y, ld_in, ld_out, B, D, n_id, n_t, inv, st:
        _mop(_MOP_COUPLING_FWD, a=x, b=prm, bias=ii, add_pre=ti, out=y, out_pre=ld_out, gamma=ld_in, M=B, N=D, Kd=n_id, lda=n_t, relu=inv),
    'pgv_coupling_bwd': lambda dy, dld, x, prm, ii, ti, dx, dprm, B, D, n_id, n_t, st:
        _mop(_MOP_COUPLING_BWD, a=dy, b=dld, bias=x, add_pre=prm, gamma=ii, beta=ti, out=dx, out_pre=dprm, M=B, N=D, Kd=n_id, lda=n_t),
    'pgv_linear_cs_fwd': lambda x, w, bias, res, y, M, N, K, relu, st:
        _mop(_MOP_CS_FWD, a=x, lda=K, b=w, ldb=K, M=M, N=N, Kd=K, bias=bias, add_pre=res, out=y, relu=relu),
    'pgv_linear_cs_dgrad': lambda dy, w, dx, M, N, K, st: _mop(_MOP_CS_DGRAD, a=dy, lda=N, b=w, ldb=K, M=M, N=K, Kd=N, out=dx),
    'pgv_linear_bn_fwd': lambda x, w, bias, res, y_pre, out, g, be, mask, sm, sr, rm, rv, mom, eps, M, N, K, st:
        _mop(_MOP_CS_BN_FWD, a=x, lda=K, b=w, ldb=K, M=M, N=N, Kd=K, bias=bias, add_pre=res, out_pre=y_pre, out=out, gamma=g, beta=be, mask=mask,
             save_mean=sm, save_rstd=sr, running_mean=rm, running_var=rv, momentum=mom, eps=eps),
    'pgv_linear_dgrad_bn_bwd': lambda dy, w, bn_x, g, be, mean, rstd, mask, add_post, dx, dg, db, M, N, K, st:
        _mop(_MOP_CS_BN_BWD, a=dy, lda=N, b=w, ldb=K, M=M, N=K, Kd=N, out=dx, gamma=g, beta=be, mask=mask, bn_x=bn_x, mean=mean, rstd=rstd,
             add_post=add_post, dgamma=dg, dbeta=db),
    # weight gradients have no consumer inside the chain: no barrier behind them, they run beside the data-gradient op that follows
    'pgv_linear_wgrad_f32': lambda h, dy, x, dw, db, M, N, K, st: _mop(_MOP_WGRAD, barrier=0, a=dy, b=x, out=dw, out_pre=db, M=M, N=N, Kd=K),
}
_mega_limits = {}


def _mega_max_ops():
    if 'max' not in _mega_limits:
        L = _lib.lib()
        assert L.pgv_flow_program_op_bytes() == ctypes.sizeof(_MegaOp), "MegaOp layout differs between pgv_flow_fused.cu and ops.py"
        _mega_limits['max'] = L.pgv_flow_program_max_ops()
    return _mega_limits['max']


def mega_begin(ref, rows):
    """Start recording (no-op when disabled or the batch does not fit the column-slice kernels)."""
    global _mega, _mega_ref
    if use_flow_program and _mega is None and colslice_ok(rows):
        _mega_max_ops()
        _mega, _mega_ref = [], (ref, rows)


def mega_flush():
    """Launch what has been recorded so far (recording continues)."""
    global _mega, launches
    if not _mega:
        del _mega_keep[:]
        return
    ops_, (ref, rows) = _mega, _mega_ref
    _mega = None                                  # (the launch itself goes through _call)
    try:
        arr = (_MegaOp * len(ops_))(*ops_)
        counter = _mega_counter(ref)
        _call('pgv_flow_program', _h(ref), ctypes.cast(arr, ctypes.c_void_p), len(ops_), rows, _f(counter), _s(ref), n=2)
    finally:
        _mega = []
        del _mega_keep[:]


def mega_end():
    global _mega
    if _mega is not None:
        mega_flush()
        _mega = None
        del _mega_keep[:]


_mega_counters = {}


def _mega_counter(ref):
    key = (ref.device.index, torch.cuda.current_stream(ref.device).cuda_stream)
    if key not in _mega_counters:
        _mega_counters[key] = torch.zeros(_mega_max_ops() + 1, dtype=torch.int32, device=ref.device)   # grid barrier + one tile counter per wgrad op
    return _mega_counters[key]


def _h(t):
    return _lib.handle(t.device)


def _empty(ref, *shape):
    return torch.empty(*shape, dtype=torch.float32, device=ref.device)


_scratch = {}


def _ws(ref, nbytes):
    """Small reusable fp64 scratch per device (reduction partials)."""
    key = (ref.device.index, torch.cuda.current_stream(ref.device).cuda_stream)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8, device=ref.device)
        _scratch[key] = buf
    return buf


# ------------------------------------------------------------------------------------------------ convolutions
# Deterministic split-K (pgv.h, pgv_conv_cl_*): one zero-headed workspace per (device, stream).  `deterministic = False` passes no
# workspace, which selects the fp32-atomic accumulation of round 1 (kept for A/B measurements).
deterministic = True
CL_WS_BYTES = 48 << 20
_cl_ws = {}


def _clws(ref):
    """(pointer, size) of this stream's channels-last workspace, or (NULL, 0)."""
    if not deterministic:
        return ctypes.c_void_p(0), 0
    key = (ref.device.index, torch.cuda.current_stream(ref.device).cuda_stream)
    buf = _cl_ws.get(key)
    if buf is None:
        buf = torch.zeros(CL_WS_BYTES, dtype=torch.uint8, device=ref.device)
        _cl_ws[key] = buf
    return ctypes.c_void_p(buf.data_ptr()), buf.numel()


# ------------------------------------------------------------------------------------------------ weight gradients off the critical path
# In a backward pass the weight gradient of a layer is a leaf: nothing but the optimizer waits for it, while the data gradient feeds the
# next layer.  `forked()` enqueues the weight-gradient kernels on a child stream of the current stream (an event edge; inside a captured
# step it becomes a graph dependency) so that they fill the SMs the HBM-bound BatchNorm / latency-bound conditioner kernels of the main
# chain leave idle; `join_forks()` at the end of the module's backward makes the parent stream wait for them.
use_wgrad_fork = True
_fork_streams = {}
_fork_keep = {}


class forked:
    def __init__(self, ref, *keep):
        self.ref, self.keep, self.cm = ref, keep, None

    def __enter__(self):
        if not use_wgrad_fork or _mega is not None:      # (recording a flow program: the op goes into the program, no stream switch)
            return self
        dev = self.ref.device
        cur = torch.cuda.current_stream(dev)
        key = (dev.index, cur.cuda_stream)
        st = _fork_streams.get(key)
        if st is None:
            st = _fork_streams[key] = torch.cuda.Stream(device=dev)
        st.wait_stream(cur)
        # operands produced on the parent stream stay referenced until the join: the allocator must not hand their memory out again
        # while the child stream still reads them
        _fork_keep.setdefault(key, []).extend(self.keep)
        self.cm = torch.cuda.stream(st)
        self.cm.__enter__()
        return self

    def __exit__(self, *exc):
        if self.cm is not None:
            self.cm.__exit__(*exc)
        return False


def join_forks(ref):
    dev = ref.device
    cur = torch.cuda.current_stream(dev)
    key = (dev.index, cur.cuda_stream)
    if key in _fork_keep:
        cur.wait_stream(_fork_streams[key])
        del _fork_keep[key]


def _thin_ws(ref):
    """Partial-sum workspace of the thin-layer weight gradient: the partial-tile area of the channels-last workspace (past its header)."""
    ptr, n = _clws(ref)
    if n == 0:
        return ptr, 0
    hdr = _lib.lib().pgv_conv_cl_workspace_bytes()
    return ctypes.c_void_p(ptr.value + hdr), n - hdr


def grad_out_of(param):
    """View into a flat gradient buffer registered on the parameter (TrainStep: `param._pgv_grad_out`): weight-gradient kernels that can
    write their result in the parameter's own layout store it there directly and the gradient needs no packing pass."""
    return getattr(param, '_pgv_grad_out', None)


def conv_out_size(h, k, stride, pad):
    return (h + 2 * pad - k) // stride + 1


def conv_route(Cin, Cout, kh, kw, stride, pad, H, W, Ho, Wo):
    """Which kernel family runs a convolution of this geometry: 'thin' (direct 5x5, one channel), 'cl' (channels-last
    tcgen05 + cp.async), 'tc' (NCHW tcgen05, register-staged gathers) or 'f32' (exact CUDA-core)."""
    if use_thin and _thin(Cin, Cout, kh, kw, stride, pad, H, W, Ho, Wo):
        return 'thin'
    if cl_mode() and _lib.lib().pgv_conv_cl_supported(Cin, Cout, kh, kw, stride, pad):
        return 'cl'
    return 'tc' if _precision == 'tf32' else 'f32'


def prepared_of(param):
    """Operand copies made ahead of time for this parameter (`param._pgv_prepared`, set by TrainStep.prepare_operands and valid for
    the current step only), or None."""
    ready = getattr(param, '_pgv_prepared', None)
    if ready is not None:
        ev = getattr(param, '_pgv_prepared_event', None)      # recorded on the stream that made the copies: wait for it lazily, per
        if ev is not None:                                     # parameter, so that the copies of later layers run under earlier layers
            torch.cuda.current_stream(param.device).wait_event(ev)
    return ready


def prep_conv_weights(w, stride, pad, fwd=True, dgrad=True, out=None):
    """TF32-rounded operand matrices of the channels-last kernels for weight w [Cout, Cin, kh, kw]: (wf, wq).
    out: (wf, wq) buffers to refresh in place (persistent operand copies)."""
    ready = prepared_of(w)
    if ready is not None and out is None:
        return ready
    Cout, Cin, kh, kw = w.shape
    K = Cin * kh * kw
    if out is not None:
        wf, wq = out
    else:
        wf = _empty(w, Cout, K) if fwd else None
        wq = (_empty(w, 4 * Cin, 4 * Cout) if kh == 4 else _empty(w, Cin, Cout)) if dgrad else None
    _call('pgv_conv_cl_prep_weights', _f(w), _f(wf), _f(wq), Cout, Cin, kh, kw, stride, pad, _s(w), nbytes=12 * w.numel())
    return wf, wq


# Data-gradient kernels can also accumulate the BatchNorm-backward sums of the block in front (layer.chain_bwd): correct and tested, but
# OFF: the extra activation read in the epilogue of the 16/32-channel data-gradient launches (already bound by their epilogue and by
# 32-byte requests) costs more than the six reduction launches it removes - captured step 6.52 -> 6.64 ms at B = 160.
fuse_bn_bwd = False
BN_SUMS_MAX_N = 128     # the conv epilogue accumulates statistics in registers, which needs a single N tile (CL_MAX_N)
use_bn_sums = True


def _bn_sums(ref, C, gemm_n, route):
    """fp64 [C][2] buffer for the (sum, sum of squares) by-product of a channels-last convolution, or None if the launch has none."""
    if route == 'cl' and use_bn_sums and gemm_n <= BN_SUMS_MAX_N:
        return torch.empty(2 * C, dtype=torch.float64, device=ref.device)
    return None


def conv2d_fwd(x, w, bias, stride, pad, slope=-1.0, out_hw=None, wf=None, round_out=False, bn_sums=False, bn_bwd_x=None):
    """round_out: round the result to TF32 (set when it feeds a channels-last tensor-core kernel directly).
    bn_sums: return (y, sums) where sums holds the batch statistics of y for ops.bn2d_train_fwd (None if the route has no such by-product).
    bn_bwd_x (backward pass; implies bn_sums): the activation that went into the BatchNorm2d whose output gradient this call computes;
    sums then holds (sum y, sum y * bn_bwd_x), the raw sums of that BatchNorm's backward (ops.bn2d_train_bwd, raw_sums)."""
    bn_sums = bn_sums or bn_bwd_x is not None
    B, Cin, H, W = x.shape
    Cout, _, kh, kw = w.shape
    Ho, Wo = out_hw if out_hw is not None else (conv_out_size(H, kh, stride, pad), conv_out_size(W, kw, stride, pad))
    route = conv_route(Cin, Cout, kh, kw, stride, pad, H, W, Ho, Wo)
    flops = 2 * B * Ho * Wo * Cout * Cin * kh * kw
    if route == 'thin':
        cl = cl_mode()
        y = _empty_cl(x, B, Cout, Ho, Wo) if cl else _empty(x, B, Cout, Ho, Wo)
        _call('pgv_conv5x5s2_c1_fwd', _f(to_nchw(x)), _f(w), _f(bias), _f(y), B, Cout, H, W, Ho, Wo, slope, int(cl), int(cl and round_out),
              _s(x), flops=flops, nbytes=4 * (x.numel() + y.numel()))
        return (y, None) if bn_sums else y
    if route == 'cl':
        x = to_cl(x, round_out=True)
        if wf is None:
            wf, _ = prep_conv_weights(w, stride, pad, dgrad=False)
        y = _empty_cl(x, B, Cout, Ho, Wo)
        sums = _bn_sums(x, Cout, Cout, route) if bn_sums else None
        if sums is not None and bn_bwd_x is not None:
            assert is_cl(bn_bwd_x) and tuple(bn_bwd_x.shape) == tuple(y.shape)
        _call('pgv_conv_cl_fwd_bn', _h(x), _f(x), _f(wf), _f(bias), _f(y), B, H, W, Cin, Cout, kh, kw, stride, pad, Ho, Wo, slope, int(round_out),
              _f(sums), _f(bn_bwd_x if sums is not None else None), *_clws(x), _s(x), n=2 if sums is not None else 1, flops=flops, nbytes=4 * (x.numel() + y.numel() + w.numel()))
        return (y, sums) if bn_sums else y
    x = to_nchw(x)
    y = _empty(x, B, Cout, Ho, Wo)
    _call('pgv_conv2d_fwd_tf32' if route == 'tc' else 'pgv_conv2d_fwd_f32', *((_h(x),) if route == 'tc' else ()), _f(x), _f(w), _f(bias), _f(y),
          B, Cin, H, W, Cout, kh, kw, stride, pad, Ho, Wo, slope, _s(x), flops=flops, nbytes=4 * (x.numel() + y.numel() + w.numel()))
    return (y, None) if bn_sums else y


def _thin(Cin, Cout, kh, kw, stride, pad, H, W, Ho, Wo):
    return bool(_lib.lib().pgv_conv5x5s2_c1_supported(Cin, Cout, kh, kw, stride, pad, H, W, Ho, Wo))


def conv2d_dgrad(dy, w, in_hw, stride, pad, bias=None, slope=-1.0, clamp=None, wq=None, round_out=False, bn_sums=False, bn_bwd_x=None):
    """dx of the convolution with weight w [Cout, Cin, kh, kw]; also the forward of ConvTranspose2d(weight=w).
    clamp=(lo, hi) fuses a Hardtanh (only available on the thin-layer kernel; callers check `tconv_clamp_fusable`).
    bn_sums / bn_bwd_x: as in conv2d_fwd, returns (dx, sums)."""
    bn_sums = bn_sums or bn_bwd_x is not None
    B, Cout, Ho, Wo = dy.shape
    _, Cin, kh, kw = w.shape
    H, W = in_hw
    route = conv_route(Cin, Cout, kh, kw, stride, pad, H, W, Ho, Wo)
    flops = 2 * B * Ho * Wo * Cout * Cin * kh * kw
    if route == 'thin' and slope < 0:
        lo, hi = clamp if clamp is not None else (-float('inf'), float('inf'))
        dx = _empty(dy, B, Cin, H, W)
        _call('pgv_conv5x5s2_c1_dgrad', _f(dy), _f(w), _f(bias), _f(dx), B, Cout, H, W, Ho, Wo, lo, hi, int(is_cl(dy)), _s(dy),
              flops=flops, nbytes=4 * (dx.numel() + dy.numel()))
        return (dx, None) if bn_sums else dx
    assert clamp is None, "fused clamp needs the thin-layer kernel"
    if route == 'cl':
        dy = to_cl(dy, round_out=True)
        if wq is None:
            _, wq = prep_conv_weights(w, stride, pad, fwd=False)
        dx = _empty_cl(dy, B, Cin, H, W)
        sums = _bn_sums(dy, Cin, (4 if kh == 4 else 1) * Cin, route) if bn_sums else None
        if sums is not None and bn_bwd_x is not None:
            assert is_cl(bn_bwd_x) and tuple(bn_bwd_x.shape) == tuple(dx.shape)
        _call('pgv_conv_cl_dgrad_bn', _h(dy), _f(dy), _f(wq), _f(bias), _f(dx), B, H, W, Cin, Cout, kh, kw, stride, pad, Ho, Wo, slope,
              int(round_out), _f(sums), _f(bn_bwd_x if sums is not None else None), *_clws(dy), _s(dy), n=2 if sums is not None else 1, flops=flops,
              nbytes=4 * (dx.numel() + dy.numel() + w.numel()))
        return (dx, sums) if bn_sums else dx
    dy = to_nchw(dy)
    dx = _empty(dy, B, Cin, H, W)
    if _precision == 'tf32' and stride <= 2:
        _call('pgv_conv2d_dgrad_tf32', _h(dy), _f(dy), _f(w), _f(bias), _f(dx), B, Cin, H, W, Cout, kh, kw, stride, pad, Ho, Wo, slope,
              _s(dy), flops=flops, nbytes=4 * (dx.numel() + dy.numel() + w.numel()))
        return (dx, None) if bn_sums else dx
    _call('pgv_conv2d_dgrad_f32', _f(dy), _f(w), _f(bias), _f(dx), B, Cin, H, W, Cout, kh, kw, stride, pad, Ho, Wo, slope, _s(dy),
          flops=flops, nbytes=4 * (dx.numel() + dy.numel() + w.numel()))
    return (dx, None) if bn_sums else dx


def conv2d_wgrad(x, dy, w_shape, stride, pad, want_bias, db=None, out=None):
    """db: bias gradient if the caller already has it (BatchNorm backward by-product); else computed here when want_bias.
    out: preallocated destination of dw in the weight's own layout (e.g. a view into the flat gradient buffer)."""
    B, Cin, H, W = x.shape
    _, Cout, Ho, Wo = dy.shape
    _, _, kh, kw = w_shape
    route = conv_route(Cin, Cout, kh, kw, stride, pad, H, W, Ho, Wo)
    flops = 2 * B * Ho * Wo * Cout * Cin * kh * kw
    dw = out if out is not None else _empty(x, *w_shape)
    assert tuple(dw.shape) == tuple(w_shape) and dw.is_contiguous()
    if route == 'thin':
        _call('pgv_conv5x5s2_c1_wgrad', _f(to_nchw(x)), _f(dy), _f(dw), B, Cout, H, W, Ho, Wo, int(is_cl(dy)), *_thin_ws(x), _s(x), n=2,
              flops=flops, nbytes=4 * (x.numel() + dy.numel()))
        return dw, (db if db is not None else (channel_sum(dy) if want_bias else None))
    if route == 'cl':
        x, dy = to_cl(x, round_out=True), to_cl(dy, round_out=True)
        # the kernel (or its split-K finish pass) writes the PyTorch layout [Cout, Cin, kh, kw] directly
        _call('pgv_conv_cl_wgrad', _h(x), _f(x), _f(dy), _f(dw), 1, B, H, W, Cin, Cout, kh, kw, stride, pad, Ho, Wo, *_clws(x), _s(x), n=2,
              flops=flops, nbytes=4 * (x.numel() + dy.numel() + dw.numel()))
        return dw, (db if db is not None else (channel_sum(dy) if want_bias else None))
    x, dy = to_nchw(x), to_nchw(dy)
    if route == 'tc':
        _call('pgv_conv2d_wgrad_tf32', _h(x), _f(x), _f(dy), _f(dw), B, Cin, H, W, Cout, kh, kw, stride, pad, Ho, Wo, _s(x), n=2,
              flops=flops, nbytes=4 * (x.numel() + dy.numel() + dw.numel()))
        return dw, (db if db is not None else (channel_sum(dy) if want_bias else None))
    db = _empty(x, Cout) if want_bias else None
    _call('pgv_conv2d_wgrad_f32', _f(x), _f(dy), _f(dw), _f(db), B, Cin, H, W, Cout, kh, kw, stride, pad, Ho, Wo, _s(x),
          n=4 if want_bias else 2, flops=flops, nbytes=4 * (x.numel() + dy.numel() + dw.numel()))
    return dw, db


def channel_sum(x):
    B, C, H, W = x.shape
    out = _empty(x, C)
    if is_cl(x):
        _call('pgv_colsum_cl', _f(x), _f(out), B * H * W, C, _f(_ws(x, 16 * C)), _s(x), n=3, nbytes=4 * x.numel())
    else:
        _call('pgv_channel_sum', _f(x), _f(out), B, C, H * W, _s(x), n=2, nbytes=4 * x.numel())
    return out


# ------------------------------------------------------------------------------------------------ normalisation
def bn2d_train_fwd(x, bn, sums=None):
    """sums: batch statistics of x already accumulated by the convolution that produced it (conv2d_fwd / conv2d_dgrad, bn_sums=True)."""
    B, C = x.shape[:2]
    HW = x[0, 0].numel()
    y, mean, rstd = torch.empty_like(x), _empty(x, C), _empty(x, C)
    if sums is not None:
        assert is_cl(x) and sums.numel() == 2 * C
        _call('pgv_bn_cl_train_apply', _f(x), _f(sums), _f(bn.weight), _f(bn.bias), _f(y), _f(mean), _f(rstd), _f(bn.running_mean),
              _f(bn.running_var), bn.momentum, bn.eps, B * HW, C, 1, _s(x), nbytes=4 * 2 * x.numel())
        return y, mean, rstd
    if is_cl(x):        # result rounded to TF32: its consumers are the cp.async-fed tensor-core kernels
        _call('pgv_bn_cl_train_fwd', _f(x), _f(bn.weight), _f(bn.bias), _f(y), _f(mean), _f(rstd), _f(bn.running_mean), _f(bn.running_var),
              bn.momentum, bn.eps, B * HW, C, 1, _f(_ws(x, 16 * C)), _s(x), n=3, nbytes=4 * 3 * x.numel())
        return y, mean, rstd
    _call('pgv_bn2d_train_fwd', _f(x), _f(bn.weight), _f(bn.bias), _f(y), _f(mean), _f(rstd), _f(bn.running_mean),
          _f(bn.running_var), bn.momentum, bn.eps, B, C, HW, _f(_ws(x, 16 * C)), _s(x), n=3, nbytes=4 * 3 * x.numel())
    return y, mean, rstd


def bn2d_eval_fwd(x, bn):
    B, C = x.shape[:2]
    y = torch.empty_like(x)
    if is_cl(x):
        _call('pgv_bn_cl_eval_fwd', _f(x), _f(bn.weight), _f(bn.bias), _f(bn.running_mean), _f(bn.running_var), _f(y), bn.eps,
              B * x[0, 0].numel(), C, 1, _s(x), nbytes=8 * x.numel())
        return y
    _call('pgv_bn2d_eval_fwd', _f(x), _f(bn.weight), _f(bn.bias), _f(bn.running_mean), _f(bn.running_var), _f(y), bn.eps, B, C,
          x[0, 0].numel(), _s(x))
    return y


def bn2d_train_bwd(dy, x, gamma, mean, rstd, slope, want_colsum=False, raw_sums=None):
    """(dx, dgamma, dbeta[, colsum(dx)]): colsum(dx) is the bias gradient of the convolution in front of the block; the
    channels-last kernel produces it while writing dx.  raw_sums: (sum dy, sum dy * x) per channel if the convolution that produced dy
    accumulated them (conv2d_fwd / conv2d_dgrad, bn_bwd_x=x): the reduction pass over dy and x is skipped."""
    B, C = x.shape[:2]
    dy = same_layout(dy, x)
    dx, dg, db = torch.empty_like(x), _empty(x, C), _empty(x, C)
    if is_cl(x):
        cs = _empty(x, C) if want_colsum else None
        assert raw_sums is None or raw_sums.numel() == 2 * C
        _call('pgv_bn_cl_train_bwd', _f(dy), _f(x), _f(gamma), _f(mean), _f(rstd), _f(dx), _f(dg), _f(db), _f(cs), slope, B * x[0, 0].numel(), C,
              1, _f(raw_sums), _f(_ws(x, 24 * C)), _s(x), n=4 if want_colsum else 3, nbytes=4 * (5 if raw_sums is None else 3) * x.numel())
        return (dx, dg, db, cs) if want_colsum else (dx, dg, db)
    _call('pgv_bn2d_train_bwd', _f(dy), _f(x), _f(gamma), _f(mean), _f(rstd), _f(dx), _f(dg), _f(db), slope, B, C, x[0, 0].numel(),
          _f(_ws(x, 16 * C)), _s(x), n=3, nbytes=4 * 5 * x.numel())
    return (dx, dg, db, channel_sum(dx)) if want_colsum else (dx, dg, db)


def lrelu_bwd(dy, a, slope=LRELU_SLOPE):
    dx = torch.empty_like(a)
    if a.dim() == 4:
        dy = same_layout(dy, a)
    if is_cl(a):
        _call('pgv_lrelu_bwd_round', _f(dy), _f(a), _f(dx), slope, a.numel(), 1, _s(a), nbytes=12 * a.numel())
        return dx
    _call('pgv_lrelu_bwd', _f(dy), _f(a), _f(dx), slope, a.numel(), _s(a))
    return dx


def bn1d_train_fwd(x, bn, relu=False, mask=None):
    B, F = x.shape
    y, mean, rstd = torch.empty_like(x), _empty(x, F), _empty(x, F)
    _call('pgv_bn1d_train_fwd', _f(x), _f(bn.weight), _f(bn.bias), _f(mask), _f(y), _f(mean), _f(rstd), _f(bn.running_mean),
          _f(bn.running_var), bn.momentum, bn.eps, int(relu), B, F, _s(x))
    return y, mean, rstd


def bn1d_eval_fwd(x, bn, relu=False):
    B, F = x.shape
    y = torch.empty_like(x)
    _call('pgv_bn1d_eval_fwd', _f(x), _f(bn.weight), _f(bn.bias), _f(bn.running_mean), _f(bn.running_var), _f(y), bn.eps, int(relu), B, F,
          _s(x))
    return y


def bn1d_train_bwd(dy, x, bn, mean, rstd, relu=False, mask=None):
    B, F = x.shape
    dx, dg, db = torch.empty_like(x), _empty(x, F), _empty(x, F)
    _call('pgv_bn1d_train_bwd', _f(dy), _f(x), _f(bn.weight), _f(bn.bias), _f(mean), _f(rstd), _f(mask), _f(dx), _f(dg), _f(db), int(relu),
          B, F, _s(x))
    return dx, dg, db


def flowbn_train_fwd(x, t):
    B, F = x.shape
    y, mean, var, ld = torch.empty_like(x), _empty(x, F), _empty(x, F), _empty(x, 1)
    _call('pgv_flowbn_train_fwd', _f(x), _f(t.unconstrained_weight), _f(t.bias), _f(y), _f(mean), _f(var), _f(t.running_mean),
          _f(t.running_var), _f(ld), t.momentum, t.eps, B, F, _s(x), n=2)
    return y, mean, var, ld


def flowbn_eval(x, t, inverse=False):
    B, F = x.shape
    y, ld = torch.empty_like(x), _empty(x, 1)
    _call('pgv_flowbn_eval', _f(x), _f(t.unconstrained_weight), _f(t.bias), _f(t.running_mean), _f(t.running_var), _f(y), _f(ld), t.eps,
          int(inverse), B, F, _s(x))
    return y, ld


def flowbn_train_bwd(dy, x, t, mean, var, g_ld_sum):
    B, F = x.shape
    dx, du, db = torch.empty_like(x), _empty(x, F), _empty(x, F)
    _call('pgv_flowbn_train_bwd', _f(dy), _f(x), _f(t.unconstrained_weight), _f(mean), _f(var), _f(g_ld_sum), _f(dx), _f(du), _f(db), t.eps,
          B, F, _s(x))
    return dx, du, db


# ------------------------------------------------------------------------------------------------ dense layers
# Below this many multiply-accumulates a dense layer is latency-bound (a 160x300x300 flow conditioner layer is 14 M): the
# persistent tensor-core kernel's fixed cost (TMEM allocation, barrier set-up, pipeline fill) exceeds the whole job, so
# such layers run on the exact-fp32 CUDA-core GEMM instead.
SMALL_GEMM_MACS = 48_000_000
SMALL_WEIGHT_ELEMS = 1 << 20      # the flow conditioner layers (<= 610 x 300 weights) never go to the tensor cores, whatever the batch


def _use_tc(M, N, K):
    return _precision == 'tf32' and M * N * K >= SMALL_GEMM_MACS and N * K >= SMALL_WEIGHT_ELEMS


use_colslice = True    # small Linear layers (flow conditioners): column-slice GEMM, BatchNorm1d fused where one follows
CS_MAX_ROWS = 256


def colslice_ok(M):
    return use_colslice and M <= CS_MAX_ROWS


def linear_bn_fwd(x, w, bias, bn, residual=None, mask=None, keep_pre=True):
    """(y_pre, t, mean, rstd): y_pre = x @ w.T + bias + residual; t = mask * relu(BatchNorm1d(y_pre)) in training mode."""
    M, K = x.shape
    N = w.shape[0]
    y_pre = _empty(x, M, N) if keep_pre else None
    t, mean, rstd = _empty(x, M, N), _empty(x, N), _empty(x, N)
    _call('pgv_linear_bn_fwd', _f(x), _f(w), _f(bias), _f(residual), _f(y_pre), _f(t), _f(bn.weight), _f(bn.bias), _f(mask), _f(mean),
          _f(rstd), _f(bn.running_mean), _f(bn.running_var), bn.momentum, bn.eps, M, N, K, _s(x),
          flops=2 * M * N * K, nbytes=4 * (M * K + N * K + 2 * M * N))
    return y_pre, t, mean, rstd


def linear_dgrad_bn_bwd(dy, w, bn_x, bn, mean, rstd, mask=None, add_post=None):
    """dt = dy @ w, pushed back through mask * relu(BatchNorm1d(bn_x)); returns (dx [+ add_post], dgamma, dbeta)."""
    M, N = dy.shape
    K = w.shape[1]
    dx, dg, db = _empty(dy, M, K), _empty(dy, K), _empty(dy, K)
    _call('pgv_linear_dgrad_bn_bwd', _f(dy), _f(w), _f(bn_x), _f(bn.weight), _f(bn.bias), _f(mean), _f(rstd), _f(mask), _f(add_post), _f(dx),
          _f(dg), _f(db), M, N, K, _s(dy), flops=2 * M * N * K, nbytes=4 * (M * N + N * K + 3 * M * K))
    return dx, dg, db


def linear_fwd(x, w, bias, relu=False, residual=None):
    """y = act(x @ w.T + bias + residual); x [M,K], w [N,K] (nn.Linear layout)."""
    M, K = x.shape
    N = w.shape[0]
    y = _empty(x, M, N)
    acct = dict(flops=2 * M * N * K, nbytes=4 * (M * K + N * K + M * N))
    if not _use_tc(M, N, K) and use_colslice:
        _call('pgv_linear_cs_fwd', _f(x), _f(w), _f(bias), _f(residual), _f(y), M, N, K, int(relu), _s(x), **acct)
        return y
    if _use_tc(M, N, K):
        _call('pgv_linear_fwd_tf32', _h(x), _f(x), _f(w), _f(bias), _f(residual), _f(y), M, N, K, int(relu), _s(x), **acct)
    else:
        _call('pgv_gemm_f32', _h(x), 0, 1, _f(x), K, _f(w), K, _f(y), N, M, N, K, _f(bias), int(relu), _f(residual), N, _s(x), **acct)
    return y


def linear_dgrad(dy, w):
    """dx = dy @ w; dy [M,N], w [N,K]."""
    M, N = dy.shape
    K = w.shape[1]
    dx = _empty(dy, M, K)
    acct = dict(flops=2 * M * N * K, nbytes=4 * (M * K + N * K + M * N))
    if not _use_tc(M, N, K) and use_colslice:
        _call('pgv_linear_cs_dgrad', _f(dy), _f(w), _f(dx), M, N, K, _s(dy), **acct)
        return dx
    if _use_tc(M, N, K):
        _call('pgv_linear_dgrad_tf32', _h(dy), _f(dy), _f(w), _f(dx), M, N, K, _s(dy), **acct)
    else:
        _call('pgv_gemm_f32', _h(dy), 0, 0, _f(dy), N, _f(w), K, _f(dx), K, M, K, N, None, 0, None, 0, _s(dy), **acct)
    return dx


def linear_wgrad(dy, x, want_bias=True, out=None):
    """dw = dy.T @ x [N,K]; db = column sums of dy.  `out`: preallocated [N, K] destination (e.g. a slice of the flat gradient buffer)."""
    M, N = dy.shape
    K = x.shape[1]
    dw = out if out is not None else _empty(dy, N, K)
    assert dw.shape == (N, K) and dw.is_contiguous()
    acct = dict(flops=2 * M * N * K, nbytes=4 * (M * K + N * K + M * N))
    if not _use_tc(M, N, K):
        db = _empty(dy, N) if want_bias else None
        _call('pgv_linear_wgrad_f32', _h(dy), _f(dy), _f(x), _f(dw), _f(db), M, N, K, _s(dy), **acct)
        return dw, db
    _call('pgv_linear_wgrad_tf32', _h(dy), _f(dy), _f(x), _f(dw), M, N, K, _s(dy), **acct)
    db = None
    if want_bias:
        db = _empty(dy, N)
        _call('pgv_colsum', _f(dy), _f(db), M, N, _s(dy))
    return dw, db


# ---- the two big fully-connected layers (encoder.py:84, decoder.py:70) on the channels-last tensor-core kernel ----
use_fc_cl = True


def round_copy(x, ld=None, out=None):
    """TF32-rounded copy of the 2-D tensor x with row pitch `ld` (>= columns, zero padded): [rows, ld]."""
    rows, cols = x.shape
    ld = cols if ld is None else ld
    y = out if out is not None else _empty(x, rows, ld)
    assert y.shape == (rows, ld) and y.is_contiguous()
    _call('pgv_round_copy', _f(x), x.stride(0), _f(y), ld, rows, cols, _s(x), nbytes=4 * (x.numel() + y.numel()))
    return y


def fc_route(M, N, K):
    """'cl': cp.async / TMA fed tcgen05 kernel on rounded, 16-byte-aligned copies of the operands; else the generic Linear path."""
    return 'cl' if (cl_mode() and use_fc_cl and _use_tc(M, N, K) and N % 4 == 0) else 'generic'


def fc_fwd(x, w, bias, training=True):
    """y = x @ w.T + bias for a large Linear.  Returns (y, ctx); ctx carries the rounded operands the backward re-uses."""
    M, K = x.shape
    N = w.shape[0]
    if fc_route(M, N, K) != 'cl':
        return linear_fwd(x, w, bias), (x, None, None, K, None)
    Kp = (K + 3) // 4 * 4
    ready = prepared_of(w)                       # rounded / re-pitched weight copy made ahead of the forward pass (TrainStep)
    xr, wr = round_copy(x, Kp), (ready if ready is not None else round_copy(w, Kp))
    # Data gradient dx = dy @ W.  With M % 4 == 0 it runs on the weight-gradient form of the kernel, whose operands are both
    # "reduction index x contiguous output index": W [N, Kp] as stored (the rounded copy the forward already made) and dy^T [N, M].
    # Otherwise it needs the rounded transpose W^T [K, N], a second pass over the whole weight matrix.
    wt = None
    if training and M % 4 != 0:
        wt = _empty(w, K, N)
        _call('pgv_transpose_inner', _f(w), _f(wt), 1, N, K, 1, _s(w), nbytes=8 * w.numel())
    y = _empty(x, M, N)
    _call('pgv_linear_cl_fwd', _h(x), _f(xr), _f(wr), _f(bias), _f(y), M, N, Kp, *_clws(x), _s(x), n=2,
          flops=2 * M * N * K, nbytes=4 * (M * K + N * K + M * N))
    return y, (xr, wt, Kp, K, wr if training and wt is None else None)


def fc_bwd(dy, ctx, w, need_dx=True, out=None):
    """(dx, dw, db) of fc_fwd; `out`: preallocated [N, K] destination for dw (e.g. a slice of the flat gradient buffer)."""
    xr, wt, Kp, K, wr = ctx
    M, N = dy.shape
    if wt is None and wr is None:
        dw, db = linear_wgrad(dy, xr, out=out)
        return (linear_dgrad(dy, w) if need_dx else None), dw, db
    db = _empty(dy, N)
    _call('pgv_colsum', _f(dy), _f(db), M, N, _s(dy))
    dyr = round_copy(dy)
    dw = out if out is not None else _empty(dy, N, K)
    assert dw.shape == (N, K) and dw.is_contiguous()
    with forked(dy, dyr, xr, dw):                   # 120 MB of output, independent of the data gradient below
        _call('pgv_linear_cl_wgrad', _h(dy), _f(dyr), _f(xr), _f(dw), K, M, N, Kp, K, *_clws(dy), _s(dy), n=2,
              flops=2 * M * N * K, nbytes=4 * (M * K + N * K + M * N))
    if not need_dx:
        join_forks(dy)
        return None, dw, db
    if wr is not None:
        # dx^T [Kp, M] = W[N, Kp]^T dy^T[N, M] (reduction over N), then a small transpose back
        dyt, dxt, dxp = _empty(dy, N, M), _empty(dy, Kp, M), _empty(dy, M, Kp)
        _call('pgv_transpose_inner', _f(dy), _f(dyt), 1, M, N, 1, _s(dy), nbytes=8 * dy.numel())
        _call('pgv_linear_cl_wgrad', _h(dy), _f(wr), _f(dyt), _f(dxt), M, N, Kp, M, M, *_clws(dy), _s(dy), n=2,
              flops=2 * M * N * K, nbytes=4 * (M * K + N * K + M * N))
        _call('pgv_transpose_inner', _f(dxt), _f(dxp), 1, Kp, M, 0, _s(dy), nbytes=8 * dxp.numel())
        join_forks(dy)
        return (dxp if Kp == K else dxp[:, :K].contiguous()), dw, db
    dx = _empty(dy, M, K)
    _call('pgv_linear_cl_dgrad', _h(dy), _f(dyr), _f(wt), _f(dx), M, N, K, *_clws(dy), _s(dy), n=2,
          flops=2 * M * N * K, nbytes=4 * (M * K + N * K + M * N))
    join_forks(dy)
    return dx, dw, db


# ------------------------------------------------------------------------------------------------ latent space / flows
def reparam_fwd(mu_logvar, eps):
    B, _, D = mu_logvar.shape
    z = _empty(mu_logvar, B, D)
    _call('pgv_reparam_fwd', _f(mu_logvar), _f(eps), _f(z), B, D, _s(z))
    return z


def reparam_bwd(dz, mu_logvar, eps, add=None):
    B, _, D = mu_logvar.shape
    d = torch.empty_like(mu_logvar)
    _call('pgv_reparam_bwd', _f(dz), _f(mu_logvar), _f(eps), _f(add), _f(d), B, D, _s(d))
    return d


def gather_cols(x, idx):
    B, D = x.shape
    out = _empty(x, B, idx.numel())
    _call('pgv_gather_cols', _f(x), _f(idx), _f(out), B, D, idx.numel(), _s(x))
    return out


def scatter_add_cols_(dst, idx, src):
    B, D = dst.shape
    _call('pgv_scatter_add_cols', _f(dst), _f(idx), _f(src), B, D, idx.numel(), _s(dst))
    return dst


def coupling_fwd(x, params, id_idx, tr_idx, logdet_in, inverse=False):
    B, D = x.shape
    y, ld = torch.empty_like(x), _empty(x, B)
    _call('pgv_coupling_fwd', _f(x), _f(params), _f(id_idx), _f(tr_idx), _f(y), _f(logdet_in), _f(ld), B, D, id_idx.numel(),
          tr_idx.numel(), int(inverse), _s(x))
    return y, ld


def coupling_bwd(dy, dlogdet, x, params, id_idx, tr_idx):
    B, D = x.shape
    dx, dp = torch.empty_like(x), torch.empty_like(params)
    _call('pgv_coupling_bwd', _f(dy), _f(dlogdet), _f(x), _f(params), _f(id_idx), _f(tr_idx), _f(dx), _f(dp), B, D, id_idx.numel(),
          tr_idx.numel(), _s(x))
    return dx, dp


def hardtanh_fwd(x, lo, hi):
    y = torch.empty_like(x)
    _call('pgv_hardtanh_fwd', _f(x), _f(y), lo, hi, x.numel(), _s(x))
    return y


def hardtanh_bwd(dy, x, lo, hi):
    dx = torch.empty_like(x)
    _call('pgv_hardtanh_bwd', _f(dy), _f(x), _f(dx), lo, hi, x.numel(), _s(x))
    return dx


def mul(x, m):
    y = torch.empty_like(x)
    _call('pgv_mul', _f(x), _f(m), _f(y), x.numel(), _s(x))
    return y


def add(a, b):
    y = torch.empty_like(a)
    _call('pgv_add', _f(a), _f(b), _f(y), a.numel(), _s(a))
    return y


def add_scalar(a, scalar, n, ref):
    y = _empty(ref, n)
    _call('pgv_add_scalar', _f(a), _f(scalar), _f(y), n, _s(ref))
    return y


def colsum(x):
    out = _empty(x, x.shape[1])
    _call('pgv_colsum', _f(x), _f(out), x.shape[0], x.shape[1], _s(x))
    return out


class DeviceTables:
    """int32 device copies of PresetIndexesHelper.device_tables() (cached per device)."""

    def __init__(self, idx_helper):
        self.host = idx_helper.device_tables()
        self.n_num = len(self.host['num_cols'])
        self.n_grp = len(self.host['grp_start'])
        self._dev = {}

    def on(self, device):
        key = (device.type, device.index)
        if key not in self._dev:
            self._dev[key] = {k: torch.from_numpy(np.ascontiguousarray(v)).to(device) for k, v in self.host.items()}
        return self._dev[key]


def preset_act_softmax_fwd(x, tables):
    t = tables.on(x.device)
    y = torch.empty_like(x)
    _call('pgv_preset_act_softmax_fwd', _f(x), _f(y), _f(t['num_cols']), tables.n_num, _f(t['grp_start']), _f(t['grp_len']), tables.n_grp,
          x.shape[0], x.shape[1], _s(x))
    return y


def preset_act_softmax_bwd(dy, x, y, tables):
    t = tables.on(x.device)
    dx = torch.empty_like(x)
    _call('pgv_preset_act_softmax_bwd', _f(dy), _f(x), _f(y), _f(dx), _f(t['num_cols']), tables.n_num, _f(t['grp_start']), _f(t['grp_len']),
          tables.n_grp, x.shape[0], x.shape[1], _s(x))
    return dx


# ------------------------------------------------------------------------------------------------ losses / optimizer
def sqerr_fwd(a, b, scale):
    out = _empty(a, 1)
    _call('pgv_sqerr_fwd', _f(a), _f(b), a.numel(), float(scale), _f(out), _f(_ws(a, 8)), _s(a), n=3)
    return out


def sqerr_bwd(a, b, scale, gout):
    da = torch.empty_like(a)
    _call('pgv_sqerr_bwd', _f(a), _f(b), a.numel(), float(scale), _f(gout), _f(da), _s(a))
    return da


def latent_loss_fwd(ml, z0, zk, logdet, normalize):
    B, _, D = ml.shape
    out = _empty(ml, 1)
    _call('pgv_latent_loss_fwd', _f(ml), _f(z0), _f(zk), _f(logdet), B, D, int(normalize), _f(out), _f(_ws(ml, 8)), _s(ml), n=3)
    return out


def latent_loss_bwd(gout, ml, z0, zk, normalize):
    B, _, D = ml.shape
    dml, dz0, dzk, dld = torch.empty_like(ml), torch.empty_like(z0), torch.empty_like(zk), _empty(ml, B)
    _call('pgv_latent_loss_bwd', _f(gout), _f(ml), _f(z0), _f(zk), B, D, int(normalize), _f(dml), _f(dz0), _f(dzk), _f(dld), _s(ml))
    return dml, dz0, dzk, dld


def dkl_fwd(ml, normalize):
    B, _, D = ml.shape
    out = _empty(ml, 1)
    _call('pgv_dkl_fwd', _f(ml), B, D, int(normalize), _f(out), _f(_ws(ml, 8)), _s(ml), n=3)
    return out


def dkl_bwd(gout, ml, normalize):
    B, _, D = ml.shape
    d = torch.empty_like(ml)
    _call('pgv_dkl_bwd', _f(gout), _f(ml), B, D, int(normalize), _f(d), _s(ml))
    return d


def synth_useful_counts(v_in, tables):
    """fp64 [n_groups]: rows of v_in that count for each categorical group (loss.py:172's normaliser)."""
    t = tables.on(v_in.device)
    counts = torch.empty(tables.n_grp, dtype=torch.float64, device=v_in.device)
    _call('pgv_synth_useful_counts', _f(v_in.contiguous()), v_in.shape[0], v_in.shape[1], _f(t['grp_vol_col']), tables.n_grp, _f(counts), _s(v_in))
    return counts


def synth_loss_fwd(v_out, v_in, tables, normalize, factor, cat_softmax, temperature, group_counts=None):
    t = tables.on(v_out.device)
    B, L = v_out.shape
    out = _empty(v_out, 1)
    ws = torch.empty(_lib.lib().pgv_synth_loss_workspace_bytes(tables.n_grp), dtype=torch.uint8, device=v_out.device)
    _call('pgv_synth_loss_fwd', _f(v_out), _f(v_in), B, L, _f(t['num_cols']), _f(t['num_vol_col']), tables.n_num, _f(t['grp_start']),
          _f(t['grp_len']), _f(t['grp_vol_col']), tables.n_grp, int(normalize), float(factor), int(cat_softmax), float(temperature),
          _f(group_counts), _f(out), _f(ws), _s(v_out), n=3)
    return out, ws


def synth_loss_bwd(gout, v_out, v_in, tables, normalize, factor, cat_softmax, temperature, ws):
    t = tables.on(v_out.device)
    B, L = v_out.shape
    d = torch.empty_like(v_out)
    _call('pgv_synth_loss_bwd', _f(gout), _f(v_out), _f(v_in), B, L, _f(t['num_cols']), _f(t['num_vol_col']), tables.n_num,
          _f(t['grp_start']), _f(t['grp_len']), _f(t['grp_vol_col']), tables.n_grp, int(normalize), float(factor), int(cat_softmax),
          float(temperature), _f(ws), _f(d), _s(v_out))
    return d


# ------------------------------------------------------------------------------------------------ monitoring metrics / inference tail
class MetricTables:
    """Device copies of data.preset.metric_tables() (cached per device)."""

    def __init__(self, idx_helper, limited_vst_params_indexes=None, default_values=None):
        from ..data.preset import metric_tables
        self.host = metric_tables(idx_helper, limited_vst_params_indexes, default_values)
        self.P = len(self.host['kind'])
        self._dev = {}

    def on(self, device):
        key = (device.type, device.index)
        if key not in self._dev:
            self._dev[key] = {k: torch.from_numpy(np.ascontiguousarray(v)).to(device) for k, v in self.host.items()}
        return self._dev[key]


def preset_metrics(v_out, v_in, tables, l1=False, acc_scale=100.0, per_param=False):
    """(out4, acc): out4 = [quantised numerical loss, mean accuracy * acc_scale, #numerical, #categorical] (device tensor);
    acc = per-VST-parameter accuracies (-1 for non-categorical parameters) when per_param."""
    t = tables.on(v_out.device)
    v_out, v_in = v_out.detach().contiguous(), v_in.detach().contiguous()
    B, L = v_out.shape
    out4, partial = _empty(v_out, 4), _empty(v_out, tables.P)
    acc = _empty(v_out, tables.P) if per_param else None
    _call('pgv_preset_metrics', _f(v_out), _f(v_in), B, L, _f(t['kind']), _f(t['col']), _f(t['len']), _f(t['card']), tables.P, int(l1),
          float(acc_scale), _f(partial), _f(out4), _f(acc), _s(v_out), n=2)
    return out4, acc


def learnable_to_full(v, tables):
    t = tables.on(v.device)
    v = v.detach().contiguous()
    B, L = v.shape
    full = _empty(v, B, tables.P)
    _call('pgv_learnable_to_full', _f(v), B, L, _f(t['kind']), _f(t['col']), _f(t['len']), _f(t['card']), _f(t['fill']), tables.P, _f(full), _s(v))
    return full


def flow_params_loss_fwd(ml, z0, ld_t, ld_u, divisor):
    B, _, D = ml.shape
    out = _empty(ml, 1)
    _call('pgv_flow_params_loss_fwd', _f(ml), _f(z0), _f(ld_t), _f(ld_u), B, D, float(divisor), _f(out), _f(_empty(ml, B)), _s(ml), n=2)
    return out


def flow_params_loss_bwd(gout, ml, z0, divisor):
    B, _, D = ml.shape
    dml, dz0, dld = torch.empty_like(ml), torch.empty_like(z0), _empty(ml, B)
    _call('pgv_flow_params_loss_bwd', _f(gout), _f(ml), _f(z0), B, D, float(divisor), _f(dml), _f(dz0), _f(dld), _s(ml))
    return dml, dz0, dld


def nan_flags_(flags, *scalars):
    """flags (int32 [1], device) |= bit i for every NaN scalar i (up to 5 one-element device tensors; train.py:245)."""
    assert len(scalars) <= 5 and flags.dtype == torch.int32
    ptrs = [_f(s.detach()) for s in scalars] + [ctypes.c_void_p(0)] * (5 - len(scalars))
    _call('pgv_nan_flags', *ptrs, _f(flags), _s(flags))
    return flags


def spectrogram_stats(x):
    """x [N, ...]: (per_item [N, 4] = min / max / mean / unbiased variance, dataset [4] = min, max, mean of means, sqrt(mean variance))."""
    x = x.contiguous()
    N = x.shape[0]
    per, ds = _empty(x, N, 4), _empty(x, 4)
    _call('pgv_spectrogram_stats', _f(x), N, x[0].numel(), _f(per), _f(ds), _s(x), n=2)
    return per, ds


def coupling_inv_bwd(dx_out, dlogdet, x_out, params, id_idx, tr_idx):
    B, D = x_out.shape
    dy, dp = torch.empty_like(x_out), torch.empty_like(params)
    _call('pgv_coupling_inv_bwd', _f(dx_out), _f(dlogdet), _f(x_out), _f(params), _f(id_idx), _f(tr_idx), _f(dy), _f(dp), B, D, id_idx.numel(),
          tr_idx.numel(), _s(x_out))
    return dy, dp


def l2_prefetch(t):
    """Hint (no result): pull the contiguous tensor `t` into L2 from a child stream of the current stream."""
    with forked(t, t):
        _call('pgv_l2_prefetch', _f(t), t.numel() * t.element_size(), _s(t))
