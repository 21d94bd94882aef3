"""GPU: the channels-last companions of the tensor-core convolution path (BatchNorm2d on [P, C], per-channel sums,
layout converters, weight re-packing, thin-layer kernels with channels-last operands), each against an fp64 PyTorch
evaluation.  The channels-last convolutions themselves are covered for every layer geometry by
test_kernels_gpu.py::test_tensor_core_convs_against_fp64[use_cl=True]."""
import pytest
import torch
import torch.nn.functional as F

from preset_gen_vae_b200.model import layer, ops

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return torch.randn(*shape, device=DEV, generator=g) * scale


def tf32_rna(x):
    """Round-to-nearest (ties away) to 10 mantissa bits, like cvt.rna.tf32.f32."""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def test_layout_converters_round_trip_and_rounding():
    x = rnd(3, 24, 5, 7, seed=1)
    cl = ops.to_cl(x)
    assert ops.is_cl(cl) and torch.equal(cl, x) and cl.data_ptr() != x.data_ptr()
    assert torch.equal(cl.permute(0, 2, 3, 1).contiguous().view(-1), cl.as_strided((cl.numel(),), (1,)))   # physically NHWC
    back = ops.to_nchw(cl)
    assert back.is_contiguous() and torch.equal(back, x)
    r = ops.to_cl(x, round_out=True)
    assert torch.equal(r, tf32_rna(x))
    assert ops.to_cl(cl) is cl and ops.to_nchw(x) is x


@pytest.mark.parametrize("cout,cin,k", [(16, 8, 4), (64, 32, 4), (2048, 512, 1)])
def test_weight_repacking(cout, cin, k):
    w = rnd(cout, cin, k, k, seed=2)
    s, p = (2, 2) if k == 4 else (1, 0)
    wf, wq = ops.prep_conv_weights(w, s, p)
    wr = tf32_rna(w)
    assert torch.equal(wf, wr.permute(0, 2, 3, 1).reshape(cout, -1))
    if k == 1:
        assert torch.equal(wq, wr.view(cout, cin).t())
    else:
        # wq[(ph, pw, ci)][(a, b, co)] = w[co, ci, ph + 2(1-a), pw + 2(1-b)]
        want = torch.empty(2, 2, cin, 2, 2, cout, device=DEV)
        for ph in range(2):
            for pw in range(2):
                for a in range(2):
                    for b in range(2):
                        want[ph, pw, :, a, b, :] = wr[:, :, ph + 2 * (1 - a), pw + 2 * (1 - b)].t()
        assert torch.equal(wq, want.view(4 * cin, 4 * cout))


@pytest.mark.parametrize("B,C,H,W", [(3, 16, 33, 45), (2, 8, 129, 174), (5, 512, 3, 4), (2, 24, 7, 5)])
def test_batchnorm2d_channels_last(B, C, H, W):
    a = ops.to_cl(F.leaky_relu(rnd(B, C, H, W, seed=7) * 2 + 0.5, 0.1))
    bn = torch.nn.BatchNorm2d(C).to(DEV)
    with torch.no_grad():
        bn.weight.copy_(rnd(C, seed=8) * 0.3 + 1)
        bn.bias.copy_(rnd(C, seed=9) * 0.2)
    ref_bn = torch.nn.BatchNorm2d(C).to(DEV).double()
    ref_bn.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in bn.state_dict().items()})
    y, mean, rstd = ops.bn2d_train_fwd(a, bn)
    ref = ref_bn(a.double())
    assert ops.is_cl(y)
    assert torch.equal(y, tf32_rna(y))                       # the channels-last route hands TF32-rounded operands to the next conv
    assert rel(y, ref) < 4e-4                                  # 2^-11 rounding
    assert rel(bn.running_mean, ref_bn.running_mean) < 1e-6 and rel(bn.running_var, ref_bn.running_var) < 1e-6
    assert rel(mean, a.double().mean((0, 2, 3))) < 1e-6
    dy = rnd(B, C, H, W, seed=10)                              # NCHW on purpose: the op converts to the layout of `a`
    z = (a / torch.where(a > 0, torch.ones_like(a), torch.full_like(a, 0.1))).double().requires_grad_()
    out = ref_bn.train()(F.leaky_relu(z, 0.1))
    gz, gg, gb = torch.autograd.grad(out, (z, ref_bn.weight, ref_bn.bias), dy.double())
    dz, dg, db = ops.bn2d_train_bwd(dy, a, bn.weight, mean, rstd, 0.1)
    assert ops.is_cl(dz) and rel(dz, gz) < 4e-4 and rel(dg, gg) < 5e-6 and rel(db, gb) < 5e-6
    dz2, _, _, colsum = ops.bn2d_train_bwd(dy, a, bn.weight, mean, rstd, 0.1, want_colsum=True)      # + bias gradient of the conv in front
    assert torch.equal(dz2, dz)
    assert float((colsum.double() - dz.double().sum((0, 2, 3))).abs().max()) <= 1e-5 * float(dz.double().abs().sum((0, 2, 3)).max())
    bn.eval()
    ref_bn.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in bn.state_dict().items()})
    ref_bn.eval()
    assert rel(ops.bn2d_eval_fwd(a, bn), ref_bn(a.double())) < 4e-4
    assert rel(ops.lrelu_bwd(dy, a, 0.1), dy * torch.where(a > 0, 1.0, 0.1)) < 4e-4
    assert rel(ops.channel_sum(ops.to_cl(dy)), dy.double().sum((0, 2, 3))) < 1e-6


def test_thin_layer_kernels_with_channels_last_operands():
    """enc1 writes / dec8 reads the 8-channel tensor in channels-last order on the default route."""
    B, H, W, C = 3, 257, 347, 8
    assert ops.cl_mode()
    x, w, b = rnd(B, 1, H, W, seed=60), rnd(C, 1, 5, 5, seed=61, scale=0.2), rnd(C, seed=62)
    xd, wd, bd = x.double().requires_grad_(), w.double().requires_grad_(), b.double().requires_grad_()
    pre = F.conv2d(xd, wd, bd, 2, 2)
    y = ops.conv2d_fwd(x, w, b, 2, 2, slope=0.1)
    assert ops.is_cl(y) and rel(y, F.leaky_relu(pre, 0.1)) < 2e-6
    yr = ops.conv2d_fwd(x, w, b, 2, 2, slope=0.1, round_out=True)
    assert torch.equal(yr, tf32_rna(y))
    dy = ops.to_cl(rnd(B, C, 129, 174, seed=63))
    gx, gw, gb = torch.autograd.grad(pre, (xd, wd, bd), dy.double())
    dw, db = ops.conv2d_wgrad(x, dy, w.shape, 2, 2, want_bias=True)
    assert rel(dw, gw) < 5e-6 and rel(db, gb) < 5e-6
    bias1 = rnd(1, seed=64)
    t = ops.conv2d_dgrad(dy, w, (H, W), 2, 2, bias=bias1, clamp=(-1.0, 1.0))      # dec8 forward + Hardtanh
    assert t.is_contiguous() and rel(t, F.hardtanh(gx + bias1.double())) < 2e-6


@pytest.mark.parametrize("transposed", [False, True])
def test_block_forward_backward_on_the_channels_last_route(transposed):
    """A whole Conv2D / TConv2D block (conv -> LeakyReLU -> BatchNorm2d, weight re-packing, rounding flags, layout
    conversion of an NCHW input and of the NCHW upstream gradient) against torch's own modules in fp64."""
    torch.manual_seed(5)
    B = 4
    if transposed:
        blk = layer.TConv2D(64, 32, [4, 4], [2, 2], 2, output_padding=[1, 1], activation=torch.nn.LeakyReLU(0.1), name_prefix='t').to(DEV)
        x = rnd(B, 64, 17, 23, seed=1)
        ref_conv = torch.nn.ConvTranspose2d(64, 32, 4, 2, 2, output_padding=1).to(DEV).double()
    else:
        blk = layer.Conv2D(32, 64, [4, 4], [2, 2], 2, [1, 1], activation=torch.nn.LeakyReLU(0.1), name_prefix='c').to(DEV)
        x = rnd(B, 32, 33, 45, seed=1)
        ref_conv = torch.nn.Conv2d(32, 64, 4, 2, 2).to(DEV).double()
    ref_bn = torch.nn.BatchNorm2d(blk.bn.num_features).to(DEV).double()
    with torch.no_grad():
        ref_conv.weight.copy_(blk.conv.weight.double()); ref_conv.bias.copy_(blk.conv.bias.double())
    blk.train()
    xg = x.clone().requires_grad_()
    y = blk(xg)
    xd = x.double().requires_grad_()
    ref = ref_bn(F.leaky_relu(ref_conv(xd), 0.1))
    assert y.shape == ref.shape and rel(y, ref) < 2e-3
    # Gradient gate: TF32 noise (3e-4) in the pre-activation flips the LeakyReLU branch of the ~3e-4 fraction of elements that
    # lie that close to zero; each flip is an O(1) relative error on that element, i.e. ~sqrt(3e-4) * 0.9 = 1.5e-2 in L2.
    dy = rnd(*y.shape, seed=3)
    y.backward(dy)
    ref.backward(dy.double())
    assert rel(xg.grad, xd.grad) < 3e-2
    assert rel(blk.conv.weight.grad, ref_conv.weight.grad) < 3e-2
    assert rel(blk.bn.weight.grad, ref_bn.weight.grad) < 3e-2 and rel(blk.bn.bias.grad, ref_bn.bias.grad) < 3e-2
    assert rel(blk.conv.bias.grad, ref_conv.bias.grad) < 1e-1       # a sum with heavy cancellation over B*H*W pixels


@pytest.mark.parametrize("cin,cout,H,W", [(8, 16, 129, 174), (16, 32, 65, 88), (32, 64, 33, 45), (64, 128, 17, 23), (128, 256, 9, 12)])
def test_conv_epilogue_accumulates_the_batchnorm_statistics(cin, cout, H, W):
    """pgv_conv_cl_fwd_bn / pgv_conv_cl_dgrad_bn: sums[c] = (sum, sum of squares) of exactly the values the kernel stored, for the
    16-column direct epilogue, the staged epilogue (1 to 4 column chunks) and the quad epilogue with odd output sizes; BatchNorm
    from these sums equals BatchNorm from its own reduction pass.  Launches with several N tiles have no such by-product."""
    B = 3
    x, w, b = rnd(B, cin, H, W, seed=70), rnd(cout, cin, 4, 4, seed=71, scale=0.1), rnd(cout, seed=72)
    y, sums = ops.conv2d_fwd(x, w, b, 2, 2, slope=0.1, bn_sums=True)
    assert torch.equal(y, ops.conv2d_fwd(x, w, b, 2, 2, slope=0.1))
    if cout > 128:
        assert sums is None
        return
    yd = y.double()
    want = torch.stack([yd.sum((0, 2, 3)), (yd * yd).sum((0, 2, 3))], 1).reshape(-1)
    scale = torch.stack([yd.abs().sum((0, 2, 3)), (yd * yd).sum((0, 2, 3))], 1).reshape(-1)
    assert float(((sums - want).abs() / scale).max()) < 2e-6
    bn = torch.nn.BatchNorm2d(cout).to(DEV)
    bn2 = torch.nn.BatchNorm2d(cout).to(DEV)
    o1, m1, r1 = ops.bn2d_train_fwd(y, bn)
    o2, m2, r2 = ops.bn2d_train_fwd(y, bn2, sums)
    assert rel(o2, o1) < 1e-5 and rel(m2, m1) < 1e-5 and rel(r2, r1) < 1e-5 and rel(bn2.running_var, bn.running_var) < 1e-5
    # transposed convolution = data gradient of the same weights: input [B, cout, Ho, Wo] -> [B, cin, H, W] (H, W odd: partial quads)
    dy, bt = rnd(B, cout, y.shape[2], y.shape[3], seed=73), rnd(cin, seed=74)
    t, tsums = ops.conv2d_dgrad(dy, w, (H, W), 2, 2, bias=bt, slope=0.1, bn_sums=True)
    assert torch.equal(t, ops.conv2d_dgrad(dy, w, (H, W), 2, 2, bias=bt, slope=0.1))
    if 4 * cin > 128:
        assert tsums is None
        return
    td = t.double()
    want = torch.stack([td.sum((0, 2, 3)), (td * td).sum((0, 2, 3))], 1).reshape(-1)
    scale = torch.stack([td.abs().sum((0, 2, 3)), (td * td).sum((0, 2, 3))], 1).reshape(-1)
    assert float(((tsums - want).abs() / scale).max()) < 2e-6


@pytest.mark.parametrize("transposed", [False, True])
def test_chain_backward_with_fused_batchnorm_sums_equals_the_unfused_chain(transposed):
    """layer.chain_bwd with ops.fuse_bn_bwd on (each data-gradient kernel hands the BatchNorm-backward sums to the block in front) and off
    (every block reduces for itself): same input gradient, same parameter gradients, for a Conv2D chain and a TConv2D chain."""
    torch.manual_seed(7)
    B = 4
    act = torch.nn.LeakyReLU(0.1)
    if transposed:
        blocks = [layer.TConv2D(64, 32, [4, 4], [2, 2], 2, output_padding=[1, 1], activation=act, name_prefix='a').to(DEV),
                  layer.TConv2D(32, 16, [4, 4], [2, 2], 2, output_padding=[1, 0], activation=act, name_prefix='b').to(DEV),
                  layer.TConv2D(16, 8, [4, 4], [2, 2], 2, output_padding=[1, 0], activation=act, name_prefix='c').to(DEV)]
        x = rnd(B, 64, 17, 23, seed=1)
    else:
        blocks = [layer.Conv2D(8, 16, [4, 4], [2, 2], 2, [1, 1], activation=act, name_prefix='a').to(DEV),
                  layer.Conv2D(16, 32, [4, 4], [2, 2], 2, [1, 1], activation=act, name_prefix='b').to(DEV),
                  layer.Conv2D(32, 64, [4, 4], [2, 2], 2, [1, 1], activation=act, name_prefix='c').to(DEV)]
        x = rnd(B, 8, 129, 174, seed=1)
    ctxs, h = [], ops.to_cl(x, True)
    for blk in blocks:
        h, c = blk.fwd(h, True)
        ctxs.append(c)
    dy = ops.to_cl(rnd(*h.shape, seed=2))
    res = {}
    for fused in (True, False):
        ops.fuse_bn_bwd = fused
        try:
            grads = {}
            dx = layer.chain_bwd(blocks, ctxs, dy, grads, True)
            ops.join_forks(dy)
            torch.cuda.synchronize()
            res[fused] = (dx.clone(), {k: v.clone() for k, v in grads.items() if v is not None})
        finally:
            ops.fuse_bn_bwd = False
    assert rel(res[True][0], res[False][0]) < 2e-5
    assert res[True][1].keys() == res[False][1].keys() and len(res[True][1]) >= 10
    for k, v in res[False][1].items():
        assert rel(res[True][1][k], v) < 5e-5


@pytest.mark.parametrize("cin,cout,H,W", [(8, 16, 129, 174), (16, 32, 65, 88), (32, 64, 33, 45), (64, 128, 17, 23)])
def test_data_gradient_epilogue_accumulates_the_batchnorm_backward_sums(cin, cout, H, W):
    """bn_bwd_x: the kernel that writes the gradient flowing into a BatchNorm2d also accumulates (sum dy, sum dy * x) with x that
    BatchNorm's input; BatchNorm backward from these raw sums equals BatchNorm backward with its own reduction pass.  Both forms of the
    data gradient: the quad GEMM of a Conv2D block (odd sizes: partial quads) and the forward-conv form of a TConv2D block."""
    B = 3
    w = rnd(cout, cin, 4, 4, seed=81, scale=0.1)
    Ho, Wo = (H + 4 - 4) // 2 + 1, (W + 4 - 4) // 2 + 1
    for form in ('quad', 'conv'):
        if form == 'quad':              # dx [B, cin, H, W] of a convolution; the block in front normalised a [B, cin, H, W] activation
            g_in, act_shape = rnd(B, cout, Ho, Wo, seed=82), (B, cin, H, W)
            run = lambda xa: ops.conv2d_dgrad(g_in, w, (H, W), 2, 2, bn_bwd_x=xa)
            plain = lambda: ops.conv2d_dgrad(g_in, w, (H, W), 2, 2)
            eligible = 4 * cin <= 128
        else:                           # data gradient of a transposed convolution = forward convolution: [B, cin, H, W] -> [B, cout, Ho, Wo]
            g_in, act_shape = rnd(B, cin, H, W, seed=83), (B, cout, Ho, Wo)
            run = lambda xa: ops.conv2d_fwd(g_in, w, None, 2, 2, bn_bwd_x=xa)
            plain = lambda: ops.conv2d_fwd(g_in, w, None, 2, 2)
            eligible = cout <= 128
        a = ops.to_cl(rnd(*act_shape, seed=84) + 0.3)
        dy, sums = run(a)
        assert torch.equal(dy, plain())
        if not eligible:
            assert sums is None
            continue
        d, ad = dy.double(), a.double()
        want = torch.stack([d.sum((0, 2, 3)), (d * ad).sum((0, 2, 3))], 1).reshape(-1)
        scale = torch.stack([d.abs().sum((0, 2, 3)), (d * ad).abs().sum((0, 2, 3))], 1).reshape(-1)
        assert float(((sums - want).abs() / scale).max()) < 2e-6
        C = act_shape[1]
        gamma = rnd(C, seed=85) * 0.2 + 1.0
        mean = a.double().mean((0, 2, 3)).float()
        rstd = (1.0 / (a.double().var((0, 2, 3), unbiased=False) + 1e-5).sqrt()).float()
        r1 = ops.bn2d_train_bwd(dy, a, gamma, mean, rstd, 0.1, want_colsum=True)
        r2 = ops.bn2d_train_bwd(dy, a, gamma, mean, rstd, 0.1, want_colsum=True, raw_sums=sums)
        for u, v in zip(r2, r1):
            assert rel(u, v) < 2e-5, form


@pytest.mark.parametrize("m,n,k", [(160, 1220, 24576), (160, 24576, 610), (64, 1024, 1030), (62, 1024, 1030), (5, 36, 26)])
def test_big_linear_layers_on_the_channels_last_kernel(m, n, k):
    """encoder / decoder FC (and a padded-K case) through fc_fwd / fc_bwd: rounded + re-pitched operand copies, weights by TMA,
    K tail, direct weight-gradient destination; the data gradient on the weight-gradient form of the kernel (M % 4 == 0: W as stored
    and dy^T, padded K of the decoder FC sliced away) and on the transposed-weight form (M = 62, unaligned-N atomic epilogue)."""
    x, w, b = rnd(m, k, seed=1), rnd(n, k, seed=2, scale=0.05), rnd(n, seed=3)
    dy = rnd(m, n, seed=4)
    route = ops.fc_route(m, n, k)
    assert route == ('generic' if m == 5 else 'cl')
    y, ctx = ops.fc_fwd(x, w, b, True)
    if route == 'cl':
        assert (ctx[1] is None) == (m % 4 == 0) and (ctx[4] is None) == (m % 4 != 0)
    out = torch.full((n, k), 7.0, device=DEV)
    dx, dw, db = ops.fc_bwd(dy, ctx, w, True, out=out)
    assert dw.data_ptr() == out.data_ptr()
    errs = (rel(y, x.double() @ w.double().T + b.double()), rel(dx, dy.double() @ w.double()), rel(dw, dy.double().T @ x.double()),
            rel(db, dy.double().sum(0)))
    print("fc", route, (m, n, k), "rel-L2 fwd %.2e dgrad %.2e wgrad %.2e db %.2e" % errs)
    assert max(errs) < 2e-3 and dx.is_contiguous() and dx.shape == (m, k)
    y_eval, ctx_eval = ops.fc_fwd(x, w, b, False)
    assert ctx_eval[1] is None and rel(y_eval, y) < 1e-6


# every channels-last layer geometry of the two CNNs as (Cin, Cout, k, stride, pad, H, W) of the underlying convolution
# (encoder.py:233-259; the decoder's transposed convolutions are the data gradients of the same shapes, decoder.py:199-220)
CL_GEOMS = [(8, 16, 4, 2, 2, 129, 174), (16, 32, 4, 2, 2, 65, 88), (32, 64, 4, 2, 2, 33, 45), (64, 128, 4, 2, 2, 17, 23),
            (128, 256, 4, 2, 2, 9, 12), (256, 512, 4, 2, 2, 5, 7), (512, 2048, 1, 1, 0, 3, 4), (1536, 768, 4, 2, 2, 5, 7)]


def _set_a_mode(mode):
    from preset_gen_vae_b200 import _lib
    _lib.check(_lib.lib().pgv_debug_set_conv_a_mode(mode))


@pytest.mark.parametrize("B", [3, 40])
@pytest.mark.parametrize("cin,cout,k,s,p,H,W", CL_GEOMS)
def test_tma_fed_activation_operand_equals_the_cpasync_gather(cin, cout, k, s, p, H, W, B):
    """The TMA-fed activation operand (tiled boxes for 1x1, im2col-mode loads for the 4x4 / stride-2 window and the 2x2 window of the
    data gradient; 8 / 16-channel tiles with 32 / 64-byte swizzles) must reproduce the cp.async gather BIT FOR BIT: both fill the
    same shared-memory tiles for the same sequence of MMAs.  Covers padding, odd sizes (partial quads), row and image wrap-around
    inside a 128-pixel tile, M tails, small tensors (< 128 KB) and split-K through the workspace."""
    if B == 40 and cin * H * W > 400000:
        B = 12
    Ho, Wo = ops.conv_out_size(H, k, s, p), ops.conv_out_size(W, k, s, p)
    x, w, b = rnd(B, cin, H, W, seed=21), rnd(cout, cin, k, k, seed=22, scale=0.1), rnd(cout, seed=23)
    dy, b_in = rnd(B, cout, Ho, Wo, seed=24), rnd(cin, seed=25)
    assert ops.conv_route(cin, cout, k, k, s, p, H, W, Ho, Wo) == 'cl'
    got = {}
    try:
        for mode in (0, -1):
            _set_a_mode(mode)
            got[mode] = (ops.conv2d_fwd(x, w, b, s, p, slope=0.1), ops.conv2d_dgrad(dy, w, (H, W), s, p, bias=b_in, slope=0.1))
    finally:
        _set_a_mode(-1)
    torch.cuda.synchronize()
    assert torch.equal(got[0][0], got[-1][0]), "forward: TMA-fed A differs from the cp.async gather (max %g)" % float((got[0][0] - got[-1][0]).abs().max())
    assert torch.equal(got[0][1], got[-1][1]), "dgrad: TMA-fed A differs from the cp.async gather (max %g)" % float((got[0][1] - got[-1][1]).abs().max())
    pre = F.conv2d(x.double(), w.double(), b.double(), s, p)
    assert rel(got[-1][0], F.leaky_relu(pre, 0.1)) < 2e-3


@pytest.mark.parametrize("cin,cout,k,s,p,H,W", CL_GEOMS)
def test_split_k_through_the_workspace_is_deterministic(cin, cout, k, s, p, H, W):
    """Deterministic mode (default): forward, data gradient and weight gradient are bit-identical run to run (fixed-order split-K,
    no atomics), write the weight gradient directly in the PyTorch layout, and agree with the atomic path of round 1."""
    B = 16
    Ho, Wo = ops.conv_out_size(H, k, s, p), ops.conv_out_size(W, k, s, p)
    x, w, b = rnd(B, cin, H, W, seed=31), rnd(cout, cin, k, k, seed=32, scale=0.1), rnd(cout, seed=33)
    dy = rnd(B, cout, Ho, Wo, seed=34)

    def run():
        out = torch.full(w.shape, 7.0, device=DEV)
        dw, _ = ops.conv2d_wgrad(x, dy, w.shape, s, p, want_bias=False, out=out)
        assert dw.data_ptr() == out.data_ptr()
        return ops.conv2d_fwd(x, w, b, s, p, slope=0.1), ops.conv2d_dgrad(dy, w, (H, W), s, p), dw
    assert ops.deterministic
    a, b2 = run(), run()
    for u, v in zip(a, b2):
        assert torch.equal(u, v)
    ops.deterministic = False
    try:
        c = run()
    finally:
        ops.deterministic = True
    for u, v in zip(a, c):
        assert rel(u, v) < 2e-4                    # different summation order of the K splits only
    wd = w.double().requires_grad_()
    gw, = torch.autograd.grad(F.conv2d(x.double(), wd, None, s, p), wd, dy.double())
    assert rel(a[2], gw) < 2e-3
