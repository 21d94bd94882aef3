"""GPU: every model-path kernel family, called through the C ABI (model/ops.py -> ctypes -> libpgv.so), against an
fp64 PyTorch evaluation of the same operation.  Tolerances are fp32 rounding (these kernels use exact fp32 products)."""
from types import SimpleNamespace

import pytest
import torch
import torch.nn.functional as F

from oracle import losses as oloss, nflows_port as nf
from preset_gen_vae_b200.model import ops

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return torch.randn(*shape, device=DEV, generator=g) * scale


# (Cin, Cout, k, stride, pad, H, W): every conv geometry of the encoder (encoder.py:233-259, 60-70)
ENC_LAYERS = [(1, 8, 5, 2, 2, 257, 347), (8, 16, 4, 2, 2, 129, 174), (16, 32, 4, 2, 2, 65, 88), (32, 64, 4, 2, 2, 33, 45),
              (64, 128, 4, 2, 2, 17, 23), (128, 256, 4, 2, 2, 9, 12), (256, 512, 4, 2, 2, 5, 7), (512, 2048, 1, 1, 0, 3, 4)]


@pytest.mark.parametrize("cin,cout,k,s,p,H,W", ENC_LAYERS)
def test_conv_fwd_dgrad_wgrad(cin, cout, k, s, p, H, W):
    B = 2
    ops.set_precision('fp32')
    ops.use_thin = False          # this test covers the generic exact-fp32 kernels; the thin kernels have their own test
    x, w, b = rnd(B, cin, H, W, seed=1), rnd(cout, cin, k, k, seed=2, scale=0.1), rnd(cout, seed=3)
    y = ops.conv2d_fwd(x, w, b, s, p, slope=0.1)
    xd, wd, bd = x.double().requires_grad_(), w.double().requires_grad_(), b.double().requires_grad_()
    ref = F.leaky_relu(F.conv2d(xd, wd, bd, s, p), 0.1)
    assert y.shape == ref.shape and rel(y, ref) < 2e-6
    dy = rnd(*y.shape, seed=4)
    pre = F.conv2d(xd, wd, bd, s, p)
    gx, gw, gb = torch.autograd.grad(pre, (xd, wd, bd), dy.double())
    assert rel(ops.conv2d_dgrad(dy, w, (H, W), s, p), gx) < 2e-6
    dw, db = ops.conv2d_wgrad(x, dy, w.shape, s, p, want_bias=True)
    ops.set_precision('tf32')
    ops.use_thin = True
    assert rel(dw, gw) < 5e-6 and rel(db, gb) < 5e-6


def test_thin_layer_kernels():
    """enc1 = Conv2d(1,8,5,2,2) and dec8 = ConvTranspose2d(8,1,5,2,2)+Hardtanh on the direct streaming kernels."""
    B, H, W, C = 3, 257, 347, 8
    x, w, b = rnd(B, 1, H, W, seed=60), rnd(C, 1, 5, 5, seed=61, scale=0.2), rnd(C, seed=62)
    assert ops.use_thin and ops._thin(1, C, 5, 5, 2, 2, H, W, 129, 174)
    y = ops.conv2d_fwd(x, w, b, 2, 2, slope=0.1)
    xd, wd, bd = x.double().requires_grad_(), w.double().requires_grad_(), b.double().requires_grad_()
    pre = F.conv2d(xd, wd, bd, 2, 2)
    assert y.shape == (B, C, 129, 174) and rel(y, F.leaky_relu(pre, 0.1)) < 2e-6
    dy = rnd(B, C, 129, 174, seed=63)
    gx, gw, gb = torch.autograd.grad(pre, (xd, wd, bd), dy.double())
    dw, db = ops.conv2d_wgrad(x, dy, w.shape, 2, 2, want_bias=True)
    assert rel(dw, gw) < 5e-6 and rel(db, gb) < 5e-6
    bias1 = rnd(1, seed=64)
    t = ops.conv2d_dgrad(dy, w, (H, W), 2, 2, bias=bias1, clamp=(-1.0, 1.0))      # dec8 forward + Hardtanh
    want = F.hardtanh(gx + bias1.double())
    assert rel(t, want) < 2e-6
    assert rel(ops.conv2d_dgrad(dy, w, (H, W), 2, 2), gx) < 2e-6


# (Cin, Cout, k, output_padding, Hin, Win): every transposed conv of the decoder (decoder.py:72-75, 205-218)
DEC_LAYERS = [(2048, 512, 1, (0, 0), 1, 3, 4), (512, 256, 4, (1, 1), 2, 3, 4), (256, 128, 4, (1, 0), 2, 5, 7),
              (128, 64, 4, (1, 1), 2, 9, 12), (64, 32, 4, (1, 1), 2, 17, 23), (32, 16, 4, (1, 0), 2, 33, 45),
              (16, 8, 4, (1, 0), 2, 65, 88), (8, 1, 5, (0, 0), 2, 129, 174)]


@pytest.mark.parametrize("cin,cout,k,op,s,H,W", DEC_LAYERS)
def test_transposed_conv_via_conv_gradients(cin, cout, k, op, s, H, W):
    from preset_gen_vae_b200.model import layer
    ops.set_precision('fp32')
    B, p = 2, (2 if k > 1 else 0)
    conv = torch.nn.ConvTranspose2d(cin, cout, k, s, p, op).to(DEV)
    x = rnd(B, cin, H, W, seed=5)
    y = layer.tconv_fwd(x, conv, slope=0.1)
    xd = x.double().requires_grad_()
    wd, bd = conv.weight.detach().double().requires_grad_(), conv.bias.detach().double().requires_grad_()
    pre = F.conv_transpose2d(xd, wd, bd, s, p, op)
    assert y.shape == pre.shape and rel(y, F.leaky_relu(pre, 0.1)) < 2e-6
    dz = rnd(*y.shape, seed=6)
    gx, gw, gb = torch.autograd.grad(pre, (xd, wd, bd), dz.double())
    grads = {}
    dx = layer.tconv_bwd(dz, x, conv, grads, True)
    ops.set_precision('tf32')
    assert rel(dx, gx) < 2e-6 and rel(grads[id(conv.weight)], gw) < 5e-6 and rel(grads[id(conv.bias)], gb) < 5e-6
    expected = {1: (3, 4), 4: None}   # output sizes of decoder.py:199-220 are checked in the model test


@pytest.mark.parametrize("use_cl", [True, False])
@pytest.mark.parametrize("cin,cout,k,s,p,H,W", ENC_LAYERS + [(8, 1, 5, 2, 2, 257, 347), (16, 8, 4, 2, 2, 129, 174)])
def test_tensor_core_convs_against_fp64(cin, cout, k, s, p, H, W, use_cl):
    """tcgen05 implicit-GEMM conv fwd / dgrad / wgrad (TF32 products, operands rounded to nearest) for every layer
    geometry, incl. the decoder's last two transposed convs viewed as conv data-gradients, on both tensor-core routes:
    channels-last + cp.async (use_cl, the default) and NCHW with register-staged gathers.  TF32 tolerance: relative-L2
    error of a K-term dot product of operands with 2^-11 relative rounding ~ 2^-11 * sqrt(2) = 7e-4; gate 2e-3."""
    B = 3
    ops.use_cl = use_cl
    Ho, Wo = ops.conv_out_size(H, k, s, p), ops.conv_out_size(W, k, s, p)
    if (cin, cout, k) == (16, 8, 4):            # dec7 geometry: odd output_padding makes H one larger than the conv's natural input
        H, W = 129, 174
        Ho, Wo = 65, 88
    x, w, b = rnd(B, cin, H, W, seed=1), rnd(cout, cin, k, k, seed=2, scale=0.1), rnd(cout, seed=3)
    dy = rnd(B, cout, Ho, Wo, seed=4)
    xd, wd, bd = x.double().requires_grad_(), w.double().requires_grad_(), b.double().requires_grad_()
    pre = F.conv2d(xd, wd, bd, s, p)[:, :, :Ho, :Wo]
    gx, gw, gb = torch.autograd.grad(pre, (xd, wd, bd), dy.double())
    ops.set_precision('tf32')
    try:
        y = ops.conv2d_fwd(x, w, b, s, p, slope=0.1, out_hw=(Ho, Wo))
        dx = ops.conv2d_dgrad(dy, w, (H, W), s, p)
        dw, db = ops.conv2d_wgrad(x, dy, w.shape, s, p, want_bias=True)
        bias_in = rnd(cin, seed=5)
        dx_act = ops.conv2d_dgrad(dy, w, (H, W), s, p, bias=bias_in, slope=0.1)     # transposed-conv forward form
        route = ops.conv_route(cin, cout, k, k, s, p, H, W, Ho, Wo)
    finally:
        ops.set_precision('tf32')
        ops.use_cl = True
    assert route == ('thin' if (cin == 1 and k == 5) else ('cl' if (use_cl and k != 5) else 'tc'))
    assert ops.is_cl(y) == (use_cl and route in ('thin', 'cl')) and ops.is_cl(dx) == (route == 'cl')
    errs = (rel(y, F.leaky_relu(pre, 0.1)), rel(dx, gx), rel(dw, gw), rel(db, gb),
            rel(dx_act, F.leaky_relu(gx + bias_in.double()[None, :, None, None], 0.1)))
    print("tc conv", route, (cin, cout, k, s, p, H, W), "rel-L2 fwd %.2e dgrad %.2e wgrad %.2e db %.2e tconv-fwd %.2e" % errs)
    assert max(errs) < 2e-3


def test_batchnorm2d_train_eval_backward():
    B, C, H, W = 3, 16, 33, 45
    a = F.leaky_relu(rnd(B, C, H, W, seed=7) * 2 + 0.5, 0.1)
    bn = torch.nn.BatchNorm2d(C).to(DEV)
    with torch.no_grad():
        bn.weight.copy_(rnd(C, seed=8) * 0.3 + 1)
        bn.bias.copy_(rnd(C, seed=9) * 0.2)
    ref_bn = torch.nn.BatchNorm2d(C).to(DEV).double()
    ref_bn.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in bn.state_dict().items()})
    y, mean, rstd = ops.bn2d_train_fwd(a, bn)
    ad = a.double().requires_grad_()
    ref = ref_bn(ad)
    assert rel(y, ref) < 2e-6
    assert rel(bn.running_mean, ref_bn.running_mean) < 1e-6 and rel(bn.running_var, ref_bn.running_var) < 1e-6
    dy = rnd(B, C, H, W, seed=10)
    # fused BN backward + LeakyReLU backward (through the activation that produced `a`)
    z = (a / torch.where(a > 0, torch.ones_like(a), torch.full_like(a, 0.1))).double().requires_grad_()
    out = ref_bn.train()(F.leaky_relu(z, 0.1))
    gz, gg, gb = torch.autograd.grad(out, (z, ref_bn.weight, ref_bn.bias), dy.double())
    dz, dg, db = ops.bn2d_train_bwd(dy, a, bn.weight, mean, rstd, 0.1)
    assert rel(dz, gz) < 5e-6 and rel(dg, gg) < 5e-6 and rel(db, gb) < 5e-6
    bn.eval()
    ref_bn.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in bn.state_dict().items()})
    ref_bn.eval()
    assert rel(ops.bn2d_eval_fwd(a, bn), ref_bn(a.double())) < 2e-6
    assert rel(ops.lrelu_bwd(dy, a, 0.1), dy * torch.where(a > 0, 1.0, 0.1)) < 1e-7


@pytest.mark.parametrize("relu,use_mask", [(False, False), (True, False), (True, True)])
def test_batchnorm1d_fused(relu, use_mask):
    B, Fe = 37, 300
    x = rnd(B, Fe, seed=11) * 1.5 + 0.3
    bn = torch.nn.BatchNorm1d(Fe, eps=1e-3).to(DEV)
    with torch.no_grad():
        bn.weight.copy_(rnd(Fe, seed=12) * 0.3 + 1)
        bn.bias.copy_(rnd(Fe, seed=13) * 0.2)
    mask = (torch.empty(B, Fe, device=DEV).bernoulli_(0.6) / 0.6) if use_mask else None
    ref_bn = torch.nn.BatchNorm1d(Fe, eps=1e-3).to(DEV).double()
    ref_bn.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in bn.state_dict().items()})
    y, mean, rstd = ops.bn1d_train_fwd(x, bn, relu=relu, mask=mask)
    xd = x.double().requires_grad_()
    ref = ref_bn(xd)
    if relu:
        ref = F.relu(ref)
    if use_mask:
        ref = ref * mask.double()
    assert rel(y, ref) < 2e-6 and rel(bn.running_var, ref_bn.running_var) < 1e-6
    dy = rnd(B, Fe, seed=14)
    gx, gg, gb = torch.autograd.grad(ref, (xd, ref_bn.weight, ref_bn.bias), dy.double())
    dx, dg, db = ops.bn1d_train_bwd(dy, x, bn, mean, rstd, relu=relu, mask=mask)
    assert rel(dx, gx) < 5e-6 and rel(dg, gg) < 5e-6 and rel(db, gb) < 5e-6
    bn.eval(); ref_bn.eval()
    e = ref_bn(x.double())
    assert rel(ops.bn1d_eval_fwd(x, bn, relu=relu), F.relu(e) if relu else e) < 2e-6


def test_flow_batchnorm_transform_matches_nflows_restatement():
    from preset_gen_vae_b200.model.flows import BatchNorm
    B, Fe = 23, 610
    x = rnd(B, Fe, seed=15) * 2 + 1
    mine = BatchNorm(Fe).to(DEV)
    ref = nf.BatchNorm(Fe).to(DEV).double()
    with torch.no_grad():
        mine.unconstrained_weight.add_(rnd(Fe, seed=16) * 0.3)
        mine.bias.add_(rnd(Fe, seed=17) * 0.2)
        ref.unconstrained_weight.copy_(mine.unconstrained_weight.double())
        ref.bias.copy_(mine.bias.double())
    y, mean, var, ld = ops.flowbn_train_fwd(x, mine)
    xd = x.double().requires_grad_()
    ry, rld = ref(xd)
    assert rel(y, ry) < 2e-6 and abs(ld.item() - rld[0].item()) < 1e-3 * abs(rld[0].item()) + 1e-4
    assert rel(mine.running_var, ref.running_var) < 1e-6 and rel(mine.running_mean, ref.running_mean) < 1e-6
    dy, dld = rnd(B, Fe, seed=18), rnd(B, seed=19)
    gx, gu, gb = torch.autograd.grad([ry, rld], (xd, ref.unconstrained_weight, ref.bias), [dy.double(), dld.double()])
    dx, du, db = ops.flowbn_train_bwd(dy, x, mine, mean, var, dld.sum().reshape(1))
    assert rel(dx, gx) < 1e-5 and rel(du, gu) < 1e-5 and rel(db, gb) < 1e-5
    mine.eval(); ref.eval()
    ye, lde = ops.flowbn_eval(x, mine)
    rye, rlde = ref(x.double())
    assert rel(ye, rye) < 2e-6 and abs(lde.item() - rlde[0].item()) < 1e-3
    back, ldi = ops.flowbn_eval(ye, mine, inverse=True)
    assert rel(back, x) < 1e-5 and abs(ldi.item() + lde.item()) < 1e-3


def test_affine_coupling_forward_backward_inverse():
    B, D = 9, 610
    x, prm = rnd(B, D, seed=20), rnd(B, D, seed=21)
    ident = torch.arange(0, D, 2, device=DEV, dtype=torch.int32)
    trans = torch.arange(1, D, 2, device=DEV, dtype=torch.int32)
    ld_in = rnd(B, seed=22)
    y, ld = ops.coupling_fwd(x, prm, ident, trans, ld_in)
    xd, pd = x.double().requires_grad_(), prm.double().requires_grad_()
    n_t = trans.numel()
    s = torch.sigmoid(pd[:, n_t:] + 2) + 1e-3
    ry = xd.clone()
    ry[:, 1::2] = xd[:, 1::2] * s + pd[:, :n_t]
    rld = ld_in.double() + torch.log(s).sum(1)
    assert rel(y, ry) < 1e-6 and rel(ld, rld) < 1e-6
    dy, dld = rnd(B, D, seed=23), rnd(B, seed=24)
    gx, gp = torch.autograd.grad([ry, rld], (xd, pd), [dy.double(), dld.double()])
    dx, dp = ops.coupling_bwd(dy, dld, x, prm, ident, trans)
    assert rel(dx, gx) < 1e-6 and rel(dp, gp) < 2e-6
    back, ld_back = ops.coupling_fwd(y, prm, ident, trans, ld, inverse=True)
    assert rel(back, x) < 1e-5 and rel(ld_back, ld_in) < 1e-4
    g = ops.gather_cols(x, trans)
    assert torch.equal(g, x[:, 1::2])
    acc = x.clone()
    ops.scatter_add_cols_(acc, trans, g)
    assert torch.allclose(acc[:, 1::2], 2 * x[:, 1::2]) and torch.equal(acc[:, 0::2], x[:, 0::2])


def test_reparam_hardtanh_elementwise():
    B, D = 7, 610
    ml, eps = rnd(B, 2, D, seed=25) * 0.5, rnd(B, D, seed=26)
    z = ops.reparam_fwd(ml, eps)
    md = ml.double().requires_grad_()
    ref = md[:, 0] + torch.exp(md[:, 1] / 2) * eps.double()
    assert rel(z, ref) < 1e-6 and torch.equal(ops.reparam_fwd(ml, None), ml[:, 0])
    dz = rnd(B, D, seed=27)
    (g,) = torch.autograd.grad(ref, md, dz.double())
    assert rel(ops.reparam_bwd(dz, ml, eps), g) < 1e-6
    x = rnd(B, D, seed=28) * 2
    assert torch.equal(ops.hardtanh_fwd(x, 0.0, 1.0), x.clamp(0, 1))
    assert torch.equal(ops.hardtanh_bwd(dz, x, 0.0, 1.0), dz * ((x > 0) & (x < 1)))
    assert torch.equal(ops.mul(x, dz), x * dz) and torch.equal(ops.add(x, dz), x + dz)
    assert rel(ops.colsum(x), x.double().sum(0)) < 1e-6


def test_gemm_f32_all_transposes_and_split_k():
    ops.set_precision('fp32')
    for (m, n, k) in [(37, 50, 19), (160, 1220, 4096), (160, 300, 305)]:
        a, b, bias, res = rnd(m, k, seed=29), rnd(n, k, seed=30), rnd(n, seed=31), rnd(m, n, seed=32)
        y = ops.linear_fwd(a, b, bias, relu=False, residual=res)
        want = a.double() @ b.double().T + bias.double() + res.double()
        assert rel(y, want) < 2e-6
        dy = rnd(m, n, seed=33)
        assert rel(ops.linear_dgrad(dy, b), dy.double() @ b.double()) < 2e-6
        dw, db = ops.linear_wgrad(dy, a)
        assert rel(dw, dy.double().T @ a.double()) < 2e-6 and rel(db, dy.double().sum(0)) < 2e-6
    ops.set_precision('tf32')


@pytest.mark.parametrize("m,n,k", [(160, 1220, 24576), (160, 24576, 610), (160, 300, 305), (160, 610, 300), (5, 17, 9)])
def test_tensor_core_linear_layers(m, n, k):
    """nn.Linear forward / dgrad / wgrad on the tcgen05 gather kernel, incl. the unaligned K = 610 / 305 layers."""
    ops.set_precision('tf32')
    a, w, bias, res = rnd(m, k, seed=50), rnd(n, k, seed=51, scale=0.05), rnd(n, seed=52), rnd(m, n, seed=53)
    dy = rnd(m, n, seed=54)
    y = ops.linear_fwd(a, w, bias, relu=True, residual=res)
    want = torch.relu(a.double() @ w.double().T + bias.double() + res.double())
    dx = ops.linear_dgrad(dy, w)
    dw, db = ops.linear_wgrad(dy, a)
    errs = (rel(y, want), rel(dx, dy.double() @ w.double()), rel(dw, dy.double().T @ a.double()), rel(db, dy.double().sum(0)))
    print("tc linear", (m, n, k), "rel-L2 fwd %.2e dgrad %.2e wgrad %.2e db %.2e" % errs)
    assert max(errs) < 2e-3


def test_losses_match_oracle(idx_helper):
    from preset_gen_vae_b200 import synthetic
    B, D = 16, 610
    tables = ops.DeviceTables(idx_helper)
    v_in = synthetic.make_preset_targets(idx_helper, B, seed=3)
    v_out = torch.rand(B, D, generator=torch.Generator().manual_seed(4))
    vo = v_out.double().requires_grad_()
    want = oloss.synth_params_loss(idx_helper, vo, v_in.double())
    (gw,) = torch.autograd.grad(want, vo)
    got, ws = ops.synth_loss_fwd(v_out.to(DEV), v_in.to(DEV), tables, True, 0.2, True, 0.2)
    assert abs(got.item() - want.item()) < 1e-5 * abs(want.item())
    g = ops.synth_loss_bwd(torch.ones(1, device=DEV), v_out.to(DEV), v_in.to(DEV), tables, True, 0.2, True, 0.2, ws)
    assert rel(g.cpu(), gw) < 1e-5
    # un-normalised + no softmax inside the loss (params_reg_softmax=True configuration)
    q = torch.softmax(v_out.double() * 3, dim=1).clamp_min(1e-4).requires_grad_()
    want2 = oloss.synth_params_loss(idx_helper, q, v_in.double(), normalize_losses=False, cat_softmax=False)
    got2, ws2 = ops.synth_loss_fwd(q.detach().float().to(DEV), v_in.to(DEV), tables, False, 0.2, False, 0.2)
    assert abs(got2.item() - want2.item()) < 1e-4 * abs(want2.item())
    # squared error / latent / KL
    a, b = rnd(3, 1, 257, 347, seed=34), rnd(3, 1, 257, 347, seed=35)
    s = ops.sqerr_fwd(a, b, 1.0 / a.numel())
    assert abs(s.item() - F.mse_loss(a.double(), b.double()).item()) < 1e-6
    assert rel(ops.sqerr_bwd(a, b, 1.0 / a.numel(), torch.full((1,), 2.0, device=DEV)), 4 * (a - b).double() / a.numel()) < 1e-6
    from oracle import model as omodel
    ml, z0, zk, ld = rnd(B, 2, D, seed=36) * 0.3, rnd(B, D, seed=37), rnd(B, D, seed=38), rnd(B, seed=39)
    mld, z0d, zkd, ldd = (t.double().requires_grad_() for t in (ml, z0, zk, ld))
    fv = omodel.FlowVAE.__new__(omodel.FlowVAE)
    torch.nn.Module.__init__(fv)
    fv.normalize_latent_loss = True
    want = fv.latent_loss(mld, z0d, zkd, ldd)
    grads = torch.autograd.grad(want, (mld, z0d, zkd, ldd))
    got = ops.latent_loss_fwd(ml, z0, zk, ld, True)
    assert abs(got.item() - want.item()) < 1e-5 * abs(want.item()) + 1e-6
    for g_, w_ in zip(ops.latent_loss_bwd(torch.ones(1, device=DEV), ml, z0, zk, True), grads):
        assert rel(g_, w_) < 1e-5
    wantk = omodel.gaussian_dkl(mld[:, 0], mld[:, 1])
    (gk,) = torch.autograd.grad(wantk, mld)
    assert abs(ops.dkl_fwd(ml, True).item() - wantk.item()) < 1e-5 * abs(wantk.item())
    assert rel(ops.dkl_bwd(torch.ones(1, device=DEV), ml, True), gk) < 1e-5


def test_fused_adam_matches_torch():
    from preset_gen_vae_b200 import _lib
    n = 100_003
    p0, g = rnd(n, seed=40), rnd(n, seed=41) * 0.1
    ref_p = p0.clone().requires_grad_()
    opt = torch.optim.Adam([ref_p], lr=2e-4, weight_decay=1e-4, betas=(0.9, 0.999))
    p, m, v = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    L = _lib.lib()
    for step in range(1, 4):
        ref_p.grad = g.clone()
        opt.step()
        _lib.check(L.pgv_adam_step(_lib.ptr(p), _lib.ptr(g), _lib.ptr(m), _lib.ptr(v), n, 2e-4, 0.9, 0.999, 1e-8, 1e-4, step, 1.0,
                                   _lib.stream_ptr()))
    assert rel(p, ref_p.detach()) < 1e-6
    assert rel(m, opt.state[ref_p]['exp_avg']) < 1e-5 and rel(v, opt.state[ref_p]['exp_avg_sq']) < 5e-5
