"""GPU: the column-slice GEMM family of the flow conditioners (Linear with BatchNorm1d / ReLU / Dropout fused into its
epilogue, Linear data-gradient with the BatchNorm backward fused) against fp64 PyTorch, and the fused ResidualNet
against the unfused kernel sequence."""
import pytest
import torch
import torch.nn.functional as F

from preset_gen_vae_b200.model import flows, ops

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return torch.randn(*shape, device=DEV, generator=g) * scale


@pytest.mark.parametrize("m,n,k", [(160, 300, 305), (160, 610, 300), (160, 300, 300), (7, 33, 19), (256, 300, 300), (1, 16, 4), (1024, 300, 300)])
def test_colslice_linear_forward_and_data_gradient(m, n, k):
    ops.set_precision('tf32')           # small layers run in exact fp32 whatever the precision mode
    x, w, b, res = rnd(m, k, seed=1), rnd(n, k, seed=2, scale=0.1), rnd(n, seed=3), rnd(m, n, seed=4)
    assert not ops._use_tc(m, n, k)          # conditioner-sized weights stay off the tensor cores whatever the batch (inference: 1024 rows)
    y = ops.linear_fwd(x, w, b, relu=True, residual=res)
    assert rel(y, torch.relu(x.double() @ w.double().T + b.double() + res.double())) < 2e-6
    assert rel(ops.linear_fwd(x, w, None), x.double() @ w.double().T) < 2e-6
    dy = rnd(m, n, seed=5)
    assert rel(ops.linear_dgrad(dy, w), dy.double() @ w.double()) < 2e-6


@pytest.mark.parametrize("m,n,k,use_mask,use_res", [(160, 300, 305, False, False), (160, 300, 300, True, False),
                                                    (160, 300, 300, False, True), (37, 50, 21, True, True), (256, 304, 300, True, True)])
def test_linear_batchnorm_fused_forward_backward(m, n, k, use_mask, use_res):
    x, w, b = rnd(m, k, seed=1), rnd(n, k, seed=2, scale=0.1), rnd(n, seed=3)
    res = rnd(m, n, seed=4) if use_res else None
    mask = (torch.rand(m, n, device=DEV, generator=torch.Generator(device=DEV).manual_seed(5)) > 0.3).float() / 0.7 if use_mask else None
    bn = torch.nn.BatchNorm1d(n, eps=1e-3).to(DEV)
    with torch.no_grad():
        bn.weight.copy_(rnd(n, seed=6) * 0.3 + 1); bn.bias.copy_(rnd(n, seed=7) * 0.2)
    ref_bn = torch.nn.BatchNorm1d(n, eps=1e-3).to(DEV).double()
    ref_bn.load_state_dict({kk: v.double() if v.is_floating_point() else v for kk, v in bn.state_dict().items()})
    y_pre, t, mean, rstd = ops.linear_bn_fwd(x, w, b, bn, residual=res, mask=mask)
    xd, wd = x.double().requires_grad_(), w.double().requires_grad_()
    yd = xd @ wd.T + b.double() + (res.double() if use_res else 0.0)
    td = torch.relu(ref_bn(yd)) * (mask.double() if use_mask else 1.0)
    assert rel(y_pre, yd) < 2e-6 and rel(t, td) < 5e-6
    assert rel(mean, yd.mean(0)) < 2e-6 and rel(rstd, 1.0 / torch.sqrt(yd.var(0, unbiased=False) + 1e-3)) < 2e-6
    assert rel(bn.running_mean, ref_bn.running_mean) < 1e-6 and rel(bn.running_var, ref_bn.running_var) < 1e-6
    # backward of mask * relu(BN(u)) fused behind the data gradient dt = dy2 @ w2 of the NEXT Linear (n -> n2 features)
    n2 = 40
    w2, dy2, post = rnd(n2, n, seed=8, scale=0.1), rnd(m, n2, seed=9), rnd(m, n, seed=10)
    u = yd.detach().requires_grad_()
    t2 = torch.relu(ref_bn.train()(u)) * (mask.double() if use_mask else 1.0)
    out = t2 @ w2.double().T
    gu, gg, gb = torch.autograd.grad(out, (u, ref_bn.weight, ref_bn.bias), dy2.double())
    du, dg, db = ops.linear_dgrad_bn_bwd(dy2, w2, y_pre, bn, mean, rstd, mask=mask, add_post=post)
    assert rel(du, gu + post.double()) < 2e-5 and rel(dg, gg) < 2e-5 and rel(db, gb) < 2e-5


def test_fused_residual_net_matches_unfused_sequence():
    torch.manual_seed(3)
    net = flows.ResidualNet(305, 610, 300, num_blocks=2, dropout_probability=0.25, use_batch_norm=True).to(DEV)
    B = 160
    x = rnd(B, 305, seed=1)
    masks = [(torch.rand(B, 300, device=DEV) > 0.25).float() / 0.75 for _ in range(2)]
    dout = rnd(B, 610, seed=2)
    state = {k: v.clone() for k, v in net.state_dict().items()}
    res = {}
    for fused in (True, False):
        ops.use_colslice = fused
        net.load_state_dict(state)
        try:
            out, ctx = net.fwd(x, True, masks)
            grads = {}
            dx = net.bwd(dout, ctx, grads)
        finally:
            ops.use_colslice = True
        assert ctx[3] == fused
        res[fused] = (out, dx, [grads[id(p)] for p in net.params()], {k: v.clone() for k, v in net.state_dict().items() if 'running' in k})
    assert rel(res[True][0], res[False][0]) < 1e-5 and rel(res[True][1], res[False][1]) < 1e-4
    for a, b in zip(res[True][2], res[False][2]):       # biases that feed a BatchNorm have analytically zero gradients (1e-8 noise)
        assert float((a - b).norm()) <= 1e-4 * float(b.norm()) + 1e-6
    for k in res[True][3]:
        assert rel(res[True][3][k], res[False][3][k]) < 1e-5


def test_flow_program_kernel_equals_the_separate_launches():
    """A whole RealNVP forward + backward recorded into ONE persistent launch (pgv_flow_program: the same tile code behind grid barriers)
    against the chain of stand-alone kernels: outputs, log-determinants and every parameter gradient, for the latent flow (couplings
    only) and the regression flow (dropout masks + flow BatchNorm transforms, which flush the recording)."""
    from preset_gen_vae_b200.model import flows
    torch.manual_seed(3)
    for reg in (False, True):
        for B in (160, 5):
            if reg:
                flow = flows.CustomRealNVP(610, 300, 6, 2, dropout_probability=0.4, batch_norm_within_layers=True, batch_norm_between_layers=True).cuda().train()
            else:
                flow = flows.SimpleRealNVP(610, 300, 6, 2, batch_norm_within_layers=True)._transform.cuda().train()
            for prm in flow.parameters():                       # the conditioners' last layers are initialised near zero: make them matter
                if prm.dim() == 2:
                    prm.data.mul_(3.0)
            x = torch.randn(B, 610, device='cuda')
            masks = flow._draw_masks(x)
            gy, gld = torch.randn(B, 610, device='cuda'), torch.randn(B, device='cuda')
            res = {}
            for on in (True, False):
                ops.use_flow_program = on
                try:
                    flow.zero_grad()
                    xi = x.clone().requires_grad_()
                    before = ops.launches
                    y, ld = flow(xi, dropout_masks=masks)
                    n_fwd = ops.launches - before
                    ((y * gy).sum() + (ld * gld).sum()).backward()
                    torch.cuda.synchronize()
                    res[on] = (y.detach(), ld.detach(), xi.grad.clone(), [p.grad.clone() for p in flow.parameters()], n_fwd)
                finally:
                    ops.use_flow_program = True
            a, b = res[True], res[False]
            assert a[4] < b[4] / 2, (a[4], b[4])                  # 2 launches instead of 48 (latent flow); the flow BatchNorms flush the recording
            for u, v in zip(a[:3], b[:3]):
                assert float((u - v).abs().max()) <= 2e-6 * float(v.abs().max()) + 1e-7
            # the program kernel's weight-gradient tile adds the batch rows in a different order; a bias in front of a BatchNorm has a
            # mathematically zero gradient (rounding noise ~1e-4 of the weight gradients' scale), hence the common absolute term
            scale = max(float(v.abs().max()) for v in b[3])
            for u, v in zip(a[3], b[3]):
                assert float((u - v).abs().max()) <= 1e-5 * float(v.abs().max()) + 1e-5 * scale
