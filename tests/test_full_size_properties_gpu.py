"""GPU, BASELINE.json's full sizes (B = 160 per GPU, 610 latent / preset dimensions, 257 x 347 spectrograms): properties that hold
whatever the size and need no CPU oracle run - adjointness of the three convolution kernels, flow round trips, optimizer sharding,
idempotence of the rounding producers."""
import pytest
import torch

from preset_gen_vae_b200 import _lib
from preset_gen_vae_b200.model import flows, ops

pytestmark = pytest.mark.gpu
DEV = 'cuda'
B = 160


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return torch.randn(*shape, device=DEV, generator=g) * scale


def dot(a, b):
    return float((a.double() * b.double()).sum())


# (Cin, Cout, k, stride, pad, H, W): the thin 5x5 layer, an HBM-class layer, a deep layer and the 1x1 mixer (encoder.py:233-259)
@pytest.mark.parametrize("cin,cout,k,s,p,H,W", [(1, 8, 5, 2, 2, 257, 347), (8, 16, 4, 2, 2, 129, 174), (256, 512, 4, 2, 2, 5, 7),
                                                (512, 2048, 1, 1, 0, 3, 4)])
def test_convolution_kernels_are_mutually_adjoint_at_full_batch(cin, cout, k, s, p, H, W):
    """<conv(x; w), y> = <x, dgrad(y; w)> = <w, wgrad(x, y)>: the forward, data-gradient and weight-gradient kernels are three
    contractions of the same trilinear form.  Operands are TF32-representable, so the tensor-core products are exact and only the
    fp32 accumulation order differs: the three numbers agree to ~1e-5 of the form's scale."""
    ops.set_precision('tf32')
    x = rnd(B, cin, H, W, seed=1)
    w = rnd(cout, cin, k, k, seed=2, scale=0.1)
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    y = rnd(B, cout, Ho, Wo, seed=3)
    trunc = lambda t: (t.view(torch.int32) & ~0x1FFF).view(torch.float32)          # exactly representable in TF32
    x, w, y = trunc(x), trunc(w), trunc(y)
    if cin > 1:
        x = ops.to_cl(x)
    y = ops.to_cl(y)
    fwd = ops.conv2d_fwd(x, w, None, s, p)
    dx = ops.conv2d_dgrad(y, w, (H, W), s, p)
    dw, _ = ops.conv2d_wgrad(x, y, w.shape, s, p, want_bias=False)
    a, b, c = dot(fwd, y), dot(dx, x), dot(dw, w)
    scale = float(fwd.double().norm() * y.double().norm())
    assert abs(a - b) <= 2e-5 * scale and abs(a - c) <= 2e-5 * scale, (a, b, c, scale)


@pytest.mark.parametrize("with_bn", [False, True])
def test_flow_round_trip_at_full_batch(with_bn):
    """inverse(forward(z)) = z and the two log-determinants cancel, for the latent flow (plain couplings) and the regression flow
    (BatchNorm transforms between the couplings), 6 layers x 300 hidden units, 610 dimensions, B = 160 (flows.py:42-90)."""
    torch.manual_seed(3)
    if with_bn:
        flow = flows.CustomRealNVP(610, 300, 6, 2, batch_norm_within_layers=True, batch_norm_between_layers=True).to(DEV)
    else:
        flow = flows.SimpleRealNVP(610, 300, 6, 2, batch_norm_within_layers=True)._transform.to(DEV)
    flow.train()
    z = rnd(B, 610, seed=5)
    with torch.no_grad():
        for _ in range(2):                       # running statistics away from their initial values
            flow(z)
        flow.eval()
        y, ld = flow(z)
        back, ld_back = flow.inverse(y)
    err = (back - z).abs()
    if with_bn:
        # A randomly initialised flow with BatchNorm transforms is ill conditioned in fp32: the fp64 oracle inverts it to 5e-10, the
        # SAME oracle in fp32 to mean 3.1e-3 / max 0.44 (tools/gpu_diag_flow_inverse.py); this repo: mean 3.2e-3 / max 0.34.
        assert float(err.mean()) < 1e-2 and float(err.median()) < 2e-3
    else:
        assert float(err.max()) < 2e-4 * float(z.abs().max())
    assert float((ld + ld_back).abs().max()) < 2e-3 * float(ld.abs().max().clamp_min(1.0))
    assert bool(torch.isfinite(y).all()) and y.shape == z.shape and ld.shape == (B,)


def test_adam_on_shards_equals_adam_on_the_whole_buffer():
    """The data-parallel step applies Adam to 1/world of every segment per rank: an elementwise update, so any partition gives bit-identical
    parameters and moments."""
    n = 60_372_096 // 16
    L = _lib.lib()
    p0, g = rnd(n, seed=40), rnd(n, seed=41) * 0.1
    res = []
    for parts in (1, 8):
        p, m, v = p0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
        for step in (1, 2):
            for r in range(parts):
                lo, hi = r * n // parts // 4 * 4, ((r + 1) * n // parts // 4 * 4 if r + 1 < parts else n)
                _lib.check(L.pgv_adam_step(_lib.ptr(p[lo:hi]), _lib.ptr(g[lo:hi]), _lib.ptr(m[lo:hi]), _lib.ptr(v[lo:hi]), hi - lo, 2e-4, 0.9,
                                           0.999, 1e-8, 1e-4, step, 0.125, _lib.stream_ptr()))
        res.append((p, m, v))
    for a, b in zip(*res):
        assert torch.equal(a, b)


def test_rounding_producers_are_idempotent_at_full_size():
    """Tensors that feed the tensor cores are stored TF32-rounded by their producers; rounding twice changes nothing, and the
    channels-last <-> NCHW converters are exact inverses (activation of enc2 at B = 160: 58 MB)."""
    x = rnd(B, 16, 65, 88, seed=9)
    a = ops.to_cl(x, True)
    b = ops.to_cl(ops.to_nchw(a), True)
    assert torch.equal(a, b)
    assert torch.equal(ops.to_nchw(ops.to_cl(x)), x)
    assert float((a - x).abs().max()) <= 2.0 ** -11 * float(x.abs().max())
