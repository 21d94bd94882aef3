"""Per-layer device time of the tensor-core conv / linear kernels at the training batch (each op replayed from a CUDA graph)."""
import sys

import torch

sys.path.insert(0, '.')
from preset_gen_vae_b200.model import ops  # noqa: E402

B = next((int(a) for a in sys.argv[1:] if a.isdigit()), 160)
dev = 'cuda'
ops.set_precision('tf32')
ops.use_cl = not (len(sys.argv) > 2 and sys.argv[2] == 'nchw')
if 'a0' in sys.argv:               # A operand by cp.async gathers instead of TMA (A/B)
    from preset_gen_vae_b200 import _lib
    _lib.check(_lib.lib().pgv_debug_set_conv_a_mode(0))
if 'noa' in sys.argv:              # measurement aid: no A loads at all (results are garbage)
    from preset_gen_vae_b200 import _lib
    _lib.check(_lib.lib().pgv_debug_set_conv_a_mode(-2))
if 'atomic' in sys.argv:           # round-1 split-K: fp32 atomics instead of the workspace
    ops.deterministic = False
print('route:', 'channels-last' if ops.use_cl else 'NCHW register-staged', '| A operand:', 'cp.async' if 'a0' in sys.argv else 'TMA where possible',
      '| split-K:', 'atomics' if 'atomic' in sys.argv else 'workspace (deterministic)')
# conv geometry (Cin, Cout, k, s, p, H, W, Ho, Wo) of the convolution whose fwd/dgrad/wgrad each layer uses
ENC = [('enc1', 1, 8, 5, 257, 347), ('enc2', 8, 16, 4, 129, 174), ('enc3', 16, 32, 4, 65, 88), ('enc4', 32, 64, 4, 33, 45),
       ('enc5', 64, 128, 4, 17, 23), ('enc6', 128, 256, 4, 9, 12), ('enc7', 256, 512, 4, 5, 7)]
DEC = [('dec2', 256, 512, 4, 5, 7), ('dec3', 128, 256, 4, 9, 12), ('dec4', 64, 128, 4, 17, 23), ('dec5', 32, 64, 4, 33, 45),
       ('dec6', 16, 32, 4, 65, 88), ('dec7', 8, 16, 4, 129, 174), ('dec8', 1, 8, 5, 257, 347)]


def timeit(fn, reps=10):
    """Device time per call: the op is captured `reps` times into a CUDA graph (no Python / launch overhead in the timed region)."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * reps)


for a in sys.argv:
    if a.startswith('groups='):
        from preset_gen_vae_b200 import _lib as _l
        _l.lib().pgv_debug_set_conv_groups(int(a.split('=')[1]))
tot = {'fwd': 0.0, 'dgrad': 0.0, 'wgrad': 0.0}
print("layer        flops(G)  fwd ms (TF/s)      dgrad ms (TF/s)    wgrad ms (TF/s)   act MB")
for name, cin, cout, k, H, W in ENC + DEC:
    s, p = 2, 2
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    x = torch.randn(B, cin, H, W, device=dev)
    w = torch.randn(cout, cin, k, k, device=dev) * 0.1
    b = torch.randn(cout, device=dev)
    dy = torch.randn(B, cout, Ho, Wo, device=dev)
    fl = 2 * B * Ho * Wo * cout * cin * k * k / 1e9
    wf = wq = None
    if ops.conv_route(cin, cout, k, k, s, p, H, W, Ho, Wo) == 'cl':       # operands as the model holds them: channels-last, pre-rounded
        x, dy = ops.to_cl(x, True), ops.to_cl(dy, True)
        wf, wq = ops.prep_conv_weights(w, s, p)
    elif ops.cl_mode() and cin == 1:
        dy = ops.to_cl(dy, True)
    t_f = timeit(lambda: ops.conv2d_fwd(x, w, b, s, p, 0.1, wf=wf))
    t_d = timeit(lambda: ops.conv2d_dgrad(dy, w, (H, W), s, p, wq=wq))
    t_w = timeit(lambda: ops.conv2d_wgrad(x, dy, w.shape, s, p, want_bias=False))
    tot['fwd'] += t_f; tot['dgrad'] += t_d; tot['wgrad'] += t_w
    mb = 4 * (x.numel() + dy.numel()) / 1e6
    print("%-10s %8.2f   %7.3f (%6.1f)   %7.3f (%6.1f)   %7.3f (%6.1f)   %7.1f" %
          (name, fl, t_f, fl / t_f, t_d, fl / t_d, t_w, fl / t_w, mb))
for name, cin, cout, H, W in [('enc8 1x1', 512, 2048, 3, 4), ('dec1 1x1', 512, 2048, 3, 4)]:
    x = torch.randn(B, cin, H, W, device=dev); w = torch.randn(cout, cin, 1, 1, device=dev) * 0.05; b = torch.randn(cout, device=dev)
    dy = torch.randn(B, cout, H, W, device=dev)
    fl = 2 * B * H * W * cout * cin / 1e9
    wf = wq = None
    if ops.conv_route(cin, cout, 1, 1, 1, 0, H, W, H, W) == 'cl':
        x, dy = ops.to_cl(x, True), ops.to_cl(dy, True)
        wf, wq = ops.prep_conv_weights(w, 1, 0)
    t_f = timeit(lambda: ops.conv2d_fwd(x, w, b, 1, 0, 0.1, wf=wf)); t_d = timeit(lambda: ops.conv2d_dgrad(dy, w, (H, W), 1, 0, wq=wq))
    t_w = timeit(lambda: ops.conv2d_wgrad(x, dy, w.shape, 1, 0, want_bias=False))
    tot['fwd'] += t_f; tot['dgrad'] += t_d; tot['wgrad'] += t_w
    print("%-10s %8.2f   %7.3f (%6.1f)   %7.3f (%6.1f)   %7.3f (%6.1f)" % (name, fl, t_f, fl / t_f, t_d, fl / t_d, t_w, fl / t_w))
for name, M, N, K in [('enc FC', B, 1220, 24576), ('dec FC', B, 24576, 610)]:        # big Linear layers: ops.fc_fwd / fc_bwd (incl. operand copies)
    a = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * 0.05; bias = torch.randn(N, device=dev); dy = torch.randn(M, N, device=dev)
    fl = 2 * M * N * K / 1e9
    y, ctx = ops.fc_fwd(a, w, bias, True)
    t_f = timeit(lambda: ops.fc_fwd(a, w, bias, True)); t_b = timeit(lambda: ops.fc_bwd(dy, ctx, w, True))
    print("%-14s %6.2f   fwd incl. rounded copies %7.3f ms (%6.1f TF/s)   dgrad + wgrad + db %7.3f ms (%6.1f TF/s)   [route %s]" %
          (name, fl, t_f, fl / t_f, t_b, 2 * fl / t_b, ops.fc_route(M, N, K)))
for name, M, N, K in [('flow 300x300', B, 300, 300), ('flow 305->300', B, 300, 305), ('flow 300->610', B, 610, 300)]:
    a = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * 0.05; bias = torch.randn(N, device=dev); dy = torch.randn(M, N, device=dev)
    fl = 2 * M * N * K / 1e9
    t_f = timeit(lambda: ops.linear_fwd(a, w, bias)); t_d = timeit(lambda: ops.linear_dgrad(dy, w)); t_w = timeit(lambda: ops.linear_wgrad(dy, a))
    print("%-14s %6.2f   %7.3f (%6.1f)   %7.3f (%6.1f)   %7.3f (%6.1f)" % (name, fl, t_f, fl / t_f, t_d, fl / t_d, t_w, fl / t_w))
print("conv totals (ms): fwd %.2f dgrad %.2f wgrad %.2f" % (tot['fwd'], tot['dgrad'], tot['wgrad']))
