"""Cost of accumulating the BatchNorm statistics in the convolution epilogue vs. BatchNorm's own reduction pass (graph-replayed)."""
import sys

import torch

sys.path.insert(0, '.')
from preset_gen_vae_b200.model import ops  # noqa: E402

ENC = [('enc2', 8, 16, 4, 129, 174), ('enc3', 16, 32, 4, 65, 88), ('enc4', 32, 64, 4, 33, 45), ('enc5', 64, 128, 4, 17, 23),
       ('enc6', 128, 256, 4, 9, 12), ('enc7', 256, 512, 4, 5, 7)]
DEC = [('dec2', 256, 512, 4, 5, 7), ('dec3', 128, 256, 4, 9, 12), ('dec4', 64, 128, 4, 17, 23), ('dec5', 32, 64, 4, 33, 45),
       ('dec6', 16, 32, 4, 65, 88), ('dec7', 8, 16, 4, 129, 174)]


def timeit(fn, reps=10):
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


B = 160
ops.set_precision('tf32')
dev = 'cuda'
print("layer   conv ms   conv+sums ms   bn(reduce+apply) ms   bn(apply) ms    saved ms")
tot = 0.0
for name, cin, cout, k, H, W in ENC + DEC:
    Ho, Wo = (H + 4 - k) // 2 + 1, (W + 4 - k) // 2 + 1
    w = torch.randn(cout, cin, k, k, device=dev) * 0.1
    wf, wq = ops.prep_conv_weights(w, 2, 2)
    bn_c = cout if name.startswith('enc') else cin
    bn = torch.nn.BatchNorm2d(bn_c).to(dev)
    if name.startswith('enc'):
        x = ops.to_cl(torch.randn(B, cin, H, W, device=dev), True)
        b = torch.randn(cout, device=dev)
        f0 = lambda: ops.conv2d_fwd(x, w, b, 2, 2, 0.1, wf=wf)
        f1 = lambda: ops.conv2d_fwd(x, w, b, 2, 2, 0.1, wf=wf, bn_sums=True)
    else:
        x = ops.to_cl(torch.randn(B, cout, Ho, Wo, device=dev), True)
        b = torch.randn(cin, device=dev)
        f0 = lambda: ops.conv2d_dgrad(x, w, (H, W), 2, 2, bias=b, slope=0.1, wq=wq)
        f1 = lambda: ops.conv2d_dgrad(x, w, (H, W), 2, 2, bias=b, slope=0.1, wq=wq, bn_sums=True)
    y, sums = f1()
    t0, t1 = timeit(f0), timeit(f1)
    b0, b1 = timeit(lambda: ops.bn2d_train_fwd(y, bn)), timeit(lambda: ops.bn2d_train_fwd(y, bn, sums))
    saved = (t0 + b0) - (t1 + b1)
    tot += saved
    print("%-6s %8.3f %12.3f %16.3f %16.3f %12.3f" % (name, t0, t1, b0, b1, saved))
print("total saved per step: %.3f ms" % tot)
