"""Device time of the small (flow-conditioner) kernels: each op is captured 40x into a CUDA graph and replayed, so the
numbers are free of Python / launch overhead.  Usage: gpu_bench_small.py [B]"""
import sys
from types import SimpleNamespace

import torch

sys.path.insert(0, '.')
from preset_gen_vae_b200.model import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 160
dev = 'cuda'
REPS = 40


def graph_time(fn):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REPS):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * REPS) * 1e3       # us


def bn(n):
    return SimpleNamespace(weight=torch.ones(n, device=dev), bias=torch.zeros(n, device=dev), running_mean=torch.zeros(n, device=dev),
                           running_var=torch.ones(n, device=dev), momentum=0.1, eps=1e-3)


for (N, K) in [(300, 300), (300, 305), (610, 300)]:
    x = torch.randn(B, K, device=dev); w = torch.randn(N, K, device=dev) * 0.05; b = torch.randn(N, device=dev)
    dy = torch.randn(B, N, device=dev); res = torch.randn(B, N, device=dev); mask = torch.ones(B, N, device=dev)
    bnn, bnk = bn(N), bn(K)
    xk = torch.randn(B, K, device=dev)
    row = {}
    for cs in (True, False):
        ops.use_colslice = cs
        tag = 'colslice' if cs else 'small-gemm'
        row[tag + ' fwd'] = graph_time(lambda: ops.linear_fwd(x, w, b, residual=res))
        row[tag + ' dgrad'] = graph_time(lambda: ops.linear_dgrad(dy, w))
    ops.use_colslice = True
    row['wgrad+db'] = graph_time(lambda: ops.linear_wgrad(dy, x))
    row['bn1d fwd'] = graph_time(lambda: ops.bn1d_train_fwd(dy, bnn, relu=True, mask=mask))
    mean, rstd = torch.zeros(N, device=dev), torch.ones(N, device=dev)
    row['bn1d bwd'] = graph_time(lambda: ops.bn1d_train_bwd(dy, res, bnn, mean, rstd, relu=True, mask=mask))
    row['fused linear+bn fwd'] = graph_time(lambda: ops.linear_bn_fwd(x, w, b, bnn, residual=res, mask=mask))
    meank, rstdk = torch.zeros(K, device=dev), torch.ones(K, device=dev)
    row['fused dgrad+bn bwd'] = graph_time(lambda: ops.linear_dgrad_bn_bwd(dy, w, xk, bnk, meank, rstdk, add_post=xk))
    row['add'] = graph_time(lambda: ops.add(dy, res))
    print('N=%d K=%d B=%d (us per launch):' % (N, K, B), '  '.join('%s %.1f' % kv for kv in row.items()))
