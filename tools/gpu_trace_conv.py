"""Pipeline trace of CTA 0 of the tensor-core conv kernel (clock64 deltas per k-block)."""
import sys
import torch
sys.path.insert(0, '.')
from preset_gen_vae_b200 import _lib
from preset_gen_vae_b200.model import ops
ops.set_precision('tf32')
B = 160
dev = 'cuda'
L = _lib.lib()
cases = {
    'dec7 dgrad (N=8,K=64)': lambda: ops.conv2d_dgrad(torch.randn(B, 16, 65, 88, device=dev), torch.randn(16, 8, 4, 4, device=dev), (129, 174), 2, 2),
    'enc4 fwd (N=64,K=512)': lambda: ops.conv2d_fwd(torch.randn(B, 32, 33, 45, device=dev), torch.randn(64, 32, 4, 4, device=dev), torch.randn(64, device=dev), 2, 2, 0.1),
    'enc7 fwd (N=512,K=4096)': lambda: ops.conv2d_fwd(torch.randn(B, 256, 5, 7, device=dev), torch.randn(512, 256, 4, 4, device=dev), torch.randn(512, device=dev), 2, 2, 0.1),
}
for name, fn in cases.items():
    fn(); torch.cuda.synchronize()
    tr = torch.zeros(3 * 64 * 8, dtype=torch.int64, device=dev)
    L.pgv_debug_set_conv_trace(_lib.ptr(tr))
    fn(); torch.cuda.synchronize()
    L.pgv_debug_set_conv_trace(None)
    t = tr.view(3, 64, 8).cpu()
    t0 = int(t[0, 0, 0])
    print("==== %s   (cycles relative to producer's first event)" % name)
    print(" producer thread 0: kb | loads issued | waited empty | stored | fenced | arrived   (deltas within the k-block) | start")
    for i in range(14):
        a, b, c, d, e, f, kb = (int(v) for v in t[0, i, :7])
        if a == 0: break
        print("   kb %3d  load %5d  wait_empty %5d  store %5d  fence %5d  arrive %5d | start %7d" % (kb, b - a, c - b, d - c, e - d, f - e, a - t0))
    print(" mma thread: kb | wait_full | issue+commit | start")
    for i in range(14):
        a, b, c = (int(v) for v in t[1, i, :3]); kb = int(t[1, i, 6])
        if a == 0: break
        print("   kb %3d  wait_full %6d  issue %5d | start %7d" % (kb, b - a, c - b, a - t0))
    print(" epilogue thread: tile | wait_tfull | drain | start")
    for i in range(8):
        a, b, c = (int(v) for v in t[2, i, :3])
        if a == 0: break
        print("   tile %2d  wait %6d  drain %6d | start %7d" % (i, b - a, c - b, a - t0))
