"""Pipeline trace of CTA 0 of the channels-last tensor-core conv kernel (clock64 timestamps per k-block / tile)."""
import sys
import torch
sys.path.insert(0, '.')
from preset_gen_vae_b200 import _lib
from preset_gen_vae_b200.model import ops
ops.set_precision('tf32')
B = 160
dev = 'cuda'
L = _lib.lib()


def fwd_case(cin, cout, H, W):
    x = ops.to_cl(torch.randn(B, cin, H, W, device=dev), True); w = torch.randn(cout, cin, 4, 4, device=dev) * 0.1; b = torch.randn(cout, device=dev)
    wf, _ = ops.prep_conv_weights(w, 2, 2)
    return lambda: ops.conv2d_fwd(x, w, b, 2, 2, 0.1, wf=wf)


def quad_case(cin, cout, H, W):
    """Transposed conv (data gradient of conv cin -> cout on an H x W input): dy [B, cout, Ho, Wo] -> [B, cin, H, W]."""
    Ho, Wo = H // 2 + 1, W // 2 + 1
    dy = ops.to_cl(torch.randn(B, cout, Ho, Wo, device=dev), True); w = torch.randn(cout, cin, 4, 4, device=dev) * 0.1; b = torch.randn(cin, device=dev)
    _, wq = ops.prep_conv_weights(w, 2, 2)
    return lambda: ops.conv2d_dgrad(dy, w, (H, W), 2, 2, bias=b, slope=0.1, wq=wq)


cases = {'enc2 fwd (N=16, 4 k-blocks per tile)': fwd_case(8, 16, 129, 174), 'dec7 fwd (quad, N=32, 2 k-blocks per tile)': quad_case(8, 16, 129, 174),
         'enc3 fwd (N=32, 8 k-blocks)': fwd_case(16, 32, 65, 88), 'enc5 fwd (N=128, 32 k-blocks)': fwd_case(64, 128, 17, 23),
         'enc7 fwd (N=128, 128 k-blocks)': fwd_case(256, 512, 5, 7)}
if 'a0' in sys.argv:
    L.pgv_debug_set_conv_a_mode(0)
if 'noa' in sys.argv:           # no A loads at all: what the tile period would be if the activation operand were free
    L.pgv_debug_set_conv_a_mode(-2)
for name, fn in cases.items():
    fn(); torch.cuda.synchronize()
    tr = torch.zeros(3 * 64 * 8, dtype=torch.int64, device=dev)
    L.pgv_debug_set_conv_trace(_lib.ptr(tr))
    fn(); torch.cuda.synchronize()
    L.pgv_debug_set_conv_trace(None)
    t = tr.view(3, 64, 8).cpu()
    t0 = int(t[0, 0, 0])
    print("==== %s   (cycles; start relative to the producer's first event)" % name)
    for role, label in ((0, 'producer thread 0 (cp.async) / TMA warp: wait_empty / issue'), (1, 'mma thread: wait_full / issue+commit'), (2, 'epilogue: wait_tfull / drain')):
        print(' ' + label)
        prev = None
        for i in list(range(0, 10)) + list(range(24, 34)):
            a, b, c = (int(v) for v in t[role, i, :3]); tag = int(t[role, i, 6])
            if a == 0:
                break
            extra = ''
            if role == 1:
                d, e = int(t[role, i, 3]), int(t[role, i, 4])
                extra = '  [fence+test %d  mma+commit %d  syncwarp %d]' % (d - b, e - d, c - e)
            if role == 2:       # only asm-volatile-bracketed intervals are trustworthy: the compiler moves plain loads / stores across clock reads
                extra = '  [tcgen05.ld + wait %d]' % (int(t[role, i, 3]) - b)
            print("   #%2d tag %4d  wait %6d  work %5d | start %8d  period %s%s" % (i, tag, b - a, c - b, a - t0, '' if prev is None else a - prev, extra))
            prev = a
