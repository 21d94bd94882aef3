"""Diagnostics for the first GPU bring-up of the tcgen05 GEMM skeleton and the front end (prints, never asserts)."""
import sys
import time
import traceback

import torch

sys.path.insert(0, '.')
from preset_gen_vae_b200 import _lib, synthetic  # noqa: E402


def trunc(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)


def gemm_case(m, n, k, three=False):
    L, h = _lib.lib(), _lib.handle()
    g = torch.Generator(device='cuda').manual_seed(0)
    a = torch.randn(m, k, device='cuda', generator=g)
    b = torch.randn(n, k, device='cuda', generator=g)
    c = torch.full((m, n), float('nan'), device='cuda')
    if three:
        ah, al, bh, bl = (torch.empty_like(t) for t in (a, a, b, b))
        L.pgv_split_tf32(_lib.ptr(a), _lib.ptr(ah), _lib.ptr(al), a.numel(), _lib.stream_ptr())
        L.pgv_split_tf32(_lib.ptr(b), _lib.ptr(bh), _lib.ptr(bl), b.numel(), _lib.stream_ptr())
        rc = L.pgv_gemm_nt_tf32(h, _lib.ptr(ah), _lib.ptr(al), k, _lib.ptr(bh), _lib.ptr(bl), k, _lib.ptr(c), n, m, n, k, None, 0, 1, _lib.stream_ptr())
    else:
        rc = L.pgv_gemm_nt_tf32(h, _lib.ptr(a), None, k, _lib.ptr(b), None, k, _lib.ptr(c), n, m, n, k, None, 0, 0, _lib.stream_ptr())
    torch.cuda.synchronize()
    want_t = trunc(a).double() @ trunc(b).double().T
    want = a.double() @ b.double().T
    nan = torch.isnan(c).sum().item()
    e_t = (c.double() - want_t).abs().max().item()
    e = (c.double() - want).abs().max().item()
    print("gemm %s m=%d n=%d k=%d rc=%d nan=%d  err(trunc ops)=%.3e  err(exact)=%.3e  scale=%.2f" %
          ('3x' if three else '1x', m, n, k, rc, nan, e_t, e, want.abs().max().item()), flush=True)
    if e_t > 1e-2 and m * n <= 128 * 128:
        bad = ((c.double() - want_t).abs() > 1e-2)
        rows = bad.any(dim=1).nonzero().flatten().tolist()
        cols = bad.any(dim=0).nonzero().flatten().tolist()
        print("   bad rows (first 16):", rows[:16], "count", len(rows), "| bad cols (first 16):", cols[:16], "count", len(cols))
        print("   c[0,:8]   ", c[0, :8].tolist())
        print("   want[0,:8]", want_t[0, :8].tolist())


def main():
    print(torch.cuda.get_device_name(0), "SMs", _lib.lib().pgv_sm_count(_lib.handle()), flush=True)
    for args in [(128, 128, 8), (128, 128, 32), (128, 128, 64), (128, 128, 512), (256, 128, 64), (128, 256, 64),
                 (1024, 1024, 1024), (160, 300, 308)]:
        try:
            gemm_case(*args)
        except Exception:
            traceback.print_exc()
    for args in [(128, 128, 64), (347, 1024, 1024)]:
        try:
            gemm_case(*args, three=True)
        except Exception:
            traceback.print_exc()
    from oracle import frontend as ofe
    from preset_gen_vae_b200.utils.audio import MelSpectrogram, Spectrogram
    audio = synthetic.make_audio(4, 1, seed=0)[:, 0]
    for name, obj, ref in [("linear", Spectrogram(1024, 256, -120.0), lambda x: ofe.spectrogram_db(x, 1024, 256, -120.0, dtype=torch.float64)),
                           ("mel", MelSpectrogram(1024, 256, -120.0, 257, 22050), lambda x: ofe.mel_spectrogram_db(x, 1024, 256, -120.0, 257, dtype=torch.float64))]:
        try:
            got = obj(audio.cuda()).cpu()
            r = ref(audio)
            d = (got.double() - r).abs()
            print("%s dB: shape %s max|d|=%.4e mean|d|=%.4e nan=%d  worst idx %s" %
                  (name, tuple(got.shape), d.max().item(), d.mean().item(), torch.isnan(got).sum().item(),
                   tuple(int(v) for v in torch.unravel_index(d.argmax(), d.shape))), flush=True)
            print("   got[0,:4,100]", got[0, :4, 100].tolist(), " ref", r[0, :4, 100].tolist())
            print("   got[0,-3:,100]", got[0, -3:, 100].tolist(), " ref", r[0, -3:, 100].tolist())
            print("   got[0,60,:3]", got[0, 60, :3].tolist(), " ref", r[0, 60, :3].tolist(), "| last frames", got[0, 60, -3:].tolist(), r[0, 60, -3:].tolist())
        except Exception:
            traceback.print_exc()
    # timing, B=256
    try:
        mel = MelSpectrogram(1024, 256, -120.0, 257, 22050)
        big = synthetic.make_audio(256, 1, seed=0)[:, 0].cuda()
        for _ in range(3):
            mel.compute(big)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            mel.compute(big)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("front end B=256: %.3f ms/batch = %.0f clips/s ; executed TF32 flops %.1f TFLOP/s" %
              (ms, 256 / ms * 1e3, 3 * 820.63e6 * 256 / ms / 1e9), flush=True)
    except Exception:
        traceback.print_exc()


if __name__ == '__main__':
    main()
