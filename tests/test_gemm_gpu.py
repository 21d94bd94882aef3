"""GPU: the tcgen05 GEMM skeleton through the C ABI, against an fp64 product of the same operands."""
import ctypes

import pytest
import torch

from preset_gen_vae_b200 import _lib

pytestmark = pytest.mark.gpu


def tf32_trunc(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)


def run_tc(a, b, bias=None, act=0, three_pass=False):
    L, h = _lib.lib(), _lib.handle()
    m, k = a.shape
    n = b.shape[0]
    c = torch.full((m, n), float('nan'), device='cuda')
    a_lo = b_lo = None
    if three_pass:
        a_hi, a_lo, b_hi, b_lo = (torch.empty_like(t) for t in (a, a, b, b))
        _lib.check(L.pgv_split_tf32(_lib.ptr(a), _lib.ptr(a_hi), _lib.ptr(a_lo), a.numel(), _lib.stream_ptr()))
        _lib.check(L.pgv_split_tf32(_lib.ptr(b), _lib.ptr(b_hi), _lib.ptr(b_lo), b.numel(), _lib.stream_ptr()))
        a, b = a_hi, b_hi
    _lib.check(L.pgv_gemm_nt_tf32(h, _lib.ptr(a), _lib.ptr(a_lo), a.stride(0), _lib.ptr(b), _lib.ptr(b_lo), b.stride(0),
                                  _lib.ptr(c), n, m, n, k, _lib.ptr(bias), act, int(three_pass), _lib.stream_ptr()), 'gemm')
    torch.cuda.synchronize()
    return c


def make(m, n, k, seed=0):
    g = torch.Generator(device='cuda').manual_seed(seed)
    ld = (k + 3) // 4 * 4
    a = torch.randn(m, ld, device='cuda', generator=g)[:, :k]
    b = torch.randn(n, ld, device='cuda', generator=g)[:, :k]
    a, b = torch.as_strided(a, (m, k), (ld, 1)), torch.as_strided(b, (n, k), (ld, 1))
    return a, b


class _Strided:
    pass


def _ptr_strided(t):
    return ctypes.c_void_p(t.data_ptr())


SHAPES = [(128, 128, 32), (128, 128, 256), (256, 256, 64), (160, 300, 305), (1000, 610, 1220), (4, 24576, 610), (37, 50, 8)]


@pytest.mark.parametrize("m,n,k", SHAPES)
def test_tc_gemm_one_pass(m, n, k):
    """1xTF32: the tensor core reads the top 19 bits of each fp32 operand, so the exact expectation is the fp64
    product of the truncated operands; fp32 accumulation error only."""
    ld = (k + 3) // 4 * 4
    g = torch.Generator(device='cuda').manual_seed(1)
    A = torch.randn(m, ld, device='cuda', generator=g)
    B = torch.randn(n, ld, device='cuda', generator=g)
    A[:, k:] = float('nan')   # padding columns must never be read: TMA zero-fills beyond k
    B[:, k:] = float('nan')
    bias = torch.randn(n, device='cuda', generator=g)
    L, h = _lib.lib(), _lib.handle()
    c = torch.full((m, n), float('nan'), device='cuda')
    _lib.check(L.pgv_gemm_nt_tf32(h, _ptr_strided(A), None, ld, _ptr_strided(B), None, ld, _lib.ptr(c), n, m, n, k,
                                  _lib.ptr(bias), 1, 0, _lib.stream_ptr()), 'gemm')
    torch.cuda.synchronize()
    want = torch.relu(tf32_trunc(A[:, :k].contiguous()).double() @ tf32_trunc(B[:, :k].contiguous()).double().T + bias.double())
    err = (c.double() - want).abs().max().item()
    scale = want.abs().max().item()
    rna = torch.relu(A[:, :k].double() @ B[:, :k].double().T + bias.double())
    print("shape", (m, n, k), "max err vs truncated-operand product", err, "scale", scale,
          "| vs exact fp32-operand product", (c.double() - rna).abs().max().item())
    assert torch.isfinite(c).all()
    assert err <= 2e-5 * scale + 1e-4


@pytest.mark.parametrize("m,n,k", [(128, 128, 64), (160, 300, 305), (347, 1024, 1024)])
def test_tc_gemm_three_pass(m, n, k):
    """3xTF32 (lo*hi + hi*lo + hi*hi): fp32-equivalent products; the residual is the round-toward-zero update of the
    fp32 TMEM accumulator (k/8 sequential adds), hence the k-dependent bound."""
    a, b = (t.contiguous() for t in make(m, n, k, seed=2))
    if k % 4:
        pytest.skip("contiguous operands need k % 4 == 0") if False else None
    ld = (k + 3) // 4 * 4
    A = torch.zeros(m, ld, device='cuda'); A[:, :k] = a
    B = torch.zeros(n, ld, device='cuda'); B[:, :k] = b
    c = run_tc(A, B, three_pass=True)
    want = a.double() @ b.double().T
    err = (c.double() - want).abs().max().item()
    scale = want.abs().max().item()
    print("3-pass", (m, n, k), "max err", err, "scale", scale)
    assert err <= (1e-6 + 6e-8 * (k / 8)) * scale


def test_simt_gemm_matches():
    L, h = _lib.lib(), _lib.handle()
    a, b = (t.contiguous() for t in make(100, 70, 52, seed=3))
    bias = torch.randn(70, device='cuda')
    c = torch.empty(100, 70, device='cuda')
    _lib.check(L.pgv_gemm_nt_f32(h, _lib.ptr(a), 52, _lib.ptr(b), 52, _lib.ptr(c), 70, 100, 70, 52, _lib.ptr(bias), 0,
                                 _lib.stream_ptr()))
    torch.cuda.synchronize()
    want = a.double() @ b.double().T + bias.double()
    assert (c.double() - want).abs().max().item() < 1e-4


def test_bad_arguments_raise():
    L, h = _lib.lib(), _lib.handle()
    a = torch.zeros(4, 6, device='cuda')
    with pytest.raises(_lib.PgvError):
        _lib.check(L.pgv_gemm_nt_tf32(h, _lib.ptr(a), None, 6, _lib.ptr(a), None, 6, _lib.ptr(a), 4, 4, 4, 6, None, 0, 0,
                                      _lib.stream_ptr()))   # lda not a multiple of 4
