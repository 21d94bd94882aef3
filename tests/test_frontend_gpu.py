"""GPU: the CUDA front end (through the C ABI) against the CPU oracle and the reference's golden vectors.

Tolerance (north_star): spectrograms within 1e-4 relative.  Stated as |d| <= 1e-4*|ref| + atol on the dB output,
checked against the fp64 oracle (the fp32 reference itself is 9e-5 relative / 0.011 dB away from fp64 on the mel
output and 3e-4 relative / 0.036 dB on the linear-frequency output), with atol = 2e-3 dB for the mel
path and 2.5e-2 dB for the linear path (measured worst case 0.022 dB, below the reference's own 0.036 dB): bins just above the -120 dB floor are pure cancellation noise, there the error
is set by the absolute accuracy of re/im (fp32 accumulation), not by the relative accuracy of the products.
"""
import os

import numpy as np
import pytest
import torch

from oracle import frontend as ofe
from preset_gen_vae_b200 import synthetic
from preset_gen_vae_b200.utils.audio import MelSpectrogram, Spectrogram

pytestmark = pytest.mark.gpu


def report(name, got, ref):
    d = (got.double() - ref.double()).abs()
    rel = d / ref.double().abs().clamp_min(1e-3)
    print("%s: max|d|=%.3e dB  max rel=%.3e  p99.9 rel=%.3e  mean|d|=%.3e" %
          (name, d.max().item(), rel.max().item(), torch.quantile(rel.flatten()[:4_000_000], 0.999).item(), d.mean().item()))
    return d, rel


def test_mel_db_parity_vs_fp64_oracle_and_golden(golden_dir):
    audio = synthetic.make_audio(4, 1, seed=0)
    mel = MelSpectrogram(1024, 256, -120.0, 257, 22050)
    got = mel(audio.cuda()).cpu()[:, 0]
    assert got.shape == (4, 257, 347)
    ref64 = ofe.mel_spectrogram_db(audio[:, 0], 1024, 256, -120.0, 257, dtype=torch.float64)
    d, rel = report("mel dB vs fp64 oracle", got, ref64)
    assert torch.all(d <= 1e-4 * ref64.abs() + 2e-3)
    gold = torch.from_numpy(np.load(os.path.join(golden_dir, 'frontend.npz'))['mel_db'])   # reference fp32 output
    d, rel = report("mel dB vs reference golden (fp32)", got, gold)
    assert torch.all(d <= 2e-4 * gold.abs() + 5e-3)    # two fp32-rounded results: twice the one-sided budget


def test_linear_db_parity(golden_dir):
    audio = synthetic.make_audio(4, 1, seed=0)[:2]     # the golden fixture holds clip 0 of the 4-clip batch
    spec = Spectrogram(1024, 256, -120.0)
    got = spec(audio.cuda()).cpu()[:, 0]
    assert got.shape == (2, 513, 347)
    ref64 = ofe.spectrogram_db(audio[:, 0], 1024, 256, -120.0, dtype=torch.float64)
    d, rel = report("linear dB vs fp64 oracle", got, ref64)
    assert torch.all(d <= 1e-4 * ref64.abs() + 2.5e-2)
    assert d.mean().item() < 5e-4
    gold = torch.from_numpy(np.load(os.path.join(golden_dir, 'frontend.npz'))['lin_db_clip0'])
    d, rel = report("linear dB clip0 vs reference golden", got[0], gold)
    assert torch.quantile(d.flatten(), 0.999).item() < 2e-2 and d.max().item() < 0.1


def test_known_answers_on_device():
    spec = Spectrogram(1024, 256, -120.0)
    silent = spec(torch.zeros(1, 88576, device='cuda'))
    assert torch.all(silent == -120.0)                                     # audio.py:53 floor
    n = torch.arange(88576, dtype=torch.float64)
    sine = torch.sin(2 * np.pi * 64 * n / 1024).float()
    out = spec(sine.cuda())
    assert out.shape == (513, 347)
    assert abs(out[64, 100].item() - 20 * np.log10(0.5)) < 1e-3           # -6.02 dB at the bin centre
    nyq = torch.cos(np.pi * n).float()                                     # alternating +-1: all energy in bin 512
    out = spec(nyq.cuda())
    assert abs(out[512, 100].item()) < 1e-3 and out[256, 100].item() < -100
    dc = torch.ones(88576)
    out = spec(dc.cuda())
    assert abs(out[0, 100].item()) < 1e-3                                  # DC: |sum(w)|/511.5 = 1 -> 0 dB


def test_host_call_shapes_and_linearity():
    """Reference calling convention: 1-D numpy waveform in, CPU [F, T] tensor out; plus a size-independent property
    (magnitude is homogeneous: scaling the audio by 10 adds 20 dB above the floor)."""
    audio = synthetic.make_audio(3, 1, seed=5)[:, 0]
    mel = MelSpectrogram(1024, 256, -120.0, 257, 22050)
    one = mel(audio[0].numpy())
    assert isinstance(one, torch.Tensor) and not one.is_cuda and one.shape == (257, 347)
    batch = mel(audio)
    assert batch.shape == (3, 257, 347) and torch.equal(batch[0], one)
    loud = mel(audio * 10.0)
    above = batch > -90.0
    assert torch.allclose(loud[above], batch[above] + 20.0, atol=2e-3)


def test_ragged_and_other_geometries():
    """Lengths that are not whole hops, the 345-frame (exactly 4.0 s) case, a short clip, and n_fft=512."""
    for length, n_fft, hop in [(88200, 1024, 256), (88576 + 77, 1024, 256), (1024, 1024, 256), (20000, 512, 256), (5000, 2048, 512)]:
        x = synthetic.make_audio(2, 1, seed=9, n_samples=length)[:, 0]
        spec = Spectrogram(n_fft, hop, -120.0)
        got = spec(x.cuda()).cpu()
        ref = ofe.spectrogram_db(x, n_fft, hop, -120.0, dtype=torch.float64)
        assert got.shape == ref.shape == (2, n_fft // 2 + 1, 1 + length // hop)
        d = (got.double() - ref).abs()
        print("geometry", (length, n_fft, hop), "max|d| dB", d.max().item())
        assert torch.all(d <= 1e-4 * ref.abs() + 2.5e-2)


def test_fused_min_max_normalisation_and_linear_output():
    audio = synthetic.make_audio(2, 1, seed=2)[:, 0].cuda()
    mel = MelSpectrogram(1024, 256, -120.0, 257, 22050)
    db = mel(audio)
    nrm = mel(audio, normalize=(-120.0, 0.0))
    assert torch.allclose(nrm, ofe.min_max_normalize(db, -120.0, 0.0), atol=1e-5)
    assert float(nrm.min()) >= -1.0 - 1e-6
    lin = Spectrogram(1024, 256, -120.0, log_scale=False)(audio).cpu()
    ref = ofe.magnitude(audio.cpu(), 1024, 256, dtype=torch.float64)
    d = (lin.double() - ref).abs()
    print('linear magnitude: max abs err %.3e (floor amplitude is 1e-6)' % d.max().item())
    assert torch.all(d <= 1e-5 * ref + 5e-7)     # half the -120 dB floor amplitude


def test_full_size_batch_properties():
    """BASELINE configs[1] size (256 clips): checked through size-independent properties — every clip of the big batch
    equals the same clip computed alone, and silent clips hit the floor exactly."""
    audio = synthetic.make_audio(256, 1, seed=0)[:, 0]
    audio[17] = 0.0
    mel = MelSpectrogram(1024, 256, -120.0, 257, 22050)
    big = mel(audio.cuda())
    assert big.shape == (256, 257, 347) and torch.isfinite(big).all()
    assert torch.all(big[17] == -120.0)
    for i in (0, 100, 255):
        assert torch.equal(mel(audio[i].cuda()), big[i])
