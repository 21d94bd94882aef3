"""CPU: the front-end oracle against the reference's golden vectors and known-answer tests."""
import os

import numpy as np
import torch

from oracle import frontend as ofe
from preset_gen_vae_b200 import synthetic
from preset_gen_vae_b200.utils.audio import slaney_mel_filterbank


def test_oracle_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'frontend.npz'))
    audio = synthetic.make_audio(4, 1, seed=0)
    lin = ofe.spectrogram_db(audio[0, 0], 1024, 256, -120.0)
    assert lin.shape == (513, 347)
    # same torch.stft call as the reference; fixture generated on another host => allow FFT-kernel rounding only
    assert np.abs(lin.numpy() - g['lin_db_clip0']).max() < 5e-3
    close = np.isclose(lin.numpy(), g['lin_db_clip0'], rtol=1e-4, atol=1e-3)
    assert close.mean() > 0.9999
    mel = ofe.batch_front_end(audio, n_mel_bins=257)[:, 0]
    assert mel.shape == (4, 257, 347)
    assert np.abs(mel.numpy() - g['mel_db']).max() < 5e-3
    assert abs(float(g['norm_factor']) - 511.5) < 1e-3


def test_known_answers():
    assert abs(ofe.norm_factor(1024) - 511.5) < 1e-3                      # Hann sum for n_fft=1024
    assert ofe.stft(torch.zeros(88576), 1024, 256).shape == (513, 347)    # 88576 samples <=> 347 frames
    assert ofe.stft(torch.zeros(88200), 1024, 256).shape == (513, 345)    # exactly 4.0 s gives 345
    silent = ofe.spectrogram_db(torch.zeros(88576), 1024, 256, -120.0)
    assert torch.all(silent == -120.0)                                    # floor everywhere
    n = torch.arange(88576, dtype=torch.float64)
    sine = torch.sin(2 * np.pi * 64 * n / 1024).float()
    peak = ofe.spectrogram_db(sine, 1024, 256, -120.0)[64, 100].item()
    assert abs(peak - 20 * np.log10(0.5)) < 1e-2                          # unit sine at a bin centre: -6.02 dB


def test_mel_filterbank_cross_checks():
    ours = ofe.slaney_mel_filterbank(1024, 257)
    assert ours.shape == (257, 513) and ours.dtype == np.float32
    assert (ours > 0).sum() == 1016 and (ours.sum(axis=1) > 0).all()       # 1016 non-zeros, no empty filter
    import torchaudio
    ta = torchaudio.functional.melscale_fbanks(513, 0.0, 11025.0, 257, 22050, norm=None, mel_scale='slaney').T.numpy()
    assert np.abs(ours - ta).max() < 2e-5                                  # independent Slaney implementation
    prod = slaney_mel_filterbank(1024, 257)                                # the product's own construction
    assert np.abs(ours - prod).max() < 1e-6


def test_min_max_normalisation():
    s = torch.tensor([-120.0, -60.0, 0.0])
    assert torch.allclose(ofe.min_max_normalize(s, -120.0, 0.0), torch.tensor([-1.0, 0.0, 1.0]))


def test_fp32_oracle_vs_fp64():
    """The reference's own fp32 rounding noise on the mel dB output: this is the floor of any 1e-4 gate."""
    audio = synthetic.make_audio(2, 1, seed=3)[:, 0]
    m32 = ofe.mel_spectrogram_db(audio, 1024, 256, -120.0, 257)
    m64 = ofe.mel_spectrogram_db(audio, 1024, 256, -120.0, 257, dtype=torch.float64)
    err = (m32.double() - m64).abs()
    assert err.max() < 0.05 and (err / m64.abs()).max() < 5e-4
