"""Deterministic synthetic inputs shared by the tests, the oracle and bench.py.

The reference needs 10.5 GB of pre-rendered Dexed wavs and an LFS preset DB that are not in its tree
(data/dexeddataset.py:371, synth/dexed_presets.sqlite is an LFS pointer), so every measurement and parity
test here runs on seeded synthetic tensors of the reference's shapes.  Everything is generated on the CPU
with an explicit torch.Generator, so the same seed gives the same tensors on every box.
"""
import math

import numpy as np
import torch

SAMPLE_RATE = 22050
# 173 RenderMan buffers of 512 samples (synth/dexed.py:223) => 1 + 88576 // 256 = 347 STFT frames, the
# spectrogram_size of config.py:46.  Exactly 4.0 s (88200 samples) would give 345 frames.
CLIP_SAMPLES = 88576
# Dataset-wide min/max used by the min-max normalisation of abstractbasedataset.py:129-131.  The reference's
# stats file is not in its tree; `min` is the dB floor (config.py:42), `max` is a documented constant just
# above the largest mel-dB value the generator below produces.
SPEC_STATS = {'min': -120.0, 'max': 0.0}


def make_audio(batch: int, channels: int = 1, seed: int = 0, n_samples: int = CLIP_SAMPLES) -> torch.Tensor:
    """[batch, channels, n_samples] float32: 1-8 exponentially decaying harmonics of a log-uniform f0 in
    [55, 1760] Hz, peak amplitude in [0.05, 0.5], N(0, 1e-4^2) noise, 0.1 s linear fade-out (dexed.py:252-255)."""
    g = torch.Generator().manual_seed(seed)
    n = batch * channels
    t = torch.arange(n_samples, dtype=torch.float64) / SAMPLE_RATE
    f0 = 55.0 * torch.pow(2.0, 5.0 * torch.rand(n, generator=g, dtype=torch.float64))
    amp = 0.05 + 0.45 * torch.rand(n, generator=g, dtype=torch.float64)
    n_harm = torch.randint(1, 9, (n,), generator=g)
    decay = 0.5 + 5.5 * torch.rand(n, 8, generator=g, dtype=torch.float64)
    phase = 2.0 * math.pi * torch.rand(n, 8, generator=g, dtype=torch.float64)
    x = torch.zeros(n, n_samples, dtype=torch.float64)
    for h in range(1, 9):
        fh = f0 * h
        on = ((n_harm >= h) & (fh < 0.45 * SAMPLE_RATE)).to(torch.float64)
        a = (amp / h * on)[:, None]
        x += a * torch.exp(-decay[:, h - 1, None] * t[None, :]) \
            * torch.sin(2.0 * math.pi * fh[:, None] * t[None, :] + phase[:, h - 1, None])
    x += 1e-4 * torch.randn(n, n_samples, generator=g, dtype=torch.float64)
    fade = min(int(0.1 * SAMPLE_RATE), n_samples)
    x[:, -fade:] *= torch.linspace(1.0, 0.0, fade, dtype=torch.float64)[None, :]
    return x.to(torch.float32).reshape(batch, channels, n_samples)


def make_preset_targets(idx_helper, batch: int, seed: int = 0, p_silent_operator: float = 0.15) -> torch.Tensor:
    """v_in [batch, learnable_preset_size] float32: numerical columns U(0,1), each Dexed operator output level
    forced to 0 with probability `p_silent_operator` (exercises data/preset.py:264-281), categorical groups one-hot."""
    g = torch.Generator().manual_seed(seed + 7919)
    v = torch.zeros(batch, idx_helper.learnable_preset_size, dtype=torch.float32)
    num_cols = idx_helper.get_numerical_learnable_indexes()
    v[:, num_cols] = torch.rand(batch, len(num_cols), generator=g)
    for op in range(6):
        vol = idx_helper.full_to_learnable[31 + 22 * op] if idx_helper.full_preset_size > 31 + 22 * op else None
        if isinstance(vol, int):
            silent = torch.rand(batch, generator=g) < p_silent_operator
            v[silent, vol] = 0.0
    for cols in idx_helper.get_categorical_learnable_indexes():
        cls = torch.randint(0, len(cols), (batch,), generator=g)
        v[torch.arange(batch), cols[0] + cls] = 1.0
    return v


def make_sample_info(batch: int, pitch: int = 60, velocity: int = 85) -> torch.Tensor:
    """[batch, 3] int32 = (preset UID, MIDI pitch, MIDI velocity), abstractbasedataset.py:142-144."""
    info = torch.empty(batch, 3, dtype=torch.int32)
    info[:, 0] = torch.arange(batch, dtype=torch.int32)
    info[:, 1] = pitch
    info[:, 2] = velocity
    return info


def make_spectrogram_like(batch: int, channels: int = 1, seed: int = 0, size=(257, 347)) -> torch.Tensor:
    """Cheap stand-in for a min-max-normalised mel-dB batch in [-1, 1] (smooth in time and frequency), for model
    tests that do not want to pay for the front end."""
    g = torch.Generator().manual_seed(seed + 104729)
    coarse = torch.rand(batch, channels, 17, 23, generator=g) * 2.0 - 1.0
    x = torch.nn.functional.interpolate(coarse, size=size, mode='bilinear', align_corners=True)
    x = x + 0.05 * torch.randn(batch, channels, *size, generator=g)
    return x.clamp_(-1.0, 1.0).contiguous()


def make_noise(batch: int, dim_z: int, fc_dropout: float, reg_fc_dropout: float, flow_hidden: int = 300,
               flow_layers: int = 6, flow_blocks: int = 2, enc_fc_in: int = 24576, dec_fc_out: int = 24576,
               seed: int = 1):
    """Every random tensor one training-mode forward consumes, in the order the reference draws them
    (SURVEY.md §7 'RNG parity'): encoder FC dropout mask (encoder.py:85), eps (VAE.py:172-173), decoder FC
    dropout mask (decoder.py:65), regression-flow conditioner dropout masks for couplings 0..L-3
    (flows.py:75,81).  Masks are already scaled by 1/(1-p), i.e. out = in * mask."""
    g = torch.Generator().manual_seed(seed)

    def mask(shape, p):
        if p <= 0.0:
            return torch.ones(shape)
        return torch.empty(shape).bernoulli_(1.0 - p, generator=g) / (1.0 - p)
    noise = {'enc_fc_mask': mask((batch, enc_fc_in), fc_dropout)}
    noise['eps'] = torch.randn(batch, dim_z, generator=g)
    noise['dec_fc_mask'] = mask((batch, dec_fc_out), fc_dropout)
    noise['reg_masks'] = [[mask((batch, flow_hidden), reg_fc_dropout if layer < flow_layers - 2 else 0.0)
                           for _ in range(flow_blocks)] for layer in range(flow_layers)]
    return noise
