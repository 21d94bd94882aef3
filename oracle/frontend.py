"""CPU oracle of the spectrogram front end.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows the reference line by line:
  utils/audio.py:24-31   Spectrogram.__init__   symmetric Hann, norm = max|rfft(window)|
  utils/audio.py:33-40   get_stft               torch.stft(center=True, pad_mode='constant', onesided)
  utils/audio.py:42-54   __call__ / linear_to_log_scale   |X|/norm -> 20*log10(max(., 10^(min_dB/20)))
  utils/audio.py:80-87   MelSpectrogram.__call__  mel_basis @ (|X|/norm) then dB (magnitude, norm=None)
  data/abstractbasedataset.py:129-131   min-max normalisation to [-1, 1]

PARITY UNPINNED for `slaney_mel_filterbank`: it restates librosa 0.8 `filters.mel(sr=22050, n_fft, n_mels,
fmin=0, fmax=sr/2, htk=False, norm=None)` (Slaney auditory-toolbox scale) because librosa is not installed;
tests cross-check it against torchaudio.functional.melscale_fbanks(mel_scale='slaney', norm=None).
"""
import numpy as np
import torch


def hann_window(n_fft: int, dtype=torch.float32) -> torch.Tensor:
    return torch.hann_window(n_fft, periodic=False, dtype=dtype)          # audio.py:30


def norm_factor(n_fft: int) -> float:
    return torch.fft.rfft(hann_window(n_fft)).abs().max().item()          # audio.py:31  (= 511.5 for 1024)


def stft(x_wav, n_fft: int, fft_hop: int, dtype=torch.float32) -> torch.Tensor:
    """Complex STFT [..., n_fft/2+1, 1 + L//hop]; audio.py:33-40."""
    x = torch.as_tensor(x_wav).to(dtype)
    return torch.stft(x, n_fft=n_fft, hop_length=fft_hop, window=hann_window(n_fft, dtype), center=True,
                      pad_mode='constant', onesided=True, return_complex=True)


def linear_to_log_scale(spectrogram: torch.Tensor, min_dB: float) -> torch.Tensor:
    floor = torch.ones_like(spectrogram) * 10 ** (min_dB / 20.0)          # audio.py:53
    return 20.0 * torch.log10(torch.maximum(spectrogram, floor))          # audio.py:54


def magnitude(x_wav, n_fft: int, fft_hop: int, dtype=torch.float32) -> torch.Tensor:
    return stft(x_wav, n_fft, fft_hop, dtype).abs() / norm_factor(n_fft)  # audio.py:44-46


def spectrogram_db(x_wav, n_fft: int, fft_hop: int, min_dB: float, dtype=torch.float32) -> torch.Tensor:
    return linear_to_log_scale(magnitude(x_wav, n_fft, fft_hop, dtype), min_dB)


def _hz_to_mel_slaney(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def _mel_to_hz_slaney(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def slaney_mel_filterbank(n_fft: int, n_mels: int, sr: float = 22050.0, fmin: float = 0.0, fmax=None) -> np.ndarray:
    """float32 [n_mels, n_fft/2+1] triangular filters, un-normalised (norm=None)."""
    if fmax is None:
        fmax = sr / 2.0
    fftfreqs = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    mel_f = _mel_to_hz_slaney(np.linspace(_hz_to_mel_slaney(fmin), _hz_to_mel_slaney(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    return weights


def mel_spectrogram_db(x_wav, n_fft: int, fft_hop: int, min_dB: float, n_mel_bins: int,
                       dtype=torch.float32) -> torch.Tensor:
    """audio.py:80-87.  librosa's default sr=22050 sets fmax (self.Fs is not forwarded)."""
    mag = magnitude(x_wav, n_fft, fft_hop, dtype)
    basis = torch.from_numpy(slaney_mel_filterbank(n_fft, n_mel_bins)).to(dtype)
    return linear_to_log_scale(torch.matmul(basis, mag), min_dB)


def min_max_normalize(spec_db: torch.Tensor, spec_min: float, spec_max: float) -> torch.Tensor:
    return -1.0 + (spec_db - spec_min) / ((spec_max - spec_min) / 2.0)    # abstractbasedataset.py:129-131


def batch_front_end(audio: torch.Tensor, n_fft=1024, fft_hop=256, min_dB=-120.0, n_mel_bins=257, spec_min=None,
                    spec_max=None, dtype=torch.float32) -> torch.Tensor:
    """[B, C, L] audio -> [B, C, F, T]; one clip at a time, exactly as the reference DataLoader worker does
    (abstractbasedataset.py:124-134).  n_mel_bins <= 0 selects the linear-frequency Spectrogram."""
    B, C, _ = audio.shape
    out = []
    for b in range(B):
        chans = []
        for c in range(C):
            if n_mel_bins > 0:
                s = mel_spectrogram_db(audio[b, c], n_fft, fft_hop, min_dB, n_mel_bins, dtype)
            else:
                s = spectrogram_db(audio[b, c], n_fft, fft_hop, min_dB, dtype)
            if spec_min is not None:
                s = min_max_normalize(s, spec_min, spec_max)
            chans.append(s)
        out.append(torch.stack(chans))
    return torch.stack(out)
