"""Spectrogram front end with the reference's `utils.audio` interface (utils/audio.py:20-92), computed on the B200.

`Spectrogram(n_fft, fft_hop, min_dB)(x_wav)` and `MelSpectrogram(n_fft, fft_hop, min_dB, n_mel_bins, Fs)(x_wav)`
keep the reference's constructor arguments, attributes (`window`, `spectrogram_norm_factor`, `n_fft`, `fft_hop`,
`min_dB`, `log_scale`, `n_mel_bins`, `Fs`) and return conventions: a 1-D waveform (numpy array or tensor, as the
dataset passes it, abstractbasedataset.py:126-128) gives a CPU float32 tensor `[F, T]`.  Two extensions serve the
batched GPU data path: an input of shape `[..., L]` is processed as a batch in one call, and a CUDA
tensor input returns a CUDA tensor without any host copy.  `normalize=(min, max)` fuses the dataset's min-max
scaling (abstractbasedataset.py:129-131) into the kernel epilogue.

All arithmetic runs in libpgv.so (pgv_frontend_fwd, include/pgv.h): windowed-DFT and mel projection as 3xTF32
tcgen05 contractions.  No torch.stft, no CPU path.
"""
import ctypes

import numpy as np
import torch

from .. import _lib


def slaney_mel_filterbank(n_fft, n_mels, sr=22050.0, fmin=0.0, fmax=None):
    """Triangular mel filters [n_mels, n_fft//2+1] on the Slaney scale, un-normalised.  This is what
    `librosa.feature.melspectrogram(S=..., n_mels=n_mels, norm=None)` builds internally at audio.py:85-86 (librosa's
    default sr=22050 and fmax=sr/2 apply because the reference does not forward `Fs`)."""
    fmax = sr / 2.0 if fmax is None else fmax
    f_sp, brk_hz, step = 200.0 / 3.0, 1000.0, np.log(6.4) / 27.0
    brk_mel = brk_hz / f_sp

    def to_mel(hz):
        hz = np.asarray(hz, dtype=np.float64)
        return np.where(hz < brk_hz, hz / f_sp, brk_mel + np.log(np.maximum(hz, brk_hz) / brk_hz) / step)

    def to_hz(mel):
        mel = np.asarray(mel, dtype=np.float64)
        return np.where(mel < brk_mel, mel * f_sp, brk_hz * np.exp(step * (mel - brk_mel)))
    edges = to_hz(np.linspace(to_mel(fmin), to_mel(fmax), n_mels + 2))            # [n_mels + 2] band edges in Hz
    freqs = np.linspace(0.0, sr / 2.0, n_fft // 2 + 1)[None, :]
    rising = (freqs - edges[:-2, None]) / (edges[1:-1] - edges[:-2])[:, None]
    falling = (edges[2:, None] - freqs) / (edges[2:] - edges[1:-1])[:, None]
    return np.maximum(0.0, np.minimum(rising, falling)).astype(np.float32)


class _DeviceState:
    """Constant operands and scratch for one device."""

    def __init__(self, device, window, n_fft, mel_basis):
        self.device = device
        self.handle = _lib.handle(device)
        self.basis_hi = torch.empty(n_fft, n_fft, dtype=torch.float32, device=device)
        self.basis_lo = torch.empty_like(self.basis_hi)
        n_mels = 0 if mel_basis is None else mel_basis.shape[0]
        self.mel_hi = self.mel_lo = None
        mel_ptr = ctypes.c_void_p(0)
        if n_mels:
            ld = _lib.lib().pgv_frontend_mel_ld(n_fft)
            self.mel_hi = torch.empty(n_mels, ld, dtype=torch.float32, device=device)
            self.mel_lo = torch.empty_like(self.mel_hi)
            mel_host = np.ascontiguousarray(mel_basis, dtype=np.float32)
            mel_ptr = ctypes.c_void_p(mel_host.ctypes.data)
        win = np.ascontiguousarray(window.numpy(), dtype=np.float32)
        with torch.cuda.device(device):
            _lib.check(_lib.lib().pgv_frontend_init_constants(
                self.handle, ctypes.c_void_p(win.ctypes.data), n_fft, _lib.ptr(self.basis_hi), _lib.ptr(self.basis_lo),
                mel_ptr, n_mels, _lib.ptr(self.mel_hi), _lib.ptr(self.mel_lo)), 'pgv_frontend_init_constants')
        self.workspace = None
        self.staging = None

    def scratch(self, nbytes):
        if self.workspace is None or self.workspace.numel() < nbytes:
            self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self.workspace


class Spectrogram:
    """dB (or linear) magnitude STFT spectrogram; interface of utils/audio.py:20-54."""

    def __init__(self, n_fft, fft_hop, min_dB, dynamic_range_dB=None, log_scale=True, device=None):
        self.n_fft = n_fft
        self.fft_hop = fft_hop
        self.log_scale = log_scale
        self.min_dB = min_dB
        self.dynamic_range_dB = dynamic_range_dB
        self.window = torch.hann_window(self.n_fft, periodic=False)                              # audio.py:30
        self.spectrogram_norm_factor = torch.fft.rfft(self.window).abs().max().item()            # audio.py:31
        self.device = device
        self._states = {}
        if _lib.lib().pgv_frontend_workspace_bytes(1, n_fft, n_fft, fft_hop, 0) == 0:
            raise NotImplementedError("front end supports n_fft = power of two in [256, 4096] and a hop that is a "
                                      "multiple of 32 dividing n_fft/2 (got n_fft={}, hop={})".format(n_fft, fft_hop))

    # mel subclasses override these two
    def _mel_basis(self):
        return None

    @property
    def n_output_bins(self):
        return self.n_fft // 2 + 1

    def _state(self, device):
        key = (device.type, device.index)
        if key not in self._states:
            self._states[key] = _DeviceState(device, self.window, self.n_fft, self._mel_basis())
        return self._states[key]

    def num_frames(self, n_samples):
        return 1 + n_samples // self.fft_hop

    def compute(self, audio_dev, normalize=None, out=None):
        """audio_dev: CUDA float32 [N, L] contiguous -> CUDA float32 [N, F, T].  The one device entry point."""
        assert audio_dev.is_cuda and audio_dev.dtype == torch.float32 and audio_dev.dim() == 2
        audio_dev = audio_dev.contiguous()
        n, length = audio_dev.shape
        st = self._state(audio_dev.device)
        n_mels = 0 if st.mel_hi is None else st.mel_hi.shape[0]
        frames = self.num_frames(length)
        if out is None:
            out = torch.empty(n, self.n_output_bins, frames, dtype=torch.float32, device=audio_dev.device)
        L = _lib.lib()
        ws_bytes = L.pgv_frontend_workspace_bytes(n, length, self.n_fft, self.fft_hop, n_mels)
        ws = st.scratch(ws_bytes)
        lo, hi = (0.0, 0.0) if normalize is None else normalize
        with torch.cuda.device(audio_dev.device):
            _lib.check(L.pgv_frontend_fwd(
                st.handle, _lib.ptr(audio_dev), n, length, self.n_fft, self.fft_hop, _lib.ptr(st.basis_hi),
                _lib.ptr(st.basis_lo), _lib.ptr(st.mel_hi), _lib.ptr(st.mel_lo), n_mels, float(self.min_dB),
                float(self.spectrogram_norm_factor), int(bool(self.log_scale)), int(normalize is not None), float(lo),
                float(hi), _lib.ptr(out), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(audio_dev.device)), 'pgv_frontend_fwd')
        return out

    def compute_host(self, audio_host, normalize=None, device=None):
        """Host float32 [N, L] (ideally pinned) -> host float32 [N, F, T] through pgv_frontend_fwd_host: H2D copy,
        kernels and D2H copy on the current stream, synchronised on return.  This is the end-to-end call bench.py times."""
        assert not audio_host.is_cuda and audio_host.dtype == torch.float32 and audio_host.dim() == 2
        audio_host = audio_host.contiguous()
        device = torch.device(device or self.device or 'cuda')
        if device.index is None:
            device = torch.device('cuda', torch.cuda.current_device())
        n, length = audio_host.shape
        st = self._state(device)
        n_mels = 0 if st.mel_hi is None else st.mel_hi.shape[0]
        frames = self.num_frames(length)
        if st.staging is None or st.staging[0].shape != (n, length):
            st.staging = (torch.empty(n, length, dtype=torch.float32, device=device),
                          torch.empty(n, self.n_output_bins, frames, dtype=torch.float32, device=device),
                          torch.empty(n, self.n_output_bins, frames, dtype=torch.float32).pin_memory())
        audio_dev, out_dev, out_host = st.staging
        L = _lib.lib()
        ws = st.scratch(L.pgv_frontend_workspace_bytes(n, length, self.n_fft, self.fft_hop, n_mels))
        lo, hi = (0.0, 0.0) if normalize is None else normalize
        with torch.cuda.device(device):
            _lib.check(L.pgv_frontend_fwd_host(
                st.handle, _lib.ptr(audio_host), _lib.ptr(audio_dev), n, length, self.n_fft, self.fft_hop,
                _lib.ptr(st.basis_hi), _lib.ptr(st.basis_lo), _lib.ptr(st.mel_hi), _lib.ptr(st.mel_lo), n_mels,
                float(self.min_dB), float(self.spectrogram_norm_factor), int(bool(self.log_scale)),
                int(normalize is not None), float(lo), float(hi), _lib.ptr(out_dev), _lib.ptr(out_host), _lib.ptr(ws),
                ws.numel(), _lib.stream_ptr(device)), 'pgv_frontend_fwd_host')
        return out_host

    def __call__(self, x_wav, normalize=None):
        """Reference call (audio.py:42-50).  x_wav: [L] or [..., L], numpy array / CPU tensor / CUDA tensor."""
        x = torch.as_tensor(x_wav)
        lead = x.shape[:-1]
        if x.is_cuda:
            y = self.compute(x.to(torch.float32).reshape(-1, x.shape[-1]), normalize)
            return y.reshape(*lead, *y.shape[-2:])
        y = self.compute_host(x.to(torch.float32).reshape(-1, x.shape[-1]), normalize)
        return y.reshape(*lead, *y.shape[-2:]).clone()

    # -- small elementwise helpers of the reference class (not on the hot path; plain tensor ops) --
    def linear_to_log_scale(self, spectrogram):                                                  # audio.py:52-54
        return 20.0 * torch.log10(torch.clamp_min(spectrogram, 10 ** (self.min_dB / 20.0)))

    def log_to_linear_scale(self, spectrogram):                                                  # audio.py:56-61
        return torch.pow(10.0, spectrogram / 20.0) * self.spectrogram_norm_factor


class MelSpectrogram(Spectrogram):
    """Mel-frequency dB spectrogram; interface of utils/audio.py:72-87."""

    def __init__(self, n_fft, fft_hop, min_dB, n_mel_bins, Fs, device=None):
        self.Fs = Fs
        self.n_mel_bins = n_mel_bins
        super().__init__(n_fft, fft_hop, min_dB, log_scale=True, device=device)

    def _mel_basis(self):
        return slaney_mel_filterbank(self.n_fft, self.n_mel_bins)

    @property
    def n_output_bins(self):
        return self.n_mel_bins


def build_spectrogram(model_config, device=None):
    """The dataset's choice between the two classes (data/dataset.py:18-25, abstractbasedataset.py:66-74): mel when
    `mel_bins` > 0, hard-coded 22050 Hz."""
    n_fft, hop = model_config.stft_args
    if model_config.mel_bins > 0:
        return MelSpectrogram(n_fft, hop, model_config.spectrogram_min_dB, model_config.mel_bins, 22050, device=device)
    return Spectrogram(n_fft, hop, model_config.spectrogram_min_dB, device=device)
