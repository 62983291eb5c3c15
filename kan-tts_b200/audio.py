"""``kantts.utils.audio_torch`` surface (MelSpectrogram, stft) on the fused STFT kernel.

MelSpectrogram mirrors kantts/utils/audio_torch.py:86-186: same kwargs, a ``melmat`` buffer of
shape (n_fft//2+1, num_mels) in the state_dict, ``forward((B,T) | (B,1,T)) -> (B, num_mels, frames)``.
The Slaney filterbank (librosa.filters.mel, :125-131) is computed here at construction time (host,
float64 -> float32) because librosa is only used for that one table.
"""
import math

import numpy as np
import torch

from . import ops


def slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) with its defaults (htk=False,
    norm='slaney'): triangular filters on the Slaney mel scale, area-normalised."""
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, math.log(6.4) / 27.0

    def hz_to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, f / f_sp)

    def mel_to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)

    n_bins = 1 + n_fft // 2
    fftfreqs = np.linspace(0, float(sr) / 2, n_bins, endpoint=True)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    weights = np.maximum(0, np.minimum(lower, upper))
    weights *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return weights.astype(np.float32)


def _padded_window(win_length, n_fft, device):
    w = torch.hann_window(win_length, dtype=torch.float32, device=device)   # periodic, as torch.stft callers pass
    if win_length < n_fft:
        left = (n_fft - win_length) // 2
        w = torch.nn.functional.pad(w, (left, n_fft - win_length - left))
    return w.contiguous()


class MelSpectrogram(torch.nn.Module):
    def __init__(self, fs=22050, fft_size=1024, hop_size=256, win_length=None, window="hann", num_mels=80,
                 fmin=80, fmax=7600, center=True, normalized=False, onesided=True, eps=1e-10, log_base=10.0,
                 pad_mode="constant"):
        super().__init__()
        if window != "hann":
            raise ValueError(f"{window} window is not implemented")
        if not center or normalized or not onesided or pad_mode != "constant":
            raise NotImplementedError("kantts_b200.MelSpectrogram: only center=True, normalized=False, "
                                      "onesided=True, pad_mode='constant' (the reference defaults)")
        if log_base not in (None, 2.0, 10.0):
            raise ValueError(f"log_base: {log_base} is not supported.")
        self.fft_size = fft_size
        self.win_length = fft_size if win_length is None else win_length
        self.hop_size = hop_size
        self.center, self.normalized, self.onesided = center, normalized, onesided
        self.window, self.eps, self.pad_mode, self.log_base = window, eps, pad_mode, log_base
        fmin = 0 if fmin is None else fmin
        fmax = fs / 2 if fmax is None else fmax
        melmat = slaney_mel_filterbank(fs, fft_size, num_mels, fmin, fmax)
        self.register_buffer("melmat", torch.from_numpy(melmat.T.copy()).float())
        self._win = None

    def _window(self, device):
        if self._win is None or self._win.device != device:
            self._win = _padded_window(self.win_length, self.fft_size, device)
        return self._win

    def forward(self, x):
        if x.dim() == 3:
            x = x.reshape(-1, x.size(2))
        return ops.StftMelFn.apply(x, self._window(x.device), self.melmat.contiguous(), self.fft_size,
                                   self.hop_size, 0, float(self.eps))


def stft(x, fft_size, hop_size, win_length, window):
    """audio_torch.py:8-31: magnitude spectrogram (B, frames, fft_size//2+1); ``window`` is the
    window tensor of length win_length (reflect padding, eps 1e-7)."""
    w = window.to(torch.float32)
    if w.numel() < fft_size:
        left = (fft_size - w.numel()) // 2
        w = torch.nn.functional.pad(w, (left, fft_size - w.numel() - left))
    return ops.StftMelFn.apply(x, w.contiguous(), None, fft_size, hop_size, 1, 1e-7)


def melspectrogram(y, sample_rate, n_fft=1024, hop_length=256, win_length=1024, n_mels=80, max_norm=1.0,
                   min_level_db=-100, ref_level_db=20, fmin=50, fmax=8000, symmetric=False, preemphasize=False,
                   device="cuda"):
    """The OFFLINE feature-extraction mel of kantts/preprocess/audio_processor/core/dsp.py:165-201 on the fused
    STFT kernel: librosa.stft framing (centre, reflect padding, periodic hann), |D| (no clamp), Slaney mel
    projection, 20*log10(max(1e-5, .)) - ref_level_db, dsp._normalize ([0, max_norm], or symmetric).
    y: 1-D waveform (numpy / tensor) -> numpy array (frames, n_mels) like the reference."""
    assert fmax <= sample_rate // 2
    yt = torch.as_tensor(np.asarray(y) if not torch.is_tensor(y) else y, dtype=torch.float32).reshape(1, -1)
    if preemphasize:                                              # dsp.py:53-56: lfilter([1, -0.98], [1], wav)
        yt = torch.cat([yt[:, :1], yt[:, 1:] - 0.98 * yt[:, :-1]], dim=1)
    yt = yt.to(device)
    melmat = torch.from_numpy(slaney_mel_filterbank(sample_rate, n_fft, n_mels, fmin, fmax).T.copy()).to(device)
    if symmetric:
        norm = (float(ref_level_db), float(min_level_db), 2.0 * max_norm, float(max_norm), -float(max_norm), float(max_norm))
    else:
        norm = (float(ref_level_db), float(min_level_db), float(max_norm), 0.0, 0.0, float(max_norm))
    mel = ops.StftMelFn.apply(yt, _padded_window(win_length, n_fft, yt.device), melmat.contiguous(), n_fft, hop_length,
                              1, 0.0, norm)
    return mel[0].transpose(0, 1).contiguous().cpu().numpy()
