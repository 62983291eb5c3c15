"""Offline (feature-extraction) mel-spectrogram: CPU restatement of
kantts/preprocess/audio_processor/core/dsp.py:8-9,20-21,53-56,66-74,135-151,165-201 (TEST INFRASTRUCTURE).

librosa (stft, filters.mel) is not vendored in /root/reference nor installed: PARITY UNPINNED for this
function; librosa 0.9.2's published ``stft`` defaults are restated (center=True, pad_mode='reflect',
window='hann' = scipy.signal.get_window('hann', win_length, fftbins=True) i.e. periodic, centre-padded to
n_fft, dtype complex64) with numpy's rfft in float64."""
import numpy as np

from . import melbasis


def stft_abs(y, n_fft, hop_length, win_length):
    y = np.asarray(y, dtype=np.float64)
    n = np.arange(win_length)
    win = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)          # periodic hann
    if win_length < n_fft:
        left = (n_fft - win_length) // 2
        win = np.pad(win, (left, n_fft - win_length - left))
    yp = np.pad(y, n_fft // 2, mode="reflect")
    frames = 1 + (len(yp) - n_fft) // hop_length
    idx = np.arange(n_fft)[None, :] + hop_length * np.arange(frames)[:, None]
    return np.abs(np.fft.rfft(yp[idx] * win[None, :], axis=1)).T        # (bins, frames)


def melspectrogram(y, sample_rate, n_fft=1024, hop_length=256, win_length=1024, n_mels=80, max_norm=1.0,
                   min_level_db=-100, ref_level_db=20, fmin=50, fmax=8000, symmetric=False, preemphasize=False):
    y = np.asarray(y, dtype=np.float64)
    if preemphasize:
        y = np.concatenate([y[:1], y[1:] - 0.98 * y[:-1]])
    D = stft_abs(y, n_fft, hop_length, win_length)
    mel = melbasis.mel_filterbank(sample_rate, n_fft, n_mels, fmin, fmax).astype(np.float64) @ D
    S = 20 * np.log10(np.maximum(1e-5, mel)) - ref_level_db
    if symmetric:
        out = np.clip((2 * max_norm) * ((S - min_level_db) / (-min_level_db)) - max_norm, -max_norm, max_norm)
    else:
        out = np.clip(max_norm * ((S - min_level_db) / (-min_level_db)), 0, max_norm)
    return out.T.astype(np.float32)
