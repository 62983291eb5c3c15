"""Slaney mel filterbank = ``librosa.filters.mel`` (0.9.2 defaults) restated
(TEST INFRASTRUCTURE).

Reference call sites: kantts/utils/audio_torch.py:125-131 (MelSpectrogram),
kantts/preprocess/audio_processor/core/dsp.py:137-139.  librosa is not vendored
in /root/reference and not installed here -> PARITY UNPINNED; the algorithm is
librosa's published one (htk=False, norm='slaney'), computed in float64 and
cast to float32 exactly as librosa does; tests cross-check it against
``torchaudio.functional.melscale_fbanks(norm='slaney', mel_scale='slaney')``.
"""
import numpy as np

_F_SP = 200.0 / 3
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = np.log(6.4) / 27.0


def hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    mel = f / _F_SP
    log_t = f >= _MIN_LOG_HZ
    mel = np.where(log_t, _MIN_LOG_MEL + np.log(np.maximum(f, 1e-30) / _MIN_LOG_HZ) / _LOGSTEP, mel)
    return mel


def mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f = _F_SP * m
    log_t = m >= _MIN_LOG_MEL
    return np.where(log_t, _MIN_LOG_HZ * np.exp(_LOGSTEP * (m - _MIN_LOG_MEL)), f)


def mel_filterbank(sr, n_fft, n_mels=128, fmin=0.0, fmax=None):
    """-> (n_mels, 1 + n_fft//2) float32, Slaney-normalised triangles."""
    if fmax is None:
        fmax = float(sr) / 2
    n_bins = 1 + n_fft // 2
    fftfreqs = np.linspace(0, float(sr) / 2, n_bins, endpoint=True)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, n_bins), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights.astype(np.float32)
