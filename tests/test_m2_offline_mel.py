"""Offline (preprocessing) mel variant, SURVEY.md section 8a row M2
(kantts/preprocess/audio_processor/core/dsp.py:165-201)."""
import numpy as np
import pytest
import torch

from oracle import dsp as ODSP


def test_oracle_stft_matches_torch_stft_reflect():
    """librosa is absent (parity unpinned): pin the restated librosa.stft framing against torch.stft."""
    y = 0.1 * np.random.RandomState(0).randn(6000)
    a = ODSP.stft_abs(y, 1024, 256, 1024)
    t = torch.stft(torch.tensor(y), 1024, 256, 1024, torch.hann_window(1024, dtype=torch.float64), center=True,
                   pad_mode="reflect", return_complex=True).abs().numpy()
    assert a.shape == t.shape and np.abs(a - t).max() < 1e-10
    m = ODSP.melspectrogram(y, 22050)
    assert m.shape == (6000 // 256 + 1, 80) and m.min() >= 0.0 and m.max() <= 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(), dict(symmetric=True, max_norm=4.0), dict(preemphasize=True, fmin=0, fmax=8000),
                                dict(n_fft=2048, hop_length=200, win_length=1000)])
def test_gpu_offline_mel_matches_oracle(kw):
    import kantts_b200 as K
    y = (0.2 * np.random.RandomState(1).randn(24000)).astype(np.float32)
    ref = ODSP.melspectrogram(y, 24000 if "n_fft" in kw else 22050, **kw)
    out = K.audio.melspectrogram(y, 24000 if "n_fft" in kw else 22050, **kw)
    assert out.shape == ref.shape
    assert float(np.abs(out - ref).mean()) < 1e-4 * max(1.0, kw.get("max_norm", 1.0))
