"""GPU data path: kantts_b200.data.GpuVocBatcher vs a numpy restatement of Voc_Dataset.__getitem__ /
collate_fn (kantts/datasets/dataset.py:225-311) fed with the OFFLINE mel features of the oracle
(oracle/dsp.py = kantts/preprocess/audio_processor/core/dsp.py:165-201) -- same numpy RNG seed, so the same crops."""
import numpy as np
import pytest
import torch

from oracle import dsp as ODSP

pytestmark = [pytest.mark.gpu]


def _reference_batch(wavs, idx, sr, hop, n_fft, bms, aux, mel_kw, mean, std):
    """dataset.py:225-311 on the host (numpy), mel features = the offline pipeline's .npy content."""
    bmf = bms // hop
    items = []
    for i in idx:
        w = wavs[i]
        mel = ODSP.melspectrogram(w, sr, n_fft=n_fft, hop_length=hop, win_length=n_fft, **mel_kw)
        if mean is not None:
            mel = (mel - mean[None, :]) / std[None, :]          # audio_processor.py:363-382 norm_mean_std
        if mel.shape[0] <= bmf:                                  # dataset.py:249-262
            mel = np.concatenate((mel, np.zeros((bmf - mel.shape[0] + 1, mel.shape[1]))), axis=0)
            wc = np.zeros(mel.shape[0] * hop, dtype=np.float32)
            wc[: len(w)] = w
            w = wc
        else:                                                    # dataset.py:263-266
            w = np.pad(w, (0, n_fft), mode="reflect")[: len(mel) * hop]
        assert len(mel) * hop == len(w)
        items.append((w, mel))
    lengths = [len(m) for _, m in items]
    start = np.array([np.random.randint(aux, n + (-(bmf + aux))) for n in lengths])      # dataset.py:282-287
    wav_b = np.asarray([w[s * hop: s * hop + bms] for (w, _), s in zip(items, start)])
    mel_b = np.asarray([m[s - aux: s - aux + bmf + aux + aux] for (_, m), s in zip(items, start)])
    return torch.tensor(wav_b, dtype=torch.float32).unsqueeze(1), torch.tensor(mel_b, dtype=torch.float32).transpose(2, 1)


@pytest.mark.parametrize("hop,aux,mean_std", [(256, 0, False), (256, 2, True), (240, 1, False)])
def test_gpu_collate_matches_reference_flow(hop, aux, mean_std):
    import kantts_b200 as K
    from kantts_b200.data import GpuVocBatcher
    rs = np.random.RandomState(3)
    sr, n_fft, bms = 24000, 1024, hop * 16
    wavs = [(0.3 * rs.randn(n)).astype(np.float32) for n in (9000, 20011, hop * 16 + 5, 3000, 14500)]   # incl. short utterances
    mel_kw = dict(n_mels=80, fmin=50, fmax=8000, max_norm=1.0, min_level_db=-100, ref_level_db=20)
    mean = (0.4 + 0.1 * rs.rand(80)).astype(np.float32) if mean_std else None
    std = (0.2 + 0.1 * rs.rand(80)).astype(np.float32) if mean_std else None
    b = GpuVocBatcher(wavs, sr, hop, n_fft=n_fft, win_length=n_fft, batch_max_steps=bms, aux_context_window=aux,
                      mel_mean=mean, mel_std=std, **mel_kw)
    # (with aux_context_window > 0 the reference itself cannot crop an utterance of <= batch_max_frames + 2 * aux frames:
    #  np.random.randint(low >= high) raises in collate_fn -- such utterances only appear in the aux = 0 runs)
    batches = ((0, [0, 1, 2, 3]), (1, [4, 4, 1, 0, 2]), (7, [3, 2])) if aux == 0 else ((0, [0, 1, 4]), (1, [4, 4, 1, 0]), (7, [1]))
    for seed, idx in batches:
        np.random.seed(seed)
        wav_ref, mel_ref = _reference_batch(wavs, idx, sr, hop, n_fft, bms, aux, mel_kw, mean, std)
        np.random.seed(seed)
        wav, mel = b.collate(idx)
        assert wav.shape == wav_ref.shape and mel.shape == mel_ref.shape
        assert torch.equal(wav.cpu(), wav_ref)                                           # crops are exact copies
        scale = 1.0 if not mean_std else float(1.0 / std.min())
        assert float((mel.cpu() - mel_ref).abs().mean()) < 1e-4 * scale                  # mel-L1 tolerance of the north star
        assert float((mel.cpu() - mel_ref).abs().max()) < 5e-3 * scale
