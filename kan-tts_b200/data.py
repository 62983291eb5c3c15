"""GPU data path of the HiFi-GAN trainer: ``Voc_Dataset.__getitem__`` length fix-up + ``collate_fn``
random crop / batching (kantts/datasets/dataset.py:225-311) and the OFFLINE mel features the reference reads from
``mel/*.npy`` (kantts/preprocess/audio_processor/core/dsp.py:165-201, mean/std normalisation of
audio_processor.py:363-382) -- here computed on the fly, on the device, for exactly the frames of each crop.

The reference keeps (wav, mel.npy) pairs on disk, crops both on the host in ``collate_fn`` and copies the batch over.
``GpuVocBatcher`` keeps the utterance waveforms resident in HBM (one padded 2-D tensor), draws the crop positions on
the host with numpy's global RNG in the reference's order (``np.random.randint(start_offset, length + end_offset)`` per
item: the same seed gives the same crops), gathers the waveform segments with one indexing kernel and computes the mel
frames of the crop (plus the ``aux_context_window`` frames on either side) with the fused STFT-mel kernel
(``kt_stft_mel_fwd``: framing, hann window, rFFT, magnitude, Slaney mel, dB, normalisation in one launch).  Nothing but
the start frames crosses PCIe per step, and the ``mel/*.npy`` feature format disappears.

A mel frame f of the offline features is centred at sample f * hop of the utterance, reflect-padded by n_fft / 2 at the
utterance ends (librosa.stft).  The crop's frames are therefore computed from the slice of the utterance-level PADDED
waveform that covers their windows, so a cropped-then-computed frame equals the computed-then-cropped frame of the
reference bit for bit in exact arithmetic (fp32 kernel vs float64 numpy: tests/test_gpu_data.py, mel-L1 <= 1e-4)."""
import numpy as np
import torch

from . import ops
from .audio import _padded_window, slaney_mel_filterbank


class GpuVocBatcher:
    def __init__(self, wavs, sampling_rate, hop_length, n_fft=1024, win_length=1024, n_mels=80, fmin=50, fmax=8000,
                 batch_max_steps=8192, aux_context_window=0, max_norm=1.0, min_level_db=-100, ref_level_db=20,
                 symmetric=False, preemphasize=False, mel_mean=None, mel_std=None, device="cuda"):
        """wavs: list of 1-D float arrays (already loaded / resampled / trimmed like librosa.load in __getitem__).
        The keyword names are those of the reference's ``audio_config`` + ``batch_max_steps`` / ``aux_context_window``."""
        if batch_max_steps % hop_length != 0:                      # dataset.py:77-79
            batch_max_steps += -(batch_max_steps % hop_length)
        self.hop, self.n_fft, self.win_length, self.n_mels = int(hop_length), int(n_fft), int(win_length), int(n_mels)
        self.batch_max_steps = int(batch_max_steps)
        self.batch_max_frames = self.batch_max_steps // self.hop
        self.aux = int(aux_context_window)
        self.start_offset = self.aux                               # dataset.py:84-85
        self.end_offset = -(self.batch_max_frames + self.aux)
        self.device = torch.device(device)
        self.preemphasize = bool(preemphasize)
        # frames of one crop: [start - aux, start + batch_max_frames + aux)
        self.n_frames = self.batch_max_frames + 2 * self.aux
        # window start = first frame centre - n_fft/2 - delta, delta making the first wanted frame an INTEGER frame index f0
        # of the slice (the fused kernel frames a signal at multiples of hop from its start, centred)
        half = self.n_fft // 2
        self.delta = (-half) % self.hop
        self.f0 = (half + self.delta) // self.hop
        self.win_samples = (self.n_frames - 1 + 2 * self.f0) * self.hop + 1     # slice length: frames f0 .. f0 + n - 1 are interior
        pad_l = half + self.delta                                   # reflect padding (n_fft / 2) + alignment zeros
        # ---- Voc_Dataset.__getitem__ length fix-up (dataset.py:248-270), then utterance-level reflect padding
        fixed, frames = [], []
        for w in wavs:
            w = np.asarray(w, dtype=np.float32)
            n_mel = len(w) // self.hop + 1                          # frames of the offline mel (centred STFT)
            if n_mel <= self.batch_max_frames:                      # short utterance: zero-extend features and audio
                n_fr = self.batch_max_frames + 1
                w2 = np.zeros(n_fr * self.hop, dtype=np.float32)
                w2[: len(w)] = w
                short = True
            else:
                n_fr = n_mel
                w2 = np.pad(w, (0, self.n_fft), mode="reflect")[: n_fr * self.hop]
                short = False
            fixed.append((w, w2, n_fr, n_mel, short))
            frames.append(n_fr)
        self.frames = np.asarray(frames)
        L = max(len(f[1]) for f in fixed)
        Lp = max(len(f[0]) for f in fixed) + 2 * half + pad_l + self.win_samples + self.hop * (self.batch_max_frames + 2)
        wav_tab = torch.zeros(len(fixed), L, dtype=torch.float32)
        src_tab = torch.zeros(len(fixed), Lp, dtype=torch.float32)  # utterance-level padded signal the mel frames read
        self.valid_mel = []
        for i, (w, w2, n_fr, n_mel, short) in enumerate(fixed):
            wav_tab[i, : len(w2)] = torch.from_numpy(w2)
            y = w.astype(np.float64)
            if self.preemphasize:                                   # dsp.py:53-56 on the ORIGINAL utterance
                y = np.concatenate([y[:1], y[1:] - 0.98 * y[:-1]])
            yp = np.pad(y, half, mode="reflect").astype(np.float32)  # librosa.stft(center=True, pad_mode="reflect")
            src_tab[i, self.delta: self.delta + len(yp)] = torch.from_numpy(yp)
            self.valid_mel.append(n_mel)                            # frames >= n_mel of a short utterance are zero rows
        self.wav_tab = wav_tab.to(self.device)
        self.src_tab = src_tab.to(self.device)
        self.valid_mel_t = torch.tensor(self.valid_mel, device=self.device)
        melmat = slaney_mel_filterbank(sampling_rate, self.n_fft, self.n_mels, fmin, fmax)
        self.melmat = torch.from_numpy(melmat.T.copy()).float().to(self.device).contiguous()
        self.window = _padded_window(self.win_length, self.n_fft, self.device)
        if symmetric:
            self.norm = (float(ref_level_db), float(min_level_db), 2.0 * max_norm, float(max_norm), -float(max_norm), float(max_norm))
        else:
            self.norm = (float(ref_level_db), float(min_level_db), float(max_norm), 0.0, 0.0, float(max_norm))
        self.mel_mean = None if mel_mean is None else torch.as_tensor(mel_mean, dtype=torch.float32, device=self.device).view(1, -1, 1)
        self.mel_std = None if mel_std is None else torch.as_tensor(mel_std, dtype=torch.float32, device=self.device).view(1, -1, 1)
        self._steps = torch.arange(self.batch_max_steps, device=self.device)
        self._win = torch.arange(self.win_samples, device=self.device)

    def __len__(self):
        return self.wav_tab.shape[0]

    def draw_start_frames(self, idx):
        """dataset.py:282-287: one np.random.randint per item, in batch order."""
        return np.array([np.random.randint(self.start_offset, int(self.frames[i]) + self.end_offset) for i in idx])

    def collate(self, idx, start_frames=None):
        """idx: utterance indices of the batch (the sampler's job, unchanged).  -> (wav (B, 1, T), mel (B, n_mels, frames))
        on the device, the reference's ``collate_fn`` return value."""
        idx = np.asarray(idx)
        if start_frames is None:
            start_frames = self.draw_start_frames(idx)
        sf = torch.as_tensor(start_frames, device=self.device, dtype=torch.long)
        ii = torch.as_tensor(idx, device=self.device, dtype=torch.long)
        wav = self.wav_tab[ii[:, None], (sf * self.hop)[:, None] + self._steps[None, :]].unsqueeze(1)
        # slice of the padded utterance whose frame f0 + j is the offline frame (start - aux + j); a frame centre c (utterance
        # samples) sits at src_tab column c + n_fft/2 + delta, the slice's first sample is f0 * hop before the first centre
        first = sf - self.aux
        col0 = first * self.hop + (self.n_fft // 2 + self.delta) - self.f0 * self.hop
        seg = self.src_tab[ii[:, None], col0[:, None] + self._win[None, :]]
        mel = ops.StftMelFn.apply(seg.contiguous(), self.window, self.melmat, self.n_fft, self.hop, 1, 0.0, self.norm)
        mel = mel[:, :, self.f0: self.f0 + self.n_frames]
        # frames past the end of a short (zero-extended) utterance are zero feature rows (dataset.py:249-262)
        fidx = first[:, None] + torch.arange(self.n_frames, device=self.device)[None, :]
        if self.mel_mean is not None:
            mel = (mel - self.mel_mean) / self.mel_std
        mel = torch.where((fidx < self.valid_mel_t[ii][:, None])[:, None, :], mel, torch.zeros((), device=self.device))
        return wav, mel.contiguous()
