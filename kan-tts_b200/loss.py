"""``kantts.train.loss`` surface for the HiFi-GAN path (KAN-TTS kantts/train/loss.py:108-441,
512-544): same class names / kwargs / return values, so ``criterion_builder`` works unchanged.

The mel / STFT spectra come from the fused kernel (audio.py); the scalar reductions over the
(tiny) discriminator outputs and spectra stay torch ops (a few KB each).  The feature-matching
loss is value-only in the reference (its second argument is detached and the trainer passes the
generator's feature maps there, trainer.py:535-538) and uses the streaming kt_l1_sum kernel.
"""
import torch
import torch.nn.functional as F

from . import ops
from .audio import MelSpectrogram, stft


class GeneratorAdversarialLoss(torch.nn.Module):
    def __init__(self, average_by_discriminators=True, loss_type="mse"):
        super().__init__()
        self.average_by_discriminators = average_by_discriminators
        assert loss_type in ["mse", "hinge"], f"{loss_type} is not supported."
        self.criterion = self._mse_loss if loss_type == "mse" else self._hinge_loss

    def forward(self, outputs):
        if isinstance(outputs, (tuple, list)):
            adv_loss = 0.0
            for i, outputs_ in enumerate(outputs):
                adv_loss += self.criterion(outputs_)
            if self.average_by_discriminators:
                adv_loss /= i + 1
        else:
            adv_loss = self.criterion(outputs)
        return adv_loss

    def _mse_loss(self, x):
        return F.mse_loss(x, x.new_ones(x.size()))

    def _hinge_loss(self, x):
        return -x.mean()


class DiscriminatorAdversarialLoss(torch.nn.Module):
    def __init__(self, average_by_discriminators=True, loss_type="mse"):
        super().__init__()
        self.average_by_discriminators = average_by_discriminators
        assert loss_type in ["mse", "hinge"], f"{loss_type} is not supported."
        if loss_type == "mse":
            self.fake_criterion, self.real_criterion = self._mse_fake_loss, self._mse_real_loss
        else:
            self.fake_criterion, self.real_criterion = self._hinge_fake_loss, self._hinge_real_loss

    def forward(self, outputs_hat, outputs):
        if isinstance(outputs, (tuple, list)):
            real_loss, fake_loss = 0.0, 0.0
            for i, (outputs_hat_, outputs_) in enumerate(zip(outputs_hat, outputs)):
                if isinstance(outputs_hat_, (tuple, list)):
                    outputs_hat_, outputs_ = outputs_hat_[-1], outputs_[-1]
                real_loss += self.real_criterion(outputs_)
                fake_loss += self.fake_criterion(outputs_hat_)
            if self.average_by_discriminators:
                fake_loss /= i + 1
                real_loss /= i + 1
        else:
            real_loss = self.real_criterion(outputs)
            fake_loss = self.fake_criterion(outputs_hat)
        return real_loss, fake_loss

    def _mse_real_loss(self, x):
        return F.mse_loss(x, x.new_ones(x.size()))

    def _mse_fake_loss(self, x):
        return F.mse_loss(x, x.new_zeros(x.size()))

    def _hinge_real_loss(self, x):
        return -torch.mean(torch.min(x - 1, x.new_zeros(x.size())))

    def _hinge_fake_loss(self, x):
        return -torch.mean(torch.min(-x - 1, x.new_zeros(x.size())))


def _l1_mean(a, b):
    """F.l1_loss(a, b.detach()) for two feature maps.  Channels-last views produced by this package
    share strides, so the permuted-back contiguous buffers are compared directly."""
    if a.requires_grad or not a.is_cuda:
        return F.l1_loss(a, b.detach())
    if a.dim() == 3 and not a.is_contiguous():
        a, b = a.transpose(1, 2), b.transpose(1, 2)
    elif a.dim() == 4 and not a.is_contiguous():
        a, b = a.permute(0, 2, 3, 1), b.permute(0, 2, 3, 1)
    if not (a.is_contiguous() and b.is_contiguous()):
        return F.l1_loss(a, b.detach())
    return ops.l1_sum(a.detach(), b.detach(), 1.0 / a.numel())


class FeatureMatchLoss(torch.nn.Module):
    def __init__(self, average_by_layers=True, average_by_discriminators=True):
        super().__init__()
        self.average_by_layers = average_by_layers
        self.average_by_discriminators = average_by_discriminators

    def forward(self, feats_hat, feats):
        feat_match_loss = 0.0
        for i, (feats_hat_, feats_) in enumerate(zip(feats_hat, feats)):
            feat_match_loss_ = 0.0
            for j, (feat_hat_, feat_) in enumerate(zip(feats_hat_, feats_)):
                feat_match_loss_ += _l1_mean(feat_hat_, feat_)
            if self.average_by_layers:
                feat_match_loss_ /= j + 1
            feat_match_loss += feat_match_loss_
        if self.average_by_discriminators:
            feat_match_loss /= i + 1
        return feat_match_loss


class MelSpectrogramLoss(torch.nn.Module):
    def __init__(self, fs=22050, fft_size=1024, hop_size=256, win_length=None, window="hann", num_mels=80,
                 fmin=80, fmax=7600, center=True, normalized=False, onesided=True, eps=1e-10, log_base=10.0):
        super().__init__()
        self.mel_spectrogram = MelSpectrogram(fs=fs, fft_size=fft_size, hop_size=hop_size, win_length=win_length,
                                              window=window, num_mels=num_mels, fmin=fmin, fmax=fmax,
                                              center=center, normalized=normalized, onesided=onesided, eps=eps,
                                              log_base=log_base)

    def forward(self, y_hat, y):
        mel_hat = self.mel_spectrogram(y_hat)
        mel = self.mel_spectrogram(y)
        return F.l1_loss(mel_hat, mel)


class SpectralConvergenceLoss(torch.nn.Module):
    def forward(self, x_mag, y_mag):
        return torch.norm(y_mag - x_mag, p="fro") / torch.norm(y_mag, p="fro")


class LogSTFTMagnitudeLoss(torch.nn.Module):
    def forward(self, x_mag, y_mag):
        return F.l1_loss(torch.log(y_mag), torch.log(x_mag))


class STFTLoss(torch.nn.Module):
    def __init__(self, fft_size=1024, shift_size=120, win_length=600, window="hann_window"):
        super().__init__()
        self.fft_size, self.shift_size, self.win_length = fft_size, shift_size, win_length
        self.spectral_convergence_loss = SpectralConvergenceLoss()
        self.log_stft_magnitude_loss = LogSTFTMagnitudeLoss()
        self.register_buffer("window", getattr(torch, window)(win_length))

    def forward(self, x, y):
        x_mag = stft(x, self.fft_size, self.shift_size, self.win_length, self.window)
        y_mag = stft(y, self.fft_size, self.shift_size, self.win_length, self.window)
        return self.spectral_convergence_loss(x_mag, y_mag), self.log_stft_magnitude_loss(x_mag, y_mag)


class MultiResolutionSTFTLoss(torch.nn.Module):
    def __init__(self, fft_sizes=[1024, 2048, 512], hop_sizes=[120, 240, 50], win_lengths=[600, 1200, 240],
                 window="hann_window"):
        super().__init__()
        assert len(fft_sizes) == len(hop_sizes) == len(win_lengths)
        self.stft_losses = torch.nn.ModuleList()
        for fs, ss, wl in zip(fft_sizes, hop_sizes, win_lengths):
            self.stft_losses += [STFTLoss(fs, ss, wl, window)]

    def forward(self, x, y):
        if len(x.shape) == 3:
            x = x.view(-1, x.size(2))
            y = y.view(-1, y.size(2))
        sc_loss, mag_loss = 0.0, 0.0
        for f in self.stft_losses:
            sc_l, mag_l = f(x, y)
            sc_loss += sc_l
            mag_loss += mag_l
        return sc_loss / len(self.stft_losses), mag_loss / len(self.stft_losses)


loss_dict = {
    "generator_adv_loss": GeneratorAdversarialLoss,
    "discriminator_adv_loss": DiscriminatorAdversarialLoss,
    "stft_loss": MultiResolutionSTFTLoss,
    "mel_loss": MelSpectrogramLoss,
    "subband_stft_loss": MultiResolutionSTFTLoss,
    "feat_match_loss": FeatureMatchLoss,
}


def criterion_builder(config, device="cpu"):
    """loss.py:528-544"""
    criterion = {}
    for key, value in config["Loss"].items():
        if key in loss_dict:
            if value["enable"]:
                criterion[key] = loss_dict[key](**value.get("params", {})).to(device)
                setattr(criterion[key], "weights", value.get("weights", 1.0))
        else:
            raise NotImplementedError("{} is not implemented".format(key))
    return criterion
