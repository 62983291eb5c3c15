"""kantts_b200 -- B200-native (sm_100a) implementation of KAN-TTS's HiFi-GAN hot path.

Public surface = the reference's own module API for this path:
  hifigan.Generator / MultiPeriodDiscriminator / MultiScaleDiscriminator   (kantts.models)
  audio.MelSpectrogram / stft                                             (kantts.utils.audio_torch)
  loss.* + criterion_builder                                              (kantts.train.loss)
  train.GanStep (GAN_Trainer.train_step + the data-parallel gradient exchange)
  install.install() patches these into an importable KAN-TTS checkout.
All tensor math runs in libkantts_b200.so (C ABI: include/kantts_b200.h); there is no fallback.
"""
from . import _lib  # noqa: F401
from ._lib import build_library  # noqa: F401
from . import ops, hifigan, audio, loss, train, install as _install  # noqa: F401
from .hifigan import Generator, MultiPeriodDiscriminator, MultiScaleDiscriminator  # noqa: F401
from .audio import MelSpectrogram, stft  # noqa: F401
from .loss import (MelSpectrogramLoss, MultiResolutionSTFTLoss, GeneratorAdversarialLoss,  # noqa: F401
                   DiscriminatorAdversarialLoss, FeatureMatchLoss, criterion_builder)
from .train import GanStep, hifigan_model_builder  # noqa: F401

install = _install.install
__version__ = "0.1.0"
