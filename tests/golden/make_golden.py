"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED
reference (/root/reference, imported through oracle/ref_shims.py) on CPU.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Seed: torch.manual_seed(1234) (the reference's DATASET_RANDOM_SEED,
kantts/datasets/dataset.py:16).  Every fixture stores the module kwargs (json),
the reference ``state_dict`` (reference keys), the inputs, the outputs and the
autograd gradients of a fixed scalar, so that both the oracle restatement
(oracle/hifigan.py, CPU) and the CUDA path can be checked against the real
reference without it being present.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.ref_shims import import_reference  # noqa: E402

import_reference()
from kantts.models.hifigan.hifigan import (  # noqa: E402
    Generator, MultiPeriodDiscriminator, MultiScaleDiscriminator)
from kantts.train.loss import (  # noqa: E402
    MelSpectrogramLoss, MultiResolutionSTFTLoss, criterion_builder)
from kantts.utils.audio_torch import MelSpectrogram  # noqa: E402

torch.set_num_threads(8)


def sd_np(module):
    return {"sd/" + k: v.detach().cpu().numpy().copy() for k, v in module.state_dict().items()}


def save(name, cfg, **arrays):
    path = os.path.join(HERE, name + ".npz")
    arrays = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()}
    np.savez_compressed(path, cfg=np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8), **arrays)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB, {len(arrays)} arrays")


def randomize(module, gen):
    """Give biases / weight_g non-trivial values so parity is not vacuous."""
    with torch.no_grad():
        for n, p in module.named_parameters():
            if n.endswith("bias"):
                p.copy_(0.05 * torch.randn(p.shape, generator=gen))
            if n.endswith("weight_g"):
                p.mul_(1.0 + 0.2 * torch.randn(p.shape, generator=gen))


def gen_fixture(name, cfg, B, T, seed):
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    g = Generator(**cfg)
    randomize(g, gen)
    before = sd_np(g)
    x = torch.randn(B, cfg.get("in_channels", 80), T, generator=gen, requires_grad=True)
    y = g(x)
    r = torch.randn(y.shape, generator=gen)
    loss = (y * r).sum()
    params = dict(g.named_parameters())
    grads = torch.autograd.grad(loss, [x] + list(params.values()))
    arrays = dict(before)
    arrays.update(x=x, y=y, r=r, grad_x=grads[0])
    for (k, _), gr in zip(params.items(), grads[1:]):
        arrays["grad/" + k] = gr
    save(name, cfg, **arrays)


def disc_fixture(name, cls, cfg, B, T, seed):
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    d = cls(**cfg)
    randomize(d, gen)
    d.train()
    before = sd_np(d)
    y = (0.3 * torch.randn(B, 1, T, generator=gen)).clamp(-1, 1).requires_grad_(True)
    outs, fmaps = d(y)
    loss = 0.0
    arrays = dict(before)
    for i, o in enumerate(outs):
        r = torch.randn(o.shape, generator=gen)
        arrays[f"r{i}"] = r
        arrays[f"out{i}"] = o
        loss = loss + (o * r).sum()
        for l, f in enumerate(fmaps[i]):
            arrays[f"fmap{i}_{l}"] = f
    params = dict(d.named_parameters())
    grads = torch.autograd.grad(loss, [y] + list(params.values()))
    arrays.update(y=y, grad_y=grads[0])
    for (k, _), gr in zip(params.items(), grads[1:]):
        arrays["grad/" + k] = gr
    for k, v in d.state_dict().items():       # spectral-norm u/v after one training forward
        if k.endswith("weight_u") or (k.endswith("weight_v") and k[:-1] + "u" in d.state_dict()):
            arrays["after/" + k] = v
    save(name, cfg, **arrays)


def mel_fixture(seed):
    gen = torch.Generator().manual_seed(seed)
    y = (0.1 * torch.randn(2, 1, 8192, generator=gen)).clamp(-1, 1)
    y_hat = (y + 0.05 * torch.randn(2, 1, 8192, generator=gen)).requires_grad_(True)
    arrays = dict(y=y, y_hat=y_hat)
    for tag, cfg in (("default", {}),
                     ("yaml24k", dict(fs=24000, fft_size=1024, hop_size=240, win_length=1024,
                                      window="hann", num_mels=80, fmin=0, fmax=8000, log_base=None)),
                     ("c2", dict(fs=22050, fft_size=1024, hop_size=256, win_length=1024,
                                 window="hann", num_mels=80, fmin=0, fmax=8000))):
        m = MelSpectrogram(**cfg)
        arrays[f"mel_{tag}"] = m(y)
        arrays[f"melmat_{tag}"] = m.melmat
        lf = MelSpectrogramLoss(**cfg)
        loss = lf(y_hat, y)
        arrays[f"loss_{tag}"] = loss
        arrays[f"grad_{tag}"] = torch.autograd.grad(loss, y_hat)[0]
    st = MultiResolutionSTFTLoss()
    sc, mag = st(y_hat, y)
    arrays["stft_sc"], arrays["stft_mag"] = sc, mag
    gsc, = torch.autograd.grad(sc + mag, y_hat)
    arrays["stft_grad"] = gsc
    save("mel_stft", {"note": "MelSpectrogram / MelSpectrogramLoss / MultiResolutionSTFTLoss"}, **arrays)


SMALL_G_CAUSAL = dict(in_channels=80, out_channels=1, channels=64, kernel_size=7,
                      upsample_scales=[8, 8, 2, 2], upsample_kernal_sizes=[16, 16, 4, 4],
                      resblock_kernel_sizes=[3, 7, 11], resblock_dilations=[[1, 3, 5]] * 3,
                      bias=True, causal=True, nonlinear_activation="LeakyReLU",
                      nonlinear_activation_params={"negative_slope": 0.1}, use_weight_norm=True)
SMALL_G_NONCAUSAL = dict(in_channels=80, out_channels=1, channels=32, kernel_size=7,
                         upsample_scales=[5, 4, 2], upsample_kernal_sizes=[11, 8, 4],
                         resblock_kernel_sizes=[3, 7], resblock_dilations=[[1, 3, 5], [1, 3]],
                         bias=True, causal=False, nonlinear_activation="LeakyReLU",
                         nonlinear_activation_params={"negative_slope": 0.1}, use_weight_norm=True)
SMALL_MPD = dict(periods=[2, 3, 5, 7, 11], discriminator_params=dict(
    in_channels=1, out_channels=1, kernel_sizes=[5, 3], channels=4,
    downsample_scales=[3, 3, 3, 3, 1], max_downsample_channels=32, bias=True,
    nonlinear_activation="LeakyReLU", nonlinear_activation_params={"negative_slope": 0.1},
    use_spectral_norm=False))
SMALL_MSD = dict(scales=3, downsample_pooling="DWT",
                 downsample_pooling_params={"kernel_size": 4, "stride": 2, "padding": 2},
                 discriminator_params=dict(
                     in_channels=1, out_channels=1, kernel_sizes=[15, 41, 5, 3], channels=16,
                     max_downsample_channels=64, max_groups=16, bias=True,
                     downsample_scales=[4, 4, 4, 4, 1], nonlinear_activation="LeakyReLU",
                     nonlinear_activation_params={"negative_slope": 0.1}),
                 follow_official_norm=True)
LOSS_CFG = {
    "generator_adv_loss": {"enable": True, "params": {"average_by_discriminators": False}, "weights": 1.0},
    "discriminator_adv_loss": {"enable": True, "params": {"average_by_discriminators": False}, "weights": 1.0},
    "stft_loss": {"enable": False},
    "mel_loss": {"enable": True, "params": dict(fs=22050, fft_size=1024, hop_size=256, win_length=1024,
                                                window="hann", num_mels=80, fmin=0, fmax=8000),
                 "weights": 45.0},
    "subband_stft_loss": {"enable": False},
    "feat_match_loss": {"enable": True, "params": {"average_by_discriminators": False,
                                                   "average_by_layers": False}, "weights": 2.0},
}


def trainstep_fixture(seed):
    """One GAN_Trainer.train_step (trainer.py:469-589) of the unmodified trainer
    on small models: logged losses + parameters after the Adam steps."""
    from kantts.train.trainer import GAN_Trainer

    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    g = Generator(**SMALL_G_CAUSAL)
    msd = MultiScaleDiscriminator(**SMALL_MSD)
    mpd = MultiPeriodDiscriminator(**SMALL_MPD)
    for m in (g, msd, mpd):
        randomize(m, gen)
    arrays = {}
    for tag, m in (("g", g), ("msd", msd), ("mpd", mpd)):
        for k, v in m.state_dict().items():
            arrays[f"before/{tag}/{k}"] = v.detach().clone()
    model = {"generator": g, "discriminator": {"MultiScaleDiscriminator": msd, "MultiPeriodDiscriminator": mpd}}
    adam = dict(lr=2e-4, betas=(0.5, 0.9), weight_decay=0.0)
    optimizer = {"generator": torch.optim.Adam(g.parameters(), **adam),
                 "discriminator": {"MultiScaleDiscriminator": torch.optim.Adam(msd.parameters(), **adam),
                                   "MultiPeriodDiscriminator": torch.optim.Adam(mpd.parameters(), **adam)}}
    sched = lambda o: torch.optim.lr_scheduler.MultiStepLR(o, milestones=[200000], gamma=0.5)  # noqa: E731
    scheduler = {"generator": sched(optimizer["generator"]),
                 "discriminator": {k: sched(v) for k, v in optimizer["discriminator"].items()}}
    config = {"Loss": LOSS_CFG, "generator_train_start_steps": 1, "discriminator_train_start_steps": 0,
              "generator_grad_norm": -1, "discriminator_grad_norm": -1, "log_interval_steps": 1000,
              "train_max_steps": 10, "save_interval_steps": 10 ** 9, "eval_interval_steps": 10 ** 9}
    criterion = criterion_builder(config)
    import tempfile
    tr = GAN_Trainer(config=config, model=model, optimizer=optimizer, scheduler=scheduler,
                     criterion=criterion, device=torch.device("cpu"), sampler={"train": None, "valid": None},
                     train_loader=None, valid_loader=None, max_steps=10, save_dir=tempfile.mkdtemp(),
                     save_interval=10 ** 9, valid_interval=10 ** 9, log_interval=10 ** 9)
    tr.steps = 1
    B, Tm = 2, 8
    y = (0.1 * torch.randn(B, 1, Tm * 256, generator=gen)).clamp(-1, 1)
    x = torch.randn(B, 80, Tm, generator=gen)
    tr.train_step((y, x))
    for k, v in tr.total_train_loss.items():
        arrays["loss/" + k.replace("train/", "")] = np.float64(v)
    for tag, m in (("g", g), ("msd", msd), ("mpd", mpd)):
        for k, v in m.state_dict().items():
            arrays[f"after/{tag}/{k}"] = v.detach().clone()
    arrays["y"], arrays["x"] = y, x
    save("trainstep_small", {"generator": SMALL_G_CAUSAL, "msd": SMALL_MSD, "mpd": SMALL_MPD,
                             "loss": LOSS_CFG, "adam": {"lr": 2e-4, "betas": [0.5, 0.9]}}, **arrays)


def c1_fixture():
    """BASELINE config 1: class-default Generator under torch.manual_seed(1234),
    x = randn(1, 80, 32) -> (1, 1, 8192).  Weights are NOT stored (60 MB): the
    product module reproduces the reference's construction-time RNG stream, which
    tests/test_modules_cpu.py verifies against this checksum."""
    torch.manual_seed(1234)
    g = Generator()
    g.eval()
    gen = torch.Generator().manual_seed(1234)
    x = torch.randn(1, 80, 32, generator=gen)
    with torch.no_grad():
        y = g(x)
    sd = g.state_dict()
    csum = {k: [float(v.double().sum()), float(v.double().abs().sum())] for k, v in sd.items()}
    save("c1_generator", {"checksums": csum}, x=x, y=y)


if __name__ == "__main__":
    gen_fixture("gen_small_causal", SMALL_G_CAUSAL, 2, 6, 1234)
    gen_fixture("gen_small_noncausal", SMALL_G_NONCAUSAL, 2, 7, 1235)
    disc_fixture("mpd_small", MultiPeriodDiscriminator, SMALL_MPD, 2, 2050, 1236)
    disc_fixture("msd_small", MultiScaleDiscriminator, SMALL_MSD, 2, 2048, 1237)
    mel_fixture(1238)
    trainstep_fixture(1239)
    c1_fixture()
