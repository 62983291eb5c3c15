"""Checkpoint format round trip (kantts/train/trainer.py:591-674): a checkpoint written in the trainer's layout
from the B200-native modules loads -- strict -- into fresh native modules AND (where a KAN-TTS checkout is importable: this
container, not the GPU box) into the unmodified reference classes, and the other way round; optimizer / scheduler states
of the reference's `torch.optim.Adam` / `MultiStepLR` restore onto the native parameters (same parameter order)."""
import os
import sys

import pytest
import torch

import kantts_b200 as K

G_CFG = dict(in_channels=80, out_channels=1, channels=32, kernel_size=7, upsample_scales=[4, 2], upsample_kernal_sizes=[8, 4],
             resblock_kernel_sizes=[3, 7], resblock_dilations=[[1, 3], [1, 3]], causal=True)
MSD_CFG = dict(scales=3, downsample_pooling="DWT", downsample_pooling_params={"kernel_size": 4, "stride": 2, "padding": 2},
               discriminator_params=dict(in_channels=1, out_channels=1, kernel_sizes=[15, 41, 5, 3], channels=16,
                                         max_downsample_channels=64, max_groups=4, bias=True, downsample_scales=[4, 4, 4, 4, 1],
                                         nonlinear_activation="LeakyReLU", nonlinear_activation_params={"negative_slope": 0.1}),
               follow_official_norm=True)
MPD_CFG = dict(periods=[2, 3], discriminator_params=dict(in_channels=1, out_channels=1, kernel_sizes=[5, 3], channels=8,
                                                         downsample_scales=[3, 3, 3, 3, 1], max_downsample_channels=64, bias=True,
                                                         nonlinear_activation="LeakyReLU",
                                                         nonlinear_activation_params={"negative_slope": 0.1}, use_spectral_norm=False))


def _build(mod):
    torch.manual_seed(3)
    model = {"generator": mod.Generator(**G_CFG),
             "discriminator": {"MultiScaleDiscriminator": mod.MultiScaleDiscriminator(**MSD_CFG),
                               "MultiPeriodDiscriminator": mod.MultiPeriodDiscriminator(**MPD_CFG)}}
    mk = lambda m: torch.optim.Adam(m.parameters(), lr=2e-4, betas=(0.5, 0.9))
    opt = {"generator": mk(model["generator"]), "discriminator": {k: mk(m) for k, m in model["discriminator"].items()}}
    sch = {"generator": torch.optim.lr_scheduler.MultiStepLR(opt["generator"], milestones=[10], gamma=0.5),
           "discriminator": {k: torch.optim.lr_scheduler.MultiStepLR(o, milestones=[10], gamma=0.5) for k, o in opt["discriminator"].items()}}
    return model, opt, sch


def _save(model, opt, sch, steps, path):      # GAN_Trainer.save_checkpoint, trainer.py:591-632
    sd = {"optimizer": {"generator": opt["generator"].state_dict(), "discriminator": {k: o.state_dict() for k, o in opt["discriminator"].items()}},
          "scheduler": {"generator": sch["generator"].state_dict(), "discriminator": {k: o.state_dict() for k, o in sch["discriminator"].items()}},
          "steps": steps,
          "model": {"generator": model["generator"].state_dict(),
                    "discriminator": {k: m.state_dict() for k, m in model["discriminator"].items()}}}
    torch.save(sd, path)


def _load(model, opt, sch, path):             # GAN_Trainer.load_checkpoint(restore_training_state=True, strict=True), :634-674
    sd = torch.load(path, map_location="cpu")
    model["generator"].load_state_dict(sd["model"]["generator"], strict=True)
    for k in sd["model"]["discriminator"]:
        model["discriminator"][k].load_state_dict(sd["model"]["discriminator"][k], strict=True)
    opt["generator"].load_state_dict(sd["optimizer"]["generator"])
    for k in sd["optimizer"]["discriminator"]:
        opt["discriminator"][k].load_state_dict(sd["optimizer"]["discriminator"][k])
    for k in sd["scheduler"]["discriminator"]:
        sch["discriminator"][k].load_state_dict(sd["scheduler"]["discriminator"][k])
    sch["generator"].load_state_dict(sd["scheduler"]["generator"])
    return sd["steps"]


def _fake_train(model, opt):
    """give the optimizers a state without any GPU: one Adam step on synthetic gradients"""
    g = torch.Generator().manual_seed(5)
    for m, o in [(model["generator"], opt["generator"])] + [(model["discriminator"][k], opt["discriminator"][k]) for k in model["discriminator"]]:
        for p in m.parameters():
            p.grad = 1e-2 * torch.randn(p.shape, generator=g)
        o.step()


def _same(a, b):
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert a[k].shape == b[k].shape and torch.equal(a[k], b[k]), k


def test_native_checkpoint_round_trip(tmp_path):
    model, opt, sch = _build(K)
    _fake_train(model, opt)
    path = os.path.join(tmp_path, "ckpt", "checkpoint_7.pth")
    os.makedirs(os.path.dirname(path))
    _save(model, opt, sch, 7, path)
    m2, o2, s2 = _build(K)
    assert _load(m2, o2, s2, path) == 7
    _same(model["generator"].state_dict(), m2["generator"].state_dict())
    for k in model["discriminator"]:
        _same(model["discriminator"][k].state_dict(), m2["discriminator"][k].state_dict())
    a, b = opt["generator"].state_dict()["state"], o2["generator"].state_dict()["state"]
    assert a.keys() == b.keys() and all(torch.equal(a[i]["exp_avg"], b[i]["exp_avg"]) for i in a)


@pytest.mark.skipif(not os.path.isdir("/root/reference/kantts"), reason="needs an importable KAN-TTS checkout (not present on the GPU box)")
def test_checkpoints_interchange_with_the_reference_classes(tmp_path):
    from oracle import ref_shims
    ref_shims.import_reference()
    import kantts.models.hifigan.hifigan as R
    model, opt, sch = _build(K)
    _fake_train(model, opt)
    p1 = os.path.join(tmp_path, "native.pth")
    _save(model, opt, sch, 3, p1)
    rm, ro, rs = _build(R)                       # reference classes, reference-side Adam / MultiStepLR
    assert _load(rm, ro, rs, p1) == 3            # native checkpoint -> reference trainer objects, strict
    _same(model["generator"].state_dict(), rm["generator"].state_dict())
    for k in model["discriminator"]:
        _same(model["discriminator"][k].state_dict(), rm["discriminator"][k].state_dict())
    _fake_train(rm, ro)
    p2 = os.path.join(tmp_path, "reference.pth")
    _save(rm, ro, rs, 4, p2)
    m3, o3, s3 = _build(K)
    assert _load(m3, o3, s3, p2) == 4            # and back
    _same(rm["generator"].state_dict(), m3["generator"].state_dict())
    a, b = ro["generator"].state_dict()["state"], o3["generator"].state_dict()["state"]
    assert a.keys() == b.keys() and all(torch.equal(a[i]["exp_avg_sq"], b[i]["exp_avg_sq"]) for i in a)
