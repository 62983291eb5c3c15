"""Host-side logic added in round 2 that needs no GPU: the optimizer-step hook that invalidates prepared weights, the
restructured losses against their definitions (kantts/train/loss.py:108-256), the pair_state plumbing."""
import torch
import torch.nn.functional as F

import kantts_b200 as K
from kantts_b200 import loss as L
from kantts_b200 import ops


def test_optimizer_step_bumps_parameter_epochs_even_when_version_does_not_move():
    """torch's fused Adam updates through its own kernel and leaves Tensor._version untouched: the prepared-weight cache key
    must change anyway (ops._bump_param_epochs, registered as a global optimizer-step post hook)."""
    for kwargs in ({"fused": True}, {"foreach": True}, {}):
        p = torch.nn.Parameter(torch.randn(16))
        q = torch.nn.Parameter(torch.randn(4))            # not owned by the optimizer: must not move
        opt = torch.optim.Adam([p], lr=1e-3, **kwargs)
        p.grad = torch.randn(16)
        e0 = getattr(p, "_kt_epoch", 0)
        opt.step()
        assert getattr(p, "_kt_epoch", 0) == e0 + 1, kwargs
        assert getattr(q, "_kt_epoch", 0) == 0
    sgd_p = torch.nn.Parameter(torch.randn(3))
    sgd_p.grad = torch.randn(3)
    torch.optim.SGD([sgd_p], lr=0.1).step()
    assert sgd_p._kt_epoch == 1


def test_adversarial_losses_match_their_definition():
    torch.manual_seed(0)
    outs = [torch.randn(3, n) for n in (7, 11, 5)]
    outs_hat = [torch.randn(3, n) for n in (7, 11, 5)]
    for avg in (True, False):
        gen = L.GeneratorAdversarialLoss(average_by_discriminators=avg)
        ref = sum(F.mse_loss(o, torch.ones_like(o)) for o in outs_hat)
        ref = ref / len(outs_hat) if avg else ref
        assert torch.allclose(gen(outs_hat), ref, rtol=1e-6, atol=1e-7)
        dis = L.DiscriminatorAdversarialLoss(average_by_discriminators=avg)
        real, fake = dis(outs_hat, outs)
        r_ref = sum(F.mse_loss(o, torch.ones_like(o)) for o in outs)
        f_ref = sum(F.mse_loss(o, torch.zeros_like(o)) for o in outs_hat)
        if avg:
            r_ref, f_ref = r_ref / 3, f_ref / 3
        assert torch.allclose(real, r_ref, rtol=1e-6, atol=1e-7) and torch.allclose(fake, f_ref, rtol=1e-6, atol=1e-7)


def test_concatenated_mse_chain_equals_the_per_discriminator_sum():
    """The CUDA path of _mse_total (cat, sub, mul, dot with cached 1 / numel weights) -- evaluated here with the same ops on
    CPU tensors -- equals sum_i mean((o_i - t)^2), value and gradient."""
    torch.manual_seed(1)
    outs = [torch.randn(2, n, requires_grad=True) for n in (5, 9, 3)]
    for avg in (True, False):
        ref = sum(torch.mean((o - 1.0) ** 2) for o in outs)
        ref = ref / 3 if avg else ref
        g_ref = torch.autograd.grad(ref, outs)
        scale = 1.0 / 3 if avg else 1.0
        w = torch.cat([torch.full((o.numel(),), scale / o.numel()) for o in outs])
        d = torch.cat([o.reshape(-1) for o in outs]) - 1.0
        val = torch.dot(d * w, d)
        g = torch.autograd.grad(val, outs)
        assert torch.allclose(val, ref, rtol=1e-6, atol=1e-7)
        for a, b in zip(g, g_ref):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-7)


def test_feature_match_loss_generic_path_matches_definition():
    torch.manual_seed(2)
    feats_hat = [[torch.randn(2, 4, 9), torch.randn(2, 8, 5)], [torch.randn(2, 3, 7, 2)]]
    feats = [[torch.randn_like(t) for t in maps] for maps in feats_hat]
    for by_layers in (True, False):
        for by_disc in (True, False):
            fm = L.FeatureMatchLoss(average_by_layers=by_layers, average_by_discriminators=by_disc)
            ref = 0.0
            for mh, m in zip(feats_hat, feats):
                v = sum(F.l1_loss(a, b) for a, b in zip(mh, m))
                ref = ref + (v / len(mh) if by_layers else v)
            ref = ref / len(feats) if by_disc else ref
            assert torch.allclose(fm(feats_hat, feats), ref, rtol=1e-6, atol=1e-7)
            assert fm._forward_accumulated(feats_hat, feats) is None      # CPU tensors: no accumulator kernel, generic path


def test_pair_state_nests_and_none_is_off():
    assert ops._pair_state is None
    with ops.pair_state("record"):
        assert ops._pair_state == ("record", None)
        with ops.pair_state("reuse", 4):
            assert ops._pair_state == ("reuse", 4)
        with ops.pair_state(None):
            assert ops._pair_state is None
        assert ops._pair_state == ("record", None)
    assert ops._pair_state is None


def test_builder_asks_for_fused_adam_only_on_cuda():
    cfg = {"Model": {"Generator": {"params": dict(in_channels=8, out_channels=1, channels=16, kernel_size=3, upsample_scales=[2],
                                                  upsample_kernal_sizes=[4], resblock_kernel_sizes=[3], resblock_dilations=[[1]]),
                                   "optimizer": {"type": "Adam", "params": {"lr": 1e-3}}, "scheduler": {"type": "StepLR", "params": {"step_size": 10}}}}}
    model, opt, sched = K.hifigan_model_builder(cfg, "cpu")
    assert not opt["generator"].defaults.get("fused")
    model, opt, sched = K.hifigan_model_builder(cfg, "cpu", fused_optimizer=True)
    assert opt["generator"].defaults.get("fused")
