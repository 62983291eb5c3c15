"""CPU: the GAN / feature-matching / STFT-magnitude loss classes (pure tensor reductions, no kernel involved) against
the unmodified reference's values on the same seeded inputs (tests/golden/make_golden_losses.py)."""
import itertools

import torch

from kantts_b200 import loss as L


def test_adversarial_and_feature_match_losses_match_reference(golden):
    g = golden("losses_small")
    outs_hat = [g.t(f"outs_hat{i}") for i in range(3)]
    outs = [g.t(f"outs{i}") for i in range(3)]
    fm_hat = [[g.t(f"fm_hat{d}_{j}") for j in range(3)] for d in range(2)]
    fm = [[g.t(f"fm{d}_{j}") for j in range(3)] for d in range(2)]
    close = lambda got, want: abs(float(got) - float(want)) <= 1e-6 * max(1.0, abs(float(want)))   # noqa: E731
    for lt, avg in itertools.product(("mse", "hinge"), (True, False)):
        tag = f"{lt}_{int(avg)}"
        assert close(L.GeneratorAdversarialLoss(avg, lt)(outs_hat), g.arrays["gen_" + tag]), tag
        assert close(L.GeneratorAdversarialLoss(avg, lt)(outs_hat[0]), g.arrays["gen1_" + tag]), tag
        real, fake = L.DiscriminatorAdversarialLoss(avg, lt)(outs_hat, outs)
        assert close(real, g.arrays["dis_" + tag][0]) and close(fake, g.arrays["dis_" + tag][1]), tag
        nested_hat = [[torch.zeros(1), o] for o in outs_hat]
        nested = [[torch.zeros(1), o] for o in outs]
        real, fake = L.DiscriminatorAdversarialLoss(avg, lt)(nested_hat, nested)
        assert close(real, g.arrays["disn_" + tag][0]) and close(fake, g.arrays["disn_" + tag][1]), tag
    for al, ad in itertools.product((True, False), (True, False)):
        assert close(L.FeatureMatchLoss(al, ad)(fm_hat, fm), g.arrays[f"fm_{int(al)}{int(ad)}"]), (al, ad)
    assert close(L.SpectralConvergenceLoss()(g.t("x_mag"), g.t("y_mag")), g.arrays["sc"])
    assert close(L.LogSTFTMagnitudeLoss()(g.t("x_mag"), g.t("y_mag")), g.arrays["logmag"])


def test_adversarial_losses_are_differentiable():
    x = torch.randn(3, 7, requires_grad=True)
    for lt in ("mse", "hinge"):
        for loss in (L.GeneratorAdversarialLoss(False, lt)([x, 2 * x]),
                     sum(L.DiscriminatorAdversarialLoss(False, lt)([x], [x * 0.5]))):
            (gx,) = torch.autograd.grad(loss, x)
            assert gx.shape == x.shape and torch.isfinite(gx).all()
