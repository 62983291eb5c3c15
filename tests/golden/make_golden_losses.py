"""tests/golden/losses_small.npz: the UNMODIFIED reference's GAN / feature-matching / STFT-magnitude loss values
(kantts/train/loss.py via oracle/ref_shims.py) on seeded random discriminator outputs.  Build container only."""
import itertools
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.ref_shims import import_reference  # noqa: E402

import_reference()
from kantts.train import loss as R  # noqa: E402


def main():
    g = torch.Generator().manual_seed(4321)
    outs_hat = [torch.randn(4, 13 + i, generator=g) for i in range(3)]
    outs = [torch.randn(4, 13 + i, generator=g) for i in range(3)]
    fm_hat = [[torch.randn(2, 4, 10 + j, generator=g) for j in range(3)] for _ in range(2)]
    fm = [[torch.randn(2, 4, 10 + j, generator=g) for j in range(3)] for _ in range(2)]
    x_mag, y_mag = torch.rand(3, 20, 17, generator=g) + 0.1, torch.rand(3, 20, 17, generator=g) + 0.1
    a = {}
    for i in range(3):
        a[f"outs_hat{i}"], a[f"outs{i}"] = outs_hat[i].numpy(), outs[i].numpy()
    for d in range(2):
        for j in range(3):
            a[f"fm_hat{d}_{j}"], a[f"fm{d}_{j}"] = fm_hat[d][j].numpy(), fm[d][j].numpy()
    a["x_mag"], a["y_mag"] = x_mag.numpy(), y_mag.numpy()
    for lt, avg in itertools.product(("mse", "hinge"), (True, False)):
        tag = f"{lt}_{int(avg)}"
        a["gen_" + tag] = np.float64(R.GeneratorAdversarialLoss(avg, lt)(outs_hat))
        a["gen1_" + tag] = np.float64(R.GeneratorAdversarialLoss(avg, lt)(outs_hat[0]))
        real, fake = R.DiscriminatorAdversarialLoss(avg, lt)(outs_hat, outs)
        a["dis_" + tag] = np.asarray([float(real), float(fake)])
        nested_hat = [[torch.zeros(1), o] for o in outs_hat]
        nested = [[torch.zeros(1), o] for o in outs]
        real, fake = R.DiscriminatorAdversarialLoss(avg, lt)(nested_hat, nested)
        a["disn_" + tag] = np.asarray([float(real), float(fake)])
    for al, ad in itertools.product((True, False), (True, False)):
        a[f"fm_{int(al)}{int(ad)}"] = np.float64(R.FeatureMatchLoss(al, ad)(fm_hat, fm))
    a["sc"] = np.float64(R.SpectralConvergenceLoss()(x_mag, y_mag))
    a["logmag"] = np.float64(R.LogSTFTMagnitudeLoss()(x_mag, y_mag))
    np.savez_compressed(os.path.join(HERE, "losses_small.npz"), cfg=np.frombuffer(b"{}", dtype=np.uint8), **a)
    print("losses_small:", len(a), "arrays")


if __name__ == "__main__":
    main()
