"""tests/golden/gen_small_nsf_{causal,noncausal}.npz: the UNMODIFIED reference Generator with the neural-source-filter
branch (``nsf_params``; hifigan.py:119-166, layers.py:229-290) on CPU.  The excitation draws random phases and noise from
torch's global RNG: the forward runs right after ``torch.manual_seed(RNG_SEED)`` (stored in the fixture) so that a
restatement drawing in the same order reproduces it.  Build container only:  python tests/golden/make_golden_nsf.py"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.ref_shims import import_reference  # noqa: E402

import_reference()
from kantts.models.hifigan.hifigan import Generator  # noqa: E402

RNG_SEED = 99


def main():
    for causal in (True, False):
        cfg = dict(in_channels=8, out_channels=1, channels=32, kernel_size=7, upsample_scales=[4, 2],
                   upsample_kernal_sizes=[8, 4], resblock_kernel_sizes=[3, 7], resblock_dilations=[[1, 3], [1, 3]],
                   causal=causal, nsf_params={"nb_harmonics": 7, "sampling_rate": 16000})
        torch.manual_seed(1234)
        g = Generator(**cfg).eval()
        gen = torch.Generator().manual_seed(5)
        mel = torch.randn(2, 8, 12, generator=gen)
        pitch = 80 + 240 * torch.rand(2, 1, 12, generator=gen)
        uv = (torch.rand(2, 1, 12, generator=gen) > 0.3).float()
        x = torch.cat([mel, pitch * uv, uv], 1)
        with torch.no_grad():
            torch.manual_seed(RNG_SEED)
            y = g(x)
        arrays = {"sd/" + k: v.detach().numpy().copy() for k, v in g.state_dict().items()}
        arrays["x"], arrays["y"] = x.numpy(), y.numpy()
        arrays["rng_seed"] = np.asarray(RNG_SEED)
        name = f"gen_small_nsf_{'causal' if causal else 'noncausal'}"
        np.savez_compressed(os.path.join(HERE, name + ".npz"), cfg=np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8),
                            **arrays)
        print(name, tuple(y.shape), float(y.abs().mean()), len(arrays), "arrays")


if __name__ == "__main__":
    main()
