"""CPU: the NSF generator variant's host side -- parameter names / shapes / construction RNG stream identical to the
reference's (the checkpoint format), and the sine-plus-noise excitation (plain tensor math, host RNG like the reference)
identical to the oracle's restatement.  The conv kernels behind it need a GPU (tests/test_gpu_pipeline.py)."""
import pytest
import torch
import torch.nn.functional as F

import kantts_b200 as K
from oracle import hifigan as O


@pytest.mark.parametrize("name", ["gen_small_nsf_causal", "gen_small_nsf_noncausal"])
def test_nsf_state_dict_and_construction_rng_match_reference(golden, name):
    g = golden(name)
    sd = g.group("sd/")
    torch.manual_seed(1234)                       # the seed tests/golden/make_golden_nsf.py constructs under
    m = K.Generator(**g.cfg)
    assert m.nsf_enable
    mine = m.state_dict()
    assert list(mine.keys()) == list(sd.keys())
    for k, v in sd.items():
        assert mine[k].shape == v.shape and torch.equal(mine[k], v), k
    m.load_state_dict(sd, strict=True)


def test_nsf_excitation_matches_oracle(golden):
    g = golden("gen_small_nsf_causal")
    sd = g.group("sd/")
    m = K.Generator(**g.cfg)
    m.load_state_dict(sd, strict=True)
    x = g.t("x")
    pitch, uv = x[:, -2:-1], x[:, -1:]
    torch.manual_seed(int(g.arrays["rng_seed"]))
    e = m.source_module.excitation(pitch, uv)                                   # (B, samples, H + 1) rows
    assert e.shape == (2, 96, 8) and not e.requires_grad
    merged = torch.tanh(F.conv1d(e.transpose(1, 2), O._resolve_weight(sd, "source_module.ffn.0."),
                                 sd["source_module.ffn.0.bias"]))
    torch.manual_seed(int(g.arrays["rng_seed"]))
    want = O.nsf_excitation(sd, pitch, uv, 7, 8, 16000)
    assert float((merged - want).abs().max()) < 1e-6
