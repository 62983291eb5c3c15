"""CPU: host-side logic of the SAM-BERT drop-in (no kernels are launched): state_dict / RNG parity with the
reference (through the golden fixture), the duration-span index arithmetic against the oracle's one-hot
formulation, ABI struct layout, and the no-CPU-fallback rule."""
import ctypes

import pytest
import torch

import kantts_b200
from kantts_b200 import _lib, sambert
from oracle import sambert as osb


def test_state_dict_keys_shapes_and_seeded_init_match_reference(golden):
    g = golden("sambert_small")
    ref = g.group("sd/")
    torch.manual_seed(1234)                      # the seed tests/golden/make_golden_sambert.py used
    model = sambert.KanTtsSAMBERT(g.cfg)
    sd = model.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    same = 0
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(ref[k].shape), k
        # the generator script perturbs biases and LayerNorm parameters after construction; every other
        # tensor must be bit-identical to the reference's seeded initialisation
        if not (k.endswith("bias") or "layer_norm" in k or k.endswith("ln.weight")):
            assert torch.equal(v, ref[k]), k
            same += 1
    assert same > 60
    model.load_state_dict(ref, strict=True)
    n_train = sum(p.numel() for p in model.parameters() if p.requires_grad)
    n_ref = sum(v.numel() for k, v in ref.items() if "position_enc" not in k and "inv_timescales" not in k)
    assert n_train == n_ref


def test_full_size_parameter_count():
    """12 297 155 trainable parameters for sambert_24k.yaml + PinYin unit sizes."""
    model = sambert.KanTtsSAMBERT(kantts_b200.sambert_24k_config())
    assert sum(p.numel() for p in model.parameters() if p.requires_grad) == 12297155


@pytest.mark.parametrize("with_mask", [False, True])
def test_duration_spans_match_one_hot_formulation(with_mask):
    gen = torch.Generator().manual_seed(5)
    B, L, r = 4, 9, 3
    dur = torch.randint(0, 5, (B, L), generator=gen)
    dur[:, 0] += 1
    x = torch.randn(B, L, 6, generator=gen)
    total = dur.sum(1)
    T = int(total.max())
    mask = None
    if with_mask:
        lens = torch.clamp(total - torch.tensor([0, 2, 0, 5]), min=1)
        mask = osb.length_mask(lens, T)
    idx, start, count, pos, out_lens = sambert._duration_spans(dur, T if with_mask else None, mask, r)
    want, want_lens = osb.length_regulator(x, dur, mask, r)
    got = torch.where(idx[..., None] >= 0, torch.gather(x, 1, idx.clamp_min(0).long()[..., None].expand(-1, -1, 6)),
                      torch.zeros(()))
    assert torch.equal(got, want)
    assert torch.equal(out_lens, want_lens)
    inv = torch.tensor([10000.0 ** (2 * (i // 2) / 8) for i in range(8)])
    want_pe = osb.dur_position_encoding(dur, mask, inv, r)
    enc = sambert.DurSinusoidalPositionEncoder(8, r)
    assert torch.allclose(enc.encode(pos), want_pe, atol=1e-6)
    # spans: every frame that copies symbol i lies inside [start, start + count)
    for b in range(B):
        for t in range(idx.shape[1]):
            i = int(idx[b, t])
            if i >= 0:
                assert int(start[b, i]) <= t < int(start[b, i]) + int(count[b, i])


def test_attn_desc_layout_matches_header():
    assert ctypes.sizeof(_lib.KtAttnDesc) == 56
    assert _lib.KtAttnDesc.mask_b_stride.offset == 40
    assert _lib.KtAttnDesc.keep_scale.offset == 52


def test_sambert_has_no_cpu_fallback(golden):
    g = golden("sambert_small")
    model = sambert.KanTtsSAMBERT(g.cfg).eval()
    b = g.group("in/")
    with pytest.raises(RuntimeError):
        model(b["inputs_ling"], b["inputs_emotion"], b["inputs_speaker"], b["input_lengths"],
              output_lengths=b["output_lengths"], mel_targets=b["mel_targets"], duration_targets=b["duration_targets"],
              pitch_targets=b["pitch_targets"], energy_targets=b["energy_targets"])
