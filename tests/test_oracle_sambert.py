"""CPU: the SAM-BERT oracle restatement (oracle/sambert.py) against the golden vectors produced by the
unmodified reference (tests/golden/make_golden_sambert.py)."""
import torch

from conftest import rel_l2
from oracle import sambert as osb


def _run(g, requires_grad=False):
    sd = g.group("sd/")
    if requires_grad:
        for k, v in sd.items():
            if v.dtype.is_floating_point and "position_enc" not in k and "inv_timescales" not in k:
                v.requires_grad_(True)
    b = g.group("in/")
    res = osb.sambert_forward(sd, g.cfg, b["inputs_ling"], b["inputs_emotion"], b["inputs_speaker"],
                              b["input_lengths"], b["output_lengths"], b["mel_targets"], b["duration_targets"],
                              b["pitch_targets"], b["energy_targets"])
    return sd, b, res


def test_sambert_forward_matches_reference(golden):
    g = golden("sambert_small")
    with torch.no_grad():
        _, _, res = _run(g)
    for k in ("dec_outputs", "postnet_outputs", "log_duration_predictions", "pitch_predictions",
              "energy_predictions", "LR_text_outputs", "LR_emo_outputs", "LR_spk_outputs"):
        assert res[k].shape == g.t("out/" + k).shape, k
        assert rel_l2(res[k], g.t("out/" + k)) < 2e-6, (k, rel_l2(res[k], g.t("out/" + k)))
    assert torch.equal(res["LR_length_rounded"], g.t("out/LR_length_rounded"))
    assert [res["x_band_width"], res["h_band_width"]] == g.t("out/band_width").tolist()
    for k in ("enc_slf_attn_lst", "pnca_x_attn_lst", "pnca_h_attn_lst"):
        for i, a in enumerate(res[k]):
            assert rel_l2(a, g.t(f"out/{k}.{i}")) < 2e-6, (k, i)


def test_sambert_losses_and_grads_match_reference(golden):
    g = golden("sambert_small")
    sd, b, res = _run(g, requires_grad=True)
    total, parts = osb.total_loss(res, b)
    want = g.t("out/losses")
    for got, w in zip(list(parts) + [total], want):
        assert abs(float(got) - float(w)) < 2e-6 * max(1.0, abs(float(w)))
    total.backward()
    grads = g.group("grad/")
    assert len(grads) > 100
    for k, w in grads.items():
        got = sd[k].grad
        assert got is not None, k
        assert rel_l2(got, w) < 2e-5 or float((got - w).abs().max()) < 1e-7, (k, rel_l2(got, w))


def test_sambert_free_running_inference_matches_reference(golden):
    """Inference branch (predicted prosody, autoregressive duration predictor, step-by-step PNCA decoding with K/V
    state) against the unmodified reference's batch-1 run (tests/golden/make_golden_sambert_infer.py)."""
    g = golden("sambert_small_infer")
    b = g.group("in/")
    with torch.no_grad():
        res = osb.sambert_infer(g.group("sd/"), g.cfg, b["inputs_ling"], b["inputs_emotion"], b["inputs_speaker"],
                                b["input_lengths"])
    assert torch.equal(res["LR_length_rounded"], g.t("out/LR_length_rounded"))
    assert [res["x_band_width"], res["h_band_width"]] == g.t("out/band_width").tolist()
    for k in ("log_duration_predictions", "pitch_predictions", "energy_predictions", "LR_text_outputs", "LR_emo_outputs",
              "LR_spk_outputs", "dec_outputs", "postnet_outputs"):
        assert res[k].shape == g.t("out/" + k).shape, (k, res[k].shape, g.t("out/" + k).shape)
        assert rel_l2(res[k], g.t("out/" + k)) < 5e-6, (k, rel_l2(res[k], g.t("out/" + k)))
    for k in ("pnca_x_attn_lst", "pnca_h_attn_lst"):
        for i, a in enumerate(res[k]):
            assert rel_l2(a, g.t(f"out/{k}.{i}")) < 5e-6, (k, i)
