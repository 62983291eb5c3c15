"""BASELINE configs[4] shape of flow: symbols -> SAM-BERT free-running decode -> mel -> HiFi-GAN generator -> wav, for a
ragged batch, against the CPU oracle's composition of the same two restatements.

(First green GPU run: round 2, gpurun call r2a -- 3 passed.)"""
import pytest
import torch

from conftest import rel_l2

pytestmark = [pytest.mark.gpu]


def test_synthesize_matches_oracle_composition(golden):
    import kantts_b200 as K
    from kantts_b200 import ops
    from oracle import hifigan as OH, sambert as OS
    from golden.make_batch import make_sambert_batch
    g = golden("sambert_small_infer")
    cfg = g.cfg
    batch = make_sambert_batch(cfg, B=3, L=9, gen=torch.Generator().manual_seed(31), short=3)
    gcfg = dict(in_channels=cfg["num_mels"], channels=32, upsample_scales=[4, 2], upsample_kernal_sizes=[8, 4],
                resblock_kernel_sizes=[3, 7], resblock_dilations=[[1, 3], [1, 3]])
    torch.manual_seed(7)
    gen = K.Generator(**gcfg)
    gsd = {k: v.detach().clone() for k, v in gen.state_dict().items()}
    with torch.no_grad():
        want = OS.sambert_infer(g.group("sd/"), cfg, batch["inputs_ling"], batch["inputs_emotion"],
                                batch["inputs_speaker"], batch["input_lengths"])
        wav_o = OH.generator_forward(gsd, want["postnet_outputs"].transpose(1, 2), **gcfg)
    am = K.KanTtsSAMBERT(cfg)
    am.load_state_dict(g.group("sd/"), strict=True)
    am, gen = am.to("cuda").eval(), gen.to("cuda").eval()
    ops.set_force_ffma(True)
    try:
        with torch.backends.cudnn.flags(enabled=False):
            wavs, res = K.synthesize(am, gen, *(batch[k].to("cuda") for k in
                                                ("inputs_ling", "inputs_emotion", "inputs_speaker", "input_lengths")))
    finally:
        ops.set_force_ffma(False)
    hop = 8
    for b, w in enumerate(wavs):
        n = int(want["LR_length_rounded"][b]) * hop
        assert w.shape == (n,)
        assert float((w.cpu() - wav_o[b, 0, :n]).pow(2).mean().sqrt()) < 1e-3      # waveform RMS tolerance
        assert rel_l2(w.cpu(), wav_o[b, 0, :n]) < 1e-4


@pytest.mark.parametrize("name", ["gen_small_nsf_causal", "gen_small_nsf_noncausal"])
def test_nsf_generator_matches_reference_golden(golden, name):
    """Generator(nsf_params=...) -- source module + per-stage source_downs -- against the unmodified
    reference's output (same RNG seed: the excitation's random phases / noise are drawn on the host like the reference)."""
    import kantts_b200 as K
    from kantts_b200 import ops
    g = golden(name)
    m = K.Generator(**g.cfg)
    m.load_state_dict(g.group("sd/"), strict=True)
    m = m.to("cuda").eval()
    for force in (True, False):
        ops.set_force_ffma(force)
        try:
            with torch.no_grad():
                torch.manual_seed(int(g.arrays["rng_seed"]))
                y = m(g.t("x").to("cuda")).cpu()
        finally:
            ops.set_force_ffma(False)
        assert y.shape == g.t("y").shape
        assert float((y - g.t("y")).pow(2).mean().sqrt()) < (1e-5 if force else 1e-3), (force,)
