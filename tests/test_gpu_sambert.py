"""GPU parity of the SAM-BERT path: kernels against plain fp32 torch math on
the same seeded inputs, the model against the golden vectors of the unmodified reference and against the CPU
oracle, through the C ABI (libkantts_b200.so via ctypes)."""
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _sops():
    from kantts_b200 import sambert_ops
    return sambert_ops


# ------------------------------------------------------------------------------------------------
# kernels
# ------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("rows,c", [(7, 8), (33, 32), (100, 48), (257, 128), (64, 512), (19, 1024), (3000, 128)])
def test_layernorm_matches_torch(rows, c):
    g = torch.Generator().manual_seed(rows * 1000 + c)
    x = torch.randn(rows, c, generator=g) * 2 + 0.5
    w = torch.randn(c, generator=g)
    b = torch.randn(c, generator=g)
    r = torch.randn(rows, c, generator=g)
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    yr = F.layer_norm(xr, (c,), wr, br, 1e-6)
    (yr * r).sum().backward()
    xg, wg, bg = (t.to(DEV).requires_grad_(True) for t in (x, w, b))
    y = _sops().layer_norm(xg.view(1, rows, c), wg, bg, 1e-6)
    (y.view(rows, c) * r.to(DEV)).sum().backward()
    assert rel_l2(y.view(rows, c).cpu(), yr) < 2e-6
    assert rel_l2(xg.grad.cpu(), xr.grad) < 1e-5
    assert rel_l2(wg.grad.cpu(), wr.grad) < 1e-5
    assert rel_l2(bg.grad.cpu(), br.grad) < 1e-5


def _ref_attention(q, k, v, mask, n_head, keep=None, p_drop=0.0):
    """sambert/__init__.py:17-29,80-100 in plain torch: q (B,Lq,HD), k/v (B,Lk,HD) -> (B,Lq,HD), (H*B,Lq,Lk)."""
    B, Lq, hd = q.shape
    d = hd // n_head

    def split(t):
        return t.view(B, t.shape[1], n_head, d).permute(2, 0, 1, 3).reshape(n_head * B, t.shape[1], d)

    a = torch.bmm(split(q), split(k).transpose(1, 2)) / math.sqrt(d)
    if mask is not None:
        m = mask if mask.dim() == 3 else mask.unsqueeze(1).expand(-1, Lq, -1)
        a = a.masked_fill(m.expand(B, -1, -1).repeat(n_head, 1, 1), float("-inf"))
    a = torch.softmax(a, dim=2)
    if keep is not None:
        a = a * keep.float() / (1.0 - p_drop)
    o = torch.bmm(a, split(v))
    return o.view(n_head, B, Lq, d).permute(1, 2, 0, 3).reshape(B, Lq, hd), a


@pytest.mark.parametrize("B,H,D,L,mask_kind,drop", [
    (2, 2, 8, 10, "pad", 0.0), (3, 2, 16, 37, "pad", 0.0), (2, 8, 16, 256, "none", 0.0), (2, 4, 16, 300, "pad", 0.0),
    (1, 2, 32, 70, "full", 0.0), (1, 2, 64, 45, "pad", 0.0), (2, 2, 16, 600, "pad", 0.0), (1, 1, 16, 1100, "none", 0.0),
    (2, 4, 16, 64, "pad", 0.25),
])
def test_self_attention_matches_torch(B, H, D, L, mask_kind, drop):
    sops = _sops()
    g = torch.Generator().manual_seed(B * 100 + L)
    hd = H * D
    qkv = torch.randn(B, L, 3 * hd, generator=g)
    r = torch.randn(B, L, hd, generator=g)
    mask = None
    if mask_kind == "pad":
        lens = torch.tensor([L - 3 * i for i in range(B)])
        mask = torch.arange(L)[None, :] >= lens[:, None]
    elif mask_kind == "full":
        mask = torch.rand(B, L, L, generator=g) < 0.3
        mask[:, :, 0] = False
    keep = (torch.rand(H * B, L, L, generator=g) >= drop) if drop > 0 else None
    qr = qkv.clone().requires_grad_(True)
    q, k, v = qr.chunk(3, -1)
    o_ref, a_ref = _ref_attention(q, k, v, mask, H, keep, drop)
    (o_ref * r).sum().backward()
    qg = qkv.to(DEV).requires_grad_(True)
    o, a = sops.SelfAttnFn.apply(qg, None if mask is None else mask.to(DEV), H, drop,
                                 None if keep is None else keep.to(DEV))
    (o * r.to(DEV)).sum().backward()
    assert rel_l2(o.cpu(), o_ref) < 5e-6
    assert rel_l2(a.cpu(), a_ref) < 5e-6
    assert rel_l2(qg.grad.cpu(), qr.grad) < 2e-5


def test_pnca_attention_matches_torch():
    sops = _sops()
    g = torch.Generator().manual_seed(11)
    B, H, D, L = 3, 4, 16, 50
    hd = H * D
    x_qkv = torch.randn(B, L, 3 * hd, generator=g)
    h_kv = torch.randn(B, L, 2 * hd, generator=g)
    rx = torch.randn(B, L, hd, generator=g)
    rh = torch.randn(B, L, hd, generator=g)
    i = torch.arange(L)[:, None]
    j = torch.arange(L)[None, :]
    lens = torch.tensor([50, 44, 31])
    pad = (torch.arange(L)[None, :] >= lens[:, None]).unsqueeze(1).expand(-1, L, -1)
    mx = (~((j >= (i - 2).clamp_min(0)) & (j <= i))[None] | pad).masked_fill(pad.transpose(1, 2), False)
    mh = (~((j >= i) & (j <= i + 2))[None] | pad).masked_fill(pad.transpose(1, 2), False)
    xr, hr = x_qkv.clone().requires_grad_(True), h_kv.clone().requires_grad_(True)
    q, k, v = xr.chunk(3, -1)
    hk, hv = hr.chunk(2, -1)
    ox_ref, ax_ref = _ref_attention(q, k, v, mx, H)
    oh_ref, ah_ref = _ref_attention(q, hk, hv, mh, H)
    ((ox_ref * rx).sum() + (oh_ref * rh).sum()).backward()
    xg, hg = x_qkv.to(DEV).requires_grad_(True), h_kv.to(DEV).requires_grad_(True)
    ox, oh, ax, ah = sops.PncaAttnFn.apply(xg, hg, mx.to(DEV), mh.to(DEV), H)
    ((ox * rx.to(DEV)).sum() + (oh * rh.to(DEV)).sum()).backward()
    for got, want in ((ox, ox_ref), (oh, oh_ref), (ax, ax_ref), (ah, ah_ref)):
        assert rel_l2(got.cpu(), want) < 5e-6
    assert rel_l2(xg.grad.cpu(), xr.grad) < 2e-5
    assert rel_l2(hg.grad.cpu(), hr.grad) < 2e-5


@pytest.mark.parametrize("B,T,C,K,shift,masked", [(2, 20, 16, 5, 0, True), (3, 300, 128, 41, 0, True),
                                                   (2, 770, 256, 41, 17, True), (1, 33, 40, 7, 2, False)])
def test_fsmn_memory_block_matches_torch(B, T, C, K, shift, masked):
    sops = _sops()
    g = torch.Generator().manual_seed(T + C)
    x = torch.randn(B, T, C, generator=g)
    w = torch.randn(C, 1, K, generator=g) * 0.2
    r = torch.randn(B, T, C, generator=g)
    lp = int(round((K - 1) / 2)) + shift
    rp = int((K - 1) / 2) - shift
    mask = (torch.arange(T)[None, :] >= torch.tensor([T - 5 * i for i in range(B)])[:, None]) if masked else None
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    xm = xr if mask is None else xr.masked_fill(mask.unsqueeze(-1), 0)
    yr = F.conv1d(F.pad(xm, (0, 0, lp, rp)).transpose(1, 2), wr, None, groups=C).transpose(1, 2) + xm
    if mask is not None:
        yr = yr.masked_fill(mask.unsqueeze(-1), 0)
    (yr * r).sum().backward()
    xg, wg = x.to(DEV).requires_grad_(True), w.to(DEV).requires_grad_(True)
    y = sops.FsmnMemoryFn.apply(xg, wg, None if mask is None else mask.to(DEV), lp)
    (y * r.to(DEV)).sum().backward()
    assert rel_l2(y.cpu(), yr) < 2e-6
    assert rel_l2(xg.grad.cpu(), xr.grad) < 1e-5
    assert rel_l2(wg.grad.cpu(), wr.grad) < 1e-5


def test_length_regulator_matches_oracle():
    from kantts_b200 import sambert
    from oracle import sambert as osb
    g = torch.Generator().manual_seed(3)
    B, L, C, r = 4, 12, 40, 3
    dur = torch.randint(0, 6, (B, L), generator=g)
    dur[:, 0] += 1
    x = torch.randn(B, L, C, generator=g)
    total = dur.sum(1)
    T = int(total.max())
    mask = osb.length_mask(torch.clamp(total - torch.tensor([0, 3, 1, 7]), min=1), T)
    xr = x.clone().requires_grad_(True)
    want, want_len = osb.length_regulator(xr, dur, mask, r)
    rr = torch.randn(want.shape, generator=g)
    (want * rr).sum().backward()
    xg = x.to(DEV).requires_grad_(True)
    got, got_len = sambert.LengthRegulator(r)(xg, dur.to(DEV), mask.to(DEV))
    (got * rr.to(DEV)).sum().backward()
    assert torch.equal(got.cpu(), want.detach())
    assert torch.equal(got_len.cpu(), want_len)
    assert rel_l2(xg.grad.cpu(), xr.grad) < 1e-6


# ------------------------------------------------------------------------------------------------
# model
# ------------------------------------------------------------------------------------------------


def _run_model(cfg, sd, batch, force_ffma):
    from kantts_b200 import ops, sambert
    model = sambert.KanTtsSAMBERT(cfg)
    model.load_state_dict(sd, strict=True)
    model = model.to(DEV).eval()
    b = {k: v.to(DEV) for k, v in batch.items()}
    ops.set_force_ffma(force_ffma)
    try:
        with torch.backends.cudnn.flags(enabled=False):      # cuDNN refuses LSTM backward in eval mode
            res = model(b["inputs_ling"], b["inputs_emotion"], b["inputs_speaker"], b["input_lengths"],
                        output_lengths=b["output_lengths"], mel_targets=b["mel_targets"],
                        duration_targets=b["duration_targets"], pitch_targets=b["pitch_targets"],
                        energy_targets=b["energy_targets"])
            l0, l1 = sambert.MelReconLoss()(b["output_lengths"], b["mel_targets"], res["dec_outputs"],
                                            res["postnet_outputs"])
            dl, pl, el = sambert.ProsodyReconLoss()(res["valid_inter_lengths"], res["duration_targets"],
                                                    res["pitch_targets"], res["energy_targets"],
                                                    res["log_duration_predictions"], res["pitch_predictions"],
                                                    res["energy_predictions"])
            total = l0 + l1 + dl + pl + el
            total.backward()
    finally:
        ops.set_force_ffma(False)
    return model, res, [float(v) for v in (l0, l1, dl, pl, el, total)]


OUT_KEYS = ("dec_outputs", "postnet_outputs", "log_duration_predictions", "pitch_predictions", "energy_predictions",
            "LR_text_outputs", "LR_emo_outputs", "LR_spk_outputs")


@pytest.mark.parametrize("path", ["ffma", "tcgen05"])
def test_sambert_small_matches_reference_golden(golden, path):
    g = golden("sambert_small")
    ffma = path == "ffma"
    tol_o, tol_g = (1e-5, 2e-4) if ffma else (1e-4, 1e-3)
    model, res, losses = _run_model(g.cfg, g.group("sd/"), g.group("in/"), ffma)
    for k in OUT_KEYS:
        assert rel_l2(res[k].cpu(), g.t("out/" + k)) < tol_o, (k, rel_l2(res[k].cpu(), g.t("out/" + k)))
    assert torch.equal(res["LR_length_rounded"].cpu(), g.t("out/LR_length_rounded"))
    assert [res["x_band_width"], res["h_band_width"]] == g.t("out/band_width").tolist()
    for k in ("enc_slf_attn_lst", "pnca_x_attn_lst", "pnca_h_attn_lst"):
        assert len(res[k]) == sum(1 for n in g.arrays if n.startswith(f"out/{k}."))
        for i, a in enumerate(res[k]):
            assert rel_l2(a.cpu(), g.t(f"out/{k}.{i}")) < tol_o, (k, i)
    want = g.t("out/losses").tolist()
    for got, w in zip(losses, want):
        assert abs(got - w) < 1e-4 * max(1.0, abs(w)), (losses, want)     # mel-L1 <= 1e-4 (north_star)
    grads = g.group("grad/")
    named = dict(model.named_parameters())
    worst = 0.0
    for k, w in grads.items():
        got = named[k].grad
        assert got is not None, k
        e = rel_l2(got.cpu(), w)
        if float(w.abs().max()) > 1e-6:
            worst = max(worst, e)
            assert e < tol_g, (k, e)
    assert worst > 0.0


@pytest.mark.parametrize("path", ["ffma", "tcgen05"])
def test_sambert_medium_matches_oracle(path):
    """Full-width sambert_24k.yaml layers on a short ragged batch vs the CPU oracle, both compute paths.
    Outputs / losses: <= 1e-4 on both.  Gradients: the exact-fp32 path agrees to 1e-5 per tensor.  On the bf16x3
    tensor-core path the forward carries ~1e-5 relative error, which flips the sign of the ~1e-5 fraction of the
    34 ReLU layers' pre-activations that lie that close to zero; each flipped unit changes the (discontinuous)
    gradient by O(1), i.e. a relative gradient difference of ~sqrt(1e-5) = 3e-3 that is not an arithmetic error
    (measured: median 3.2e-3, max 9e-3; profiles/r01_notes.md).  That path is therefore held to a direction
    test (cosine similarity of the whole gradient) plus a loose per-tensor bound."""
    import kantts_b200
    from kantts_b200 import sambert
    from oracle import sambert as osb
    from golden.make_batch import make_sambert_batch
    cfg = kantts_b200.sambert_24k_config()
    torch.manual_seed(77)
    ref = sambert.KanTtsSAMBERT(cfg)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    batch = make_sambert_batch(cfg, B=3, L=24, gen=torch.Generator().manual_seed(78), short=5)
    torch.set_num_threads(16)
    sdo = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "position_enc" not in k
               and "inv_timescales" not in k else v) for k, v in sd.items()}
    want = osb.sambert_forward(sdo, cfg, batch["inputs_ling"], batch["inputs_emotion"], batch["inputs_speaker"],
                               batch["input_lengths"], batch["output_lengths"], batch["mel_targets"],
                               batch["duration_targets"], batch["pitch_targets"], batch["energy_targets"])
    total, parts = osb.total_loss(want, batch)
    total.backward()
    model, res, losses = _run_model(cfg, sd, batch, path == "ffma")
    for k in OUT_KEYS:
        assert rel_l2(res[k].cpu(), want[k].detach()) < 1e-4, (k, rel_l2(res[k].cpu(), want[k].detach()))
    for got, w in zip(losses, list(parts) + [total]):
        assert abs(got - float(w)) < 1e-4 * max(1.0, abs(float(w)))
    errs, dot, n1, n2 = [], 0.0, 0.0, 0.0
    for k, p in model.named_parameters():
        if p.requires_grad:
            w = sdo[k].grad
            assert p.grad is not None and w is not None, k
            g = p.grad.cpu().double()
            dot += float((g * w.double()).sum()); n1 += float((g * g).sum()); n2 += float((w.double() ** 2).sum())
            if float(w.abs().max()) > 1e-7:
                errs.append(rel_l2(p.grad.cpu(), w))
    errs = torch.tensor(errs)
    cos = dot / math.sqrt(n1 * n2)
    if path == "ffma":
        assert float(errs.median()) < 1e-5 and float(errs.max()) < 1e-4, (float(errs.median()), float(errs.max()))
    else:
        assert cos > 1 - 1e-4, cos
        assert float(errs.median()) < 1e-2 and float(errs.max()) < 5e-2, (float(errs.median()), float(errs.max()))


INFER_KEYS = ("log_duration_predictions", "pitch_predictions", "energy_predictions", "LR_text_outputs", "LR_emo_outputs",
              "LR_spk_outputs", "dec_outputs", "postnet_outputs")


def _infer_model(cfg, sd, inputs, ffma):
    from kantts_b200 import ops, sambert
    ops.set_force_ffma(ffma)
    try:
        model = sambert.KanTtsSAMBERT(cfg)
        model.load_state_dict(sd, strict=True)
        model = model.to(DEV).eval()
        with torch.no_grad(), torch.backends.cudnn.flags(enabled=False):
            res = model(inputs["inputs_ling"].to(DEV), inputs["inputs_emotion"].to(DEV), inputs["inputs_speaker"].to(DEV),
                        inputs["input_lengths"].to(DEV))
        torch.cuda.synchronize()
    finally:
        ops.set_force_ffma(False)
    return res


@pytest.mark.parametrize("path", ["ffma", "tcgen05"])
def test_sambert_free_running_inference_matches_reference_golden(golden, path):
    """inference (no targets) -- predicted prosody, autoregressive duration predictor, step-by-step PNCA
    decoding on a preallocated K/V state -- against the unmodified reference's batch-1 run."""
    g = golden("sambert_small_infer")
    res = _infer_model(g.cfg, g.group("sd/"), g.group("in/"), path == "ffma")
    tol = 2e-5 if path == "ffma" else 2e-4
    assert torch.equal(res["LR_length_rounded"].cpu(), g.t("out/LR_length_rounded"))
    assert [res["x_band_width"], res["h_band_width"]] == g.t("out/band_width").tolist()
    for k in INFER_KEYS:
        assert res[k].shape == g.t("out/" + k).shape, (k, res[k].shape)
        assert rel_l2(res[k].cpu(), g.t("out/" + k)) < tol, (k, rel_l2(res[k].cpu(), g.t("out/" + k)))
    for k in ("pnca_x_attn_lst", "pnca_h_attn_lst"):
        assert len(res[k]) == g.cfg["decoder_num_layers"]
        for i, a in enumerate(res[k]):
            assert rel_l2(a.cpu(), g.t(f"out/{k}.{i}")) < tol, (k, i)


def test_sambert_free_running_inference_batch_matches_oracle(golden):
    """The same on a ragged batch of 3 (the reference cannot: its decoder masks drop the batch dimension) against the
    CPU oracle's per-item restatement."""
    from oracle import sambert as osb
    from golden.make_batch import make_sambert_batch
    g = golden("sambert_small_infer")
    batch = make_sambert_batch(g.cfg, B=3, L=9, gen=torch.Generator().manual_seed(31), short=3)
    inputs = {k: batch[k] for k in ("inputs_ling", "inputs_emotion", "inputs_speaker", "input_lengths")}
    with torch.no_grad():
        want = osb.sambert_infer(g.group("sd/"), g.cfg, inputs["inputs_ling"], inputs["inputs_emotion"],
                                 inputs["inputs_speaker"], inputs["input_lengths"])
    dur = torch.exp(want["log_duration_predictions"]) - 1
    frac = (dur + 0.5) - torch.floor(dur + 0.5)
    assert float(torch.minimum(frac, 1 - frac)[dur > 0].min()) > 2e-3       # no duration sits on a rounding boundary
    res = _infer_model(g.cfg, g.group("sd/"), inputs, True)
    assert torch.equal(res["LR_length_rounded"].cpu(), want["LR_length_rounded"])
    assert res["x_band_width"] == want["x_band_width"]
    for k in INFER_KEYS:
        assert res[k].shape == want[k].shape, (k, res[k].shape, want[k].shape)
        assert rel_l2(res[k].cpu(), want[k]) < 2e-5, (k, rel_l2(res[k].cpu(), want[k]))


def test_sambert_c4_train_step_runs_and_learns():
    """BASELINE configs[3] shape (batch 32, 256 symbols, 768 mel frames, 80 mels), train() mode with the yaml's
    dropouts, through SambertStep: losses finite, parameters move, the loss goes down over a few steps."""
    import kantts_b200
    from kantts_b200 import sambert
    from golden.make_batch import make_c4_batch
    cfg = kantts_b200.sambert_24k_config()
    torch.manual_seed(1234)
    model = sambert.KanTtsSAMBERT(cfg).to(DEV).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.98), eps=1e-9)
    sch = kantts_b200.train.NoamLR(opt, warmup_steps=40)
    step = kantts_b200.SambertStep(model, opt, sch, {"MelReconLoss": sambert.MelReconLoss(),
                                                     "ProsodyReconLoss": sambert.ProsodyReconLoss()})
    batch = {k: v.to(DEV) for k, v in make_c4_batch(cfg, torch.Generator().manual_seed(1234)).items()}
    before = model.mel_decoder.mel_dec.dec_out_proj.weight.detach().clone()
    hist = []
    for _ in range(8):
        out = step.step(batch)
        hist.append(float(out["TotalLoss"]))
    assert all(math.isfinite(v) for v in hist), hist
    assert hist[-1] < hist[0], hist
    assert not torch.equal(before, model.mel_decoder.mel_dec.dec_out_proj.weight.detach())
    assert out["x_band_width"] == 1
