"""Pin the CPU oracle (oracle/hifigan.py) against golden vectors produced by
the unmodified reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import hifigan as O
from oracle import dwt as ODWT
from oracle import melbasis
from conftest import rel_l2


def _leaf(sd):
    out = {}
    for k, v in sd.items():
        v = v.clone()
        if v.is_floating_point() and not (k.endswith("weight_u") or "meanpools" in k):
            v.requires_grad_(True)
        out[k] = v
    return out


@pytest.mark.parametrize("name", ["gen_small_causal", "gen_small_noncausal"])
def test_generator_matches_reference(golden, name):
    g = golden(name)
    sd = _leaf(g.group("sd/"))
    x = g.t("x").requires_grad_(True)
    y = O.generator_forward(sd, x, **g.cfg)
    assert y.shape == g.t("y").shape
    assert float((y - g.t("y")).abs().max()) < 2e-6
    keys = [k for k in sd if sd[k].requires_grad]
    grads = torch.autograd.grad((y * g.t("r")).sum(), [x] + [sd[k] for k in keys])
    assert rel_l2(grads[0], g.t("grad_x")) < 1e-5
    ref = g.group("grad/")
    for k, gr in zip(keys, grads[1:]):
        assert rel_l2(gr, ref[k]) < 2e-5, k


@pytest.mark.parametrize("name,fwd", [("mpd_small", O.mpd_forward), ("msd_small", O.msd_forward)])
def test_discriminators_match_reference(golden, name, fwd):
    g = golden(name)
    sd = _leaf(g.group("sd/"))
    for k in list(sd):                                   # spectral-norm v is a buffer too
        if k.endswith("weight_v") and k[:-1] + "u" in sd:
            sd[k] = sd[k].detach()
    y = g.t("y").requires_grad_(True)
    outs, fmaps = fwd(sd, y, True, **g.cfg)
    loss = 0.0
    for i, o in enumerate(outs):
        assert float((o - g.t(f"out{i}")).abs().max()) < 1e-5
        for l, f in enumerate(fmaps[i]):
            assert f.shape == g.t(f"fmap{i}_{l}").shape
            assert rel_l2(f, g.t(f"fmap{i}_{l}")) < 1e-5
        loss = loss + (o * g.t(f"r{i}")).sum()
    keys = [k for k in sd if sd[k].requires_grad]
    grads = torch.autograd.grad(loss, [y] + [sd[k] for k in keys])
    assert rel_l2(grads[0], g.t("grad_y")) < 1e-5
    ref = g.group("grad/")
    for k, gr in zip(keys, grads[1:]):
        assert rel_l2(gr, ref[k]) < 5e-5, k
    for k, v in g.group("after/").items():                # power-iteration buffers updated in place
        assert rel_l2(sd[k], v) < 1e-5, k


def test_mel_and_stft_losses_match_reference(golden):
    g = golden("mel_stft")
    y, y_hat = g.t("y"), g.t("y_hat").requires_grad_(True)
    cfgs = {"default": {}, "yaml24k": dict(fs=24000, fft_size=1024, hop_size=240, win_length=1024, fmin=0, fmax=8000),
            "c2": dict(fs=22050, fft_size=1024, hop_size=256, win_length=1024, fmin=0, fmax=8000)}
    for tag, cfg in cfgs.items():
        mel = O.mel_spectrogram(y, **cfg)
        assert mel.shape == g.t(f"mel_{tag}").shape
        assert float((mel - g.t(f"mel_{tag}")).abs().mean()) < 1e-5
        loss = O.mel_spectrogram_loss(y_hat, y, **cfg)
        assert abs(float(loss) - float(g.t(f"loss_{tag}"))) < 1e-5
        gr, = torch.autograd.grad(loss, y_hat)
        assert rel_l2(gr, g.t(f"grad_{tag}")) < 1e-3
    sc, mag = O.multi_resolution_stft_loss(y_hat, y)
    assert abs(float(sc) - float(g.t("stft_sc"))) < 1e-5
    assert abs(float(mag) - float(g.t("stft_mag"))) < 1e-5
    gr, = torch.autograd.grad(sc + mag, y_hat)
    assert rel_l2(gr, g.t("stft_grad")) < 1e-3


def test_train_step_matches_reference_trainer(golden):
    g = golden("trainstep_small")
    gan = O.OracleGAN(g.group("before/g/"),
                      {"MultiScaleDiscriminator": g.group("before/msd/"),
                       "MultiPeriodDiscriminator": g.group("before/mpd/")},
                      g.cfg["generator"],
                      {"MultiScaleDiscriminator": g.cfg["msd"], "MultiPeriodDiscriminator": g.cfg["mpd"]},
                      g.cfg["loss"])
    log = gan.train_step(g.t("y"), g.t("x"))
    for k in ("mel_loss", "feature_matching_loss", "generator_loss", "real_loss", "fake_loss", "discriminator_loss"):
        ref = float(g.arrays["loss/" + k])
        assert abs(log[k] - ref) <= 2e-5 * max(1.0, abs(ref)), (k, log[k], ref)
    # parameters after the Adam steps (Adam's first step is +-lr*sign(grad): compare loosely)
    for tag, sd in (("g", gan.g), ("msd", gan.d["MultiScaleDiscriminator"]), ("mpd", gan.d["MultiPeriodDiscriminator"])):
        after = g.group(f"after/{tag}/")
        before = g.group(f"before/{tag}/")
        num = den = 0.0
        for k, v in after.items():
            if not v.is_floating_point():
                continue
            num += float(((sd[k].detach() - v).double() ** 2).sum())
            den += float(((before[k] - v).double() ** 2).sum())
        assert num <= 1e-3 * den, (tag, num, den)


def test_dwt_properties():
    """db3 analysis: orthonormal (energy preserving on zero-extended signals),
    low-pass DC gain sqrt(2), documented output lengths."""
    assert abs(sum(ODWT.DEC_LO) - 2 ** 0.5) < 1e-10
    assert abs(sum(a * a for a in ODWT.DEC_LO) - 1) < 1e-10
    assert abs(sum(a * b for a, b in zip(ODWT.DEC_LO, ODWT.DEC_HI))) < 1e-10
    x = torch.randn(3, 1, 8192, dtype=torch.float64)
    yl, yh = ODWT.dwt_db3_zero(x)
    assert yl.shape[-1] == 4098 and ODWT.dwt_out_len(4098) == 2051
    assert abs(float((yl.pow(2).sum() + yh.pow(2).sum()) / x.pow(2).sum()) - 1) < 1e-10


def test_dwt_against_independent_derivation():
    """Pin of row D3 that does not go through oracle/dwt.py's own arithmetic (the reference-shim goldens are
    circular for the DWT).  (1) The db3 taps from Daubechies' closed form for N = 3 (spectral factorisation of
    1 + 3y + 6y^2: Daubechies, "Ten Lectures on Wavelets", table 6.1) must reproduce the PyWavelets table used by the
    oracle and the CUDA kernel.  (2) PyWavelets' documented definition of the zero-mode analysis step -- the FULL linear
    convolution with the decomposition filter, keeping the odd samples, floor((N + 5) / 2) coefficients -- evaluated
    with numpy.convolve must equal the oracle's strided-conv1d-on-a-padded-signal formulation, for even and odd N."""
    s10 = 10 ** 0.5
    r = (5 + 2 * s10) ** 0.5
    rec_lo = np.array([1 + s10 + r, 5 + s10 + 3 * r, 10 - 2 * s10 + 2 * r, 10 - 2 * s10 - 2 * r, 5 + s10 - 3 * r,
                       1 + s10 - r]) / (16 * 2 ** 0.5)
    dec_lo = rec_lo[::-1]
    dec_hi = np.array([(-1) ** (k + 1) * rec_lo[k] for k in range(6)])     # quadrature mirror of the scaling filter
    # (PyWavelets tabulates the taps to ~1e-11; far below fp32 resolution)
    assert np.abs(dec_lo - np.array(ODWT.DEC_LO)).max() < 1e-10
    assert np.abs(dec_hi - np.array(ODWT.DEC_HI)).max() < 1e-10
    rng = np.random.default_rng(3)
    for n in (8192, 4098, 2051, 17, 6, 5):
        x = rng.standard_normal(n)
        want_lo = np.convolve(x, dec_lo)[1::2]
        want_hi = np.convolve(x, dec_hi)[1::2]
        yl, yh = ODWT.dwt_db3_zero(torch.from_numpy(x).view(1, 1, n))
        assert want_lo.shape[0] == (n + 5) // 2 == yl.shape[-1]
        assert np.abs(yl.numpy().ravel() - want_lo).max() < 1e-9, n
        assert np.abs(yh.numpy().ravel() - want_hi).max() < 1e-9, n


def test_mel_basis_matches_torchaudio():
    ta = pytest.importorskip("torchaudio")
    for sr, n_fft, n_mels, fmin, fmax in [(22050, 1024, 80, 80, 7600), (24000, 1024, 80, 0, 8000), (16000, 2048, 80, 0, 8000)]:
        ours = melbasis.mel_filterbank(sr, n_fft, n_mels, fmin, fmax)
        ref = ta.functional.melscale_fbanks(n_fft // 2 + 1, fmin, fmax, n_mels, sr, norm="slaney", mel_scale="slaney").T.numpy()
        assert np.abs(ours - ref).max() < 1e-6


@pytest.mark.parametrize("name", ["gen_small_nsf_causal", "gen_small_nsf_noncausal"])
def test_nsf_generator_matches_reference(golden, name):
    """the neural-source-filter generator variant (sine + noise excitation, per-stage strided
    ``source_downs``), same RNG seed -> same random phases / noise as the reference run."""
    g = golden(name)
    with torch.no_grad():
        torch.manual_seed(int(g.arrays["rng_seed"]))
        y = O.generator_forward(g.group("sd/"), g.t("x"), **g.cfg)
    assert y.shape == g.t("y").shape
    assert float((y - g.t("y")).abs().max()) < 2e-6
