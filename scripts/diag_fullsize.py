"""Diagnostics for the full-size parity failures (development aid): (1) MSD scale-0 first-layer weight_orig gradient, pair vs
two-call vs oracle; (2) exact (FFMA) vs bf16x3 path gradients of G under a mel-only loss and under a discriminator loss."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import kantts_b200 as K
from kantts_b200 import ops
from oracle import hifigan as O

dev = torch.device("cuda")
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
torch.manual_seed(1234)
msd = K.MultiScaleDiscriminator(**bench.MSD_PARAMS)
sd = {k: v.detach().clone() for k, v in msd.state_dict().items()}
y, x = bench.synth_batch(16, 1234)
y2 = (0.1 * torch.randn(16, 1, 8192)).clamp(-1, 1)

# ---- (1) oracle: D-phase style loss on (y, y2)
leaf = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not (k.endswith("weight_u") or (k.endswith("weight_v") and k[:-8] + "weight_orig" in sd) or "meanpools" in k) else v.clone()) for k, v in sd.items()}
torch.set_num_threads(32)
pa, _ = O.msd_forward(leaf, y, True, **bench.MSD_PARAMS)
pb, _ = O.msd_forward(leaf, y2, True, **bench.MSD_PARAMS)
(sum(((o - 1) ** 2).mean() for o in pa) + sum((o ** 2).mean() for o in pb)).backward()
for mode in ("two", "pair"):
    m = K.MultiScaleDiscriminator(**bench.MSD_PARAMS)
    m.load_state_dict(sd)
    m = m.to(dev).train()
    if mode == "two":
        oa, _ = m(y.to(dev)); ob, _ = m(y2.to(dev))
    else:
        (oa, _), (ob, _) = m.forward_pair(y.to(dev), y2.to(dev))
    (sum(((o - 1) ** 2).mean() for o in oa) + sum((o ** 2).mean() for o in ob)).backward()
    K.hifigan.join_side_streams(dev); torch.cuda.synchronize()
    errs = sorted(((rel(p.grad.cpu(), leaf[k].grad), k) for k, p in m.named_parameters()), reverse=True)
    print(f"MSD {mode}: worst", [(f"{e:.2e}", k) for e, k in errs[:4]], "median %.2e" % errs[len(errs) // 2][0], flush=True)

# ---- (2) G gradients: FFMA vs TC, mel-only loss and MPD-adversarial loss
torch.manual_seed(1234)
G = K.Generator(**bench.G_PARAMS).to(dev)
mpd = K.MultiPeriodDiscriminator(**bench.MPD_PARAMS).to(dev)
msd_d = K.MultiScaleDiscriminator(**bench.MSD_PARAMS).to(dev)
mel = K.MelSpectrogramLoss(**bench.LOSS["mel_loss"]["params"]).to(dev)
xg, yg = x.to(dev), y.to(dev)
# oracle: mel-only loss gradients of G
gsd = {k: v.detach().cpu().clone() for k, v in G.state_dict().items()}
gleaf = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in gsd.items()}
yo = O.generator_forward(gleaf, x, **bench.G_PARAMS)
lo = O.mel_spectrogram_loss(yo, y, **bench.LOSS["mel_loss"]["params"])
lo.backward()
oracle_g = {k: v.grad for k, v in gleaf.items() if v.grad is not None}
res = {}
for loss_kind in ("mel", "mpd", "msd"):
    for ffma in (False, True):
        ops.set_force_ffma(ffma)
        for p in G.parameters():
            p.grad = None
        y_ = G(xg)
        if loss_kind == "mel":
            loss = mel(y_, yg)
        else:
            d = mpd if loss_kind == "mpd" else msd_d
            for q in d.parameters():
                q.requires_grad_(False)
            outs, _ = d(y_)
            loss = sum(((o - 1) ** 2).mean() for o in outs)
        loss.backward()
        K.hifigan.join_side_streams(dev); torch.cuda.synchronize()
        res[(loss_kind, ffma)] = ({k: p.grad.clone() for k, p in G.named_parameters()}, y_.detach().clone(), float(loss))
    ops.set_force_ffma(False)
    a, b = res[(loss_kind, False)], res[(loss_kind, True)]
    errs = sorted(((rel(a[0][k], b[0][k]), k) for k in a[0]), reverse=True)
    if loss_kind == "mel":
        for tag, rr in (("TC", a), ("FFMA", b)):
            e2 = sorted(((rel(rr[0][k].cpu(), oracle_g[k]), k) for k in rr[0]), reverse=True)
            print(f"G grads, loss=mel: {tag} vs ORACLE worst", [(f"{e:.2e}", k) for e, k in e2[:3]], "median %.2e" % e2[len(e2) // 2][0],
                  "| y rel %.2e loss %.6f vs %.6f" % (rel(rr[1].cpu(), yo.detach()), rr[2], float(lo)), flush=True)
    print(f"G grads, loss={loss_kind}: TC vs FFMA worst", [(f"{e:.2e}", k) for e, k in errs[:3]], "median %.2e" % errs[len(errs) // 2][0],
          "| y rel %.2e loss %.6f vs %.6f" % (rel(a[1], b[1]), a[2], b[2]), flush=True)


# ---- (3) FFMA path, mel-only loss, vs ORACLE under stream variations
from kantts_b200 import hifigan
for par, wga in ((True, True), (False, True), (True, False), (False, False)):
    hifigan._PARALLEL_STREAMS, ops._WGRAD_ASYNC = par, wga
    ops.set_force_ffma(True)
    for p_ in G.parameters():
        p_.grad = None
    y_ = G(xg)
    mel(y_, yg).backward()
    K.hifigan.join_side_streams(dev); torch.cuda.synchronize()
    ops.set_force_ffma(False)
    e2 = sorted(((rel(p_.grad.cpu(), oracle_g[k]), k) for k, p_ in G.named_parameters()), reverse=True)
    print(f"FFMA mel-only vs ORACLE, parallel streams {par}, wgrad streams {wga}: worst", [(f"{e:.2e}", k) for e, k in e2[:2]],
          "median %.2e" % e2[len(e2) // 2][0], flush=True)
hifigan._PARALLEL_STREAMS, ops._WGRAD_ASYNC = True, True
