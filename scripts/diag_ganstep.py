"""Which ingredient of GanStep breaks the gradient of MSD scale 0's first layer (weight_orig) at full size?  (development aid)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import kantts_b200 as K
from kantts_b200 import ops, hifigan
from oracle import hifigan as O

dev = torch.device("cuda")
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
torch.manual_seed(1234)
model, opt, sched = K.hifigan_model_builder(bench.CONFIG, dev)
sds = {"g": {k: v.detach().cpu().clone() for k, v in model["generator"].state_dict().items()},
       **{n: {k: v.detach().cpu().clone() for k, v in m.state_dict().items()} for n, m in model["discriminator"].items()}}
y, x = bench.synth_batch(16, 1234)
torch.set_num_threads(32)
gan = O.OracleGAN(sds["g"], {n: sds[n] for n in model["discriminator"]}, bench.G_PARAMS,
                  {"MultiScaleDiscriminator": bench.MSD_PARAMS, "MultiPeriodDiscriminator": bench.MPD_PARAMS}, bench.LOSS)
gan.train_step(y, x)
ref = {k: v.grad for k, v in gan.d["MultiScaleDiscriminator"].items() if v.grad is not None}
KEY = "discriminators.0.convs.0.0.weight_orig"
runs = []
for rep in range(3):
    runs += [("spectral on main #%d" % rep, {}, True, True, True), ("spectral on side #%d" % rep, {}, True, True, False)]
for name, kw, par, wga, som in runs:
    hifigan._SPECTRAL_ON_MAIN = som
    torch.manual_seed(1234)
    model, opt, sched = K.hifigan_model_builder(bench.CONFIG, dev)
    model["generator"].load_state_dict(sds["g"])
    for n, m in model["discriminator"].items():
        m.load_state_dict(sds[n])
    crit = K.criterion_builder(bench.CONFIG, dev)
    hifigan._PARALLEL_STREAMS, ops._WGRAD_ASYNC = par, wga
    step = K.GanStep(model, opt, sched, crit, bench.CONFIG, **kw)
    step.step((y.to(dev), x.to(dev)))
    torch.cuda.synchronize()
    hifigan._PARALLEL_STREAMS, ops._WGRAD_ASYNC = True, True
    hifigan._SPECTRAL_ON_MAIN = True
    msd = model["discriminator"]["MultiScaleDiscriminator"]
    errs = sorted(((rel(p.grad.cpu(), ref[k]), k) for k, p in msd.named_parameters()), reverse=True)
    g = dict(msd.named_parameters())[KEY].grad
    print(f"{name:26s}: {KEY} err {rel(g.cpu(), ref[KEY]):.3e} |g| {float(g.norm()):.3e} |ref| {float(ref[KEY].norm()):.3e}; worst "
          + str([(f"{e:.1e}", k) for e, k in errs[:3]]), flush=True)
