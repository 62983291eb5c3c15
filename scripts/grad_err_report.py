"""Per-tensor parameter-gradient error of the default (bf16x3) path vs the reference goldens (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import kantts_b200 as K
from conftest import Golden, rel_l2  # noqa

DEV = "cuda"
def report(name, cls, fwd):
    g = Golden(name)
    m = getattr(K, cls)(**g.cfg)
    m.load_state_dict(g.group("sd/"), strict=True)
    m = m.to(DEV).train()
    fwd(m, g)
    refg = g.group("grad/")
    errs = sorted(((rel_l2(p.grad.cpu(), refg[k]), k, tuple(p.shape), float(refg[k].norm())) for k, p in m.named_parameters()), reverse=True)
    print(f"== {name}: worst tensors (rel err, name, shape, |ref grad|)")
    for e in errs[:12]:
        print("   %.2e  %-50s %-18s %.3e" % e)
    print("   median %.2e" % errs[len(errs) // 2][0])

def gen_fwd(m, g):
    y = m(g.t("x").to(DEV))
    (y * g.t("r").to(DEV)).sum().backward()

def disc_fwd(m, g):
    outs, _ = m(g.t("y").to(DEV))
    sum((o * g.t(f"r{i}").to(DEV)).sum() for i, o in enumerate(outs)).backward()

report("gen_small_causal", "Generator", gen_fwd)
report("gen_small_noncausal", "Generator", gen_fwd)
report("mpd_small", "MultiPeriodDiscriminator", disc_fwd)
report("msd_small", "MultiScaleDiscriminator", disc_fwd)
