import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import kantts_b200 as K
from conftest import Golden, rel_l2
g = Golden("msd_small")
m = K.MultiScaleDiscriminator(**g.cfg)
m.load_state_dict(g.group("sd/"), strict=True)
m = m.to("cuda:0").train()
y = g.t("y").to("cuda:0")
outs, fmaps = m(y)
for i in range(3):
    print("scale", i, "out err", float((outs[i].cpu() - g.t(f"out{i}")).abs().max()))
    for l, f in enumerate(fmaps[i]):
        print("   fmap", l, tuple(f.shape), "rel", rel_l2(f.detach().cpu(), g.t(f"fmap{i}_{l}")))
sd = m.state_dict()
for k, v in g.group("after/").items():
    print(k, rel_l2(sd[k].cpu(), v))
