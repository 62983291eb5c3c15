"""In-situ cost of the parts of one C2 train step (development aid; the ablated variants compute WRONG results):
graph-replay step time with one part removed at a time, plus the split of the plain step into its four segments."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import kantts_b200 as K
from kantts_b200 import ops, train

dev = torch.device("cuda", 0)
VARIANTS = sys.argv[1:] or ["base", "segments", "no_adam", "no_wgrad", "no_prepare", "no_wgrad_streams", "flags256", "flags512", "flags1024", "flags2048", "flags3840"]


def build():
    torch.manual_seed(1234)
    model, opt, sched = K.hifigan_model_builder(bench.CONFIG, dev)
    crit = K.criterion_builder(bench.CONFIG, dev)
    step = K.GanStep(model, opt, sched, crit, bench.CONFIG, cuda_graph=True)
    y, x = bench.synth_batch(bench.B_PER_GPU, 1234)
    return step, y.to(dev), x.to(dev)


def timed(step, y, x, n=10):
    for _ in range(6):
        step.step((y, x))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step.step((y, x))
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


orig_wb, orig_prep = ops._weight_backward, ops.prepare_weight
for var in VARIANTS:
    ops._weight_backward, ops.prepare_weight, ops._WGRAD_ASYNC = orig_wb, orig_prep, True
    if var == "no_wgrad":
        ops._weight_backward = lambda *a, **k: (None, None, None)
    elif var == "no_prepare":
        def prep(cache, spec, v, g):
            if cache.w_fwd is not None and isinstance(v, torch.nn.Parameter):
                return cache
            return orig_prep(cache, spec, v, g)
        ops.prepare_weight = prep
    elif var == "no_wgrad_streams":
        ops._WGRAD_ASYNC = False
    from kantts_b200 import _lib
    _lib.load().kt_debug_set_flags(int(var[5:]) if var.startswith("flags") else 0)
    step, y, x = build()
    if var == "no_adam":
        step._seg_gopt = lambda: None
        step._seg_dopt = lambda: None
    if var == "segments":
        timed(step, y, x, 2)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        acc = [0.0] * 4
        n = 10
        for _ in range(n):
            g1, g2 = step._graphs
            ev[0].record(); g1.replay()
            ev[1].record(); step._seg_gopt()
            ev[2].record(); g2.replay()
            ev[3].record(); step._seg_dopt()
            ev[4].record()
            torch.cuda.synchronize()
            for i in range(4):
                acc[i] += ev[i].elapsed_time(ev[i + 1]) / n
        print(f"segments: G graph {acc[0]:.2f}  G adam {acc[1]:.2f}  D graph {acc[2]:.2f}  D adam {acc[3]:.2f}  sum {sum(acc):.2f} ms", flush=True)
    else:
        print(f"{var:18s} {timed(step, y, x):8.2f} ms/step", flush=True)
    del step
    torch.cuda.empty_cache()
