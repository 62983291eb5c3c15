"""Finer probe of the g2 segment (generator Adam + no-grad G forward + discriminator fwd/bwd): which piece breaks
cudaGraphLaunch at full size?  One subprocess per variant."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANTS = ["adam_only", "dphase_noadam", "g2_ownpool", "g2_nosched", "g1_then_dphase_noadam"]


def run(v):
    import faulthandler; faulthandler.enable()
    import torch
    import bench
    import kantts_b200 as K
    dev = torch.device("cuda", 0)
    torch.manual_seed(1234)
    y, x = bench.synth_batch(bench.B_PER_GPU, 1234)
    y, x = y.to(dev), x.to(dev)
    model, opt, sched = K.hifigan_model_builder(bench.CONFIG, dev, capturable=True)
    crit = K.criterion_builder(bench.CONFIG, dev)
    step = K.GanStep(model, opt, sched, crit, bench.CONFIG)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            step.step((y, x))
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    step._log = {}
    graphs = []

    def cap(fn, pool=None):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, pool=pool):
            fn()
        graphs.append(g)
        return g

    if v == "adam_only":
        def fn():
            step.optimizer["generator"].step()
            step.scheduler["generator"].step()
        cap(fn)
    elif v == "dphase_noadam":
        step._g_active = lambda: False
        cap(lambda: step._seg_gopt_discriminator(y, x))
    elif v == "g2_ownpool":
        cap(lambda: step._seg_gopt_discriminator(y, x))
    elif v == "g2_nosched":
        class _NoSched:
            def step(self): pass
        step.scheduler["generator"] = _NoSched()
        cap(lambda: step._seg_gopt_discriminator(y, x))
    elif v == "g1_then_dphase_noadam":
        g1 = cap(lambda: step._seg_generator(y, x))
        step._g_active = lambda: False
        cap(lambda: step._seg_gopt_discriminator(y, x), pool=g1.pool())
    print(v, "captured", flush=True)
    for _ in range(3):
        for g in graphs:
            g.replay()
    torch.cuda.synchronize()
    print(v, "OK")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1])
    else:
        for v in VARIANTS:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), v], capture_output=True, text=True, timeout=240)
            tail = (r.stdout + r.stderr).strip().splitlines()
            keep = [l for l in tail if v in l or "Error" in l or "error" in l or "File" in l][-6:]
            print(f"=== {v}: rc={r.returncode}")
            for l in keep:
                print("   ", l[:250])
