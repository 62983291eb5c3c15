"""Which part of the full-size step breaks cudaGraphLaunch?  Each variant runs in its own subprocess
(a segfault must not take the probe down).  python scripts/graph_probe.py [variant]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANTS = ["gen_fwd", "gen_fwdbwd", "mpd_fwdbwd", "msd_fwdbwd", "mpd_fwdbwd_serial", "gen_fwdbwd_serial", "step_serial", "step_bigstack"]


def run(v):
    import faulthandler; faulthandler.enable()
    if v.endswith("_serial"):
        os.environ["KANTTS_B200_STREAMS"] = "0"
    import torch, threading
    import bench
    import kantts_b200 as K
    dev = torch.device("cuda", 0)
    torch.manual_seed(1234)
    y, x = bench.synth_batch(bench.B_PER_GPU, 1234)
    y, x = y.to(dev), x.to(dev)
    if v.startswith("step"):
        model, opt, sched = K.hifigan_model_builder(bench.CONFIG, dev, capturable=True)
        crit = K.criterion_builder(bench.CONFIG, dev)
        step = K.GanStep(model, opt, sched, crit, bench.CONFIG, cuda_graph=True)
        def body():
            for i in range(6):
                log = step.step((y, x))
            torch.cuda.synchronize()
            print(v, "OK", {k: float(t) for k, t in log.items() if torch.is_tensor(t)})
        if v == "step_bigstack":
            threading.stack_size(1 << 30)
            t = threading.Thread(target=body); t.start(); t.join()
        else:
            body()
        return
    if v.startswith("gen"):
        m = K.Generator(**bench.G_PARAMS).to(dev)
        inp = x
    elif v.startswith("mpd"):
        m = K.MultiPeriodDiscriminator(**bench.MPD_PARAMS).to(dev)
        inp = y
    else:
        m = K.MultiScaleDiscriminator(**bench.MSD_PARAMS).to(dev)
        inp = y
    fg = K.train.FlatGrads(m)
    def fwdbwd():
        out = m(inp)
        outs = out[0] if isinstance(out, tuple) else [out]
        if "fwdbwd" in v:
            loss = sum((o * o).mean() for o in outs)
            fg.zero()
            loss.backward()
            K.hifigan.join_side_streams(dev)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            if v == "gen_fwd":
                with torch.no_grad():
                    fwdbwd()
            else:
                fwdbwd()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        if v == "gen_fwd":
            with torch.no_grad():
                fwdbwd()
        else:
            fwdbwd()
    print(v, "captured", flush=True)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    print(v, "OK", float(fg.flat.abs().sum()))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1])
    else:
        for v in VARIANTS:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), v], capture_output=True, text=True, timeout=240)
            tail = (r.stdout + r.stderr).strip().splitlines()
            keep = [l for l in tail if v in l or "Error" in l or "error" in l or "File" in l][-8:]
            print(f"=== {v}: rc={r.returncode}")
            for l in keep:
                print("   ", l[:250])
