"""Loss trajectories of the small train-step fixture under the launch-mode toggles (debug aid)."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
CONFIGS = {
    "base":        dict(KANTTS_B200_STREAMS="0", KANTTS_B200_WGRAD_STREAMS="0", KANTTS_B200_PREFETCH="0", GRAPH="0"),
    "streams":     dict(KANTTS_B200_STREAMS="1", KANTTS_B200_WGRAD_STREAMS="0", KANTTS_B200_PREFETCH="0", GRAPH="0"),
    "wg":          dict(KANTTS_B200_STREAMS="1", KANTTS_B200_WGRAD_STREAMS="1", KANTTS_B200_PREFETCH="0", GRAPH="0"),
    "pf":          dict(KANTTS_B200_STREAMS="1", KANTTS_B200_WGRAD_STREAMS="0", KANTTS_B200_PREFETCH="1", GRAPH="0"),
    "all":         dict(KANTTS_B200_STREAMS="1", KANTTS_B200_WGRAD_STREAMS="1", KANTTS_B200_PREFETCH="1", GRAPH="0"),
    "graph_base":  dict(KANTTS_B200_STREAMS="0", KANTTS_B200_WGRAD_STREAMS="0", KANTTS_B200_PREFETCH="0", GRAPH="1"),
    "graph_wg":    dict(KANTTS_B200_STREAMS="1", KANTTS_B200_WGRAD_STREAMS="1", KANTTS_B200_PREFETCH="0", GRAPH="1"),
    "graph_pf":    dict(KANTTS_B200_STREAMS="1", KANTTS_B200_WGRAD_STREAMS="0", KANTTS_B200_PREFETCH="1", GRAPH="1"),
    "graph_all":   dict(KANTTS_B200_STREAMS="1", KANTTS_B200_WGRAD_STREAMS="1", KANTTS_B200_PREFETCH="1", GRAPH="1"),
}


def run():
    import torch
    import kantts_b200 as K
    from conftest import Golden
    from test_gpu_parity import _small_config
    g = Golden("trainstep_small")
    cfg = _small_config(g)
    torch.manual_seed(0)
    model, opt, sched = K.hifigan_model_builder(cfg, "cuda")
    model["generator"].load_state_dict(g.group("before/g/"))
    model["discriminator"]["MultiScaleDiscriminator"].load_state_dict(g.group("before/msd/"))
    model["discriminator"]["MultiPeriodDiscriminator"].load_state_dict(g.group("before/mpd/"))
    crit = K.criterion_builder(cfg, "cuda")
    step = K.GanStep(model, opt, sched, crit, cfg, cuda_graph=os.environ["GRAPH"] == "1", graph_warmup=2)
    y, x = g.t("y").to("cuda"), g.t("x").to("cuda")
    out = []
    for i in range(5):
        log = K.train.losses_to_float(step.step((y, x)))
        out.append([round(log[k], 5) for k in ("mel_loss", "adversarial_loss", "feature_matching_loss", "real_loss", "fake_loss")])
    print("TRAJ " + json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run()
    else:
        for name, env in CONFIGS.items():
            e = dict(os.environ); e.update(env)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "run"], env=e, capture_output=True, text=True, timeout=200)
            line = [l for l in r.stdout.splitlines() if l.startswith("TRAJ ")]
            print(f"{name:11s} rc={r.returncode} ", line[0][5:] if line else (r.stdout + r.stderr)[-600:])
