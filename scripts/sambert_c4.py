"""Time the SAM-BERT C4 train step (BASELINE configs[3]: batch 32, 256 symbols, 768 frames) on one GPU.
Usage: python scripts/sambert_c4.py [--steps 10] [--eval] [--prof]"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import kantts_b200  # noqa: E402
from kantts_b200 import ops, sambert  # noqa: E402
from golden.make_batch import make_c4_batch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--eval", action="store_true")
    ap.add_argument("--prof", action="store_true")
    ap.add_argument("--batch", type=int, default=32)
    a = ap.parse_args()
    dev = "cuda"
    cfg = kantts_b200.sambert_24k_config()
    torch.manual_seed(1234)
    model = sambert.KanTtsSAMBERT(cfg).to(dev)
    model.eval() if a.eval else model.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.98), eps=1e-9)
    sch = kantts_b200.train.NoamLR(opt, warmup_steps=4000)
    step = kantts_b200.SambertStep(model, opt, sch, {"MelReconLoss": sambert.MelReconLoss(),
                                                     "ProsodyReconLoss": sambert.ProsodyReconLoss()})
    batch = {k: v.to(dev) for k, v in make_c4_batch(cfg, torch.Generator().manual_seed(1234), B=a.batch).items()}
    ctx = torch.backends.cudnn.flags(enabled=not a.eval)
    with ctx:
        for _ in range(a.warmup):
            out = step.step(batch)
        torch.cuda.synchronize()
        n0 = ops.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(a.steps):
            out = step.step(batch)
        e1.record()
        t_host = (time.perf_counter() - t0) / a.steps * 1e3
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        frames = a.batch * 768
        print(f"sambert C4 {'eval' if a.eval else 'train'} step: {ms:.2f} ms/step (host enqueue {t_host:.2f} ms), "
              f"{frames / ms * 1e3:.0f} mel frames/s, library launches/step {(ops.launch_count() - n0) / a.steps:.0f}, "
              f"loss {float(out['TotalLoss']):.4f}")
        if a.prof:
            from torch.profiler import profile, ProfilerActivity
            with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
                step.step(batch)
                torch.cuda.synchronize()
            print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))


if __name__ == "__main__":
    main()
