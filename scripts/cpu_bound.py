"""Is the eager step CPU- (launch-) bound?  Compare host enqueue time with device time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import kantts_b200 as K

dev = torch.device("cuda", 0)
torch.manual_seed(1234)
model, opt, sched = K.hifigan_model_builder(bench.CONFIG, dev)
crit = K.criterion_builder(bench.CONFIG, dev)
step = K.GanStep(model, opt, sched, crit, bench.CONFIG)
y, x = bench.synth_batch(bench.B_PER_GPU, 1234)
y, x = y.to(dev), x.to(dev)
for _ in range(3):
    step.step((y, x))
torch.cuda.synchronize()
enq, tot = [], []
for _ in range(5):
    t0 = time.perf_counter()
    step.step((y, x))
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    enq.append((t1 - t0) * 1e3); tot.append((t2 - t0) * 1e3)
print("host enqueue ms/step:", [round(v, 1) for v in enq])
print("wall incl. sync ms/step:", [round(v, 1) for v in tot])
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); step.step((y, x)); pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(45)
pstats.Stats(pr).sort_stats("cumulative").print_stats(40)
