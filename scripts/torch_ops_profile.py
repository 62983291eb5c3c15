"""Which torch (non-library) ops run inside one eager C2 train step, and from where (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
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
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True) as prof:
    step.step((y, x))
    torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=60, max_shapes_column_width=70))
print(prof.key_averages(group_by_stack_n=6).table(sort_by="cuda_time_total", row_limit=50, max_name_column_width=50, max_src_column_width=90))
