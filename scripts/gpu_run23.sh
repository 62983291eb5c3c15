#!/bin/bash
# pair-batched discriminators + in-kernel grad accumulation: tests, per-layer breakdown (serial streams), bench, graph segfault trace
mkdir -p gpurun_out; S=gpurun_out/summary23.txt; rm -f $S
timeout 600 python -m pytest tests -m gpu -q --tb=short 2>&1 > gpurun_out/tests23.log; echo "tests rc=$?" >> $S
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/tests23.log | cut -c1-300 | head -20 >> $S
KANTTS_B200_STREAMS=0 timeout 300 python scripts/breakdown.py > gpurun_out/breakdown23.log 2>&1; echo "breakdown rc=$?" >> $S
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench23.log 2>&1; echo "bench rc=$?" >> $S
timeout 300 python scripts/cpu_bound.py 2>&1 | head -n 3 > gpurun_out/cpu_bound23.log
timeout 300 python -X faulthandler bench.py --graph --steps 5 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/graph23.log 2>&1; echo "graph rc=$?" >> $S
cat $S; tail -n 1 gpurun_out/bench23.log | cut -c1-700; cat gpurun_out/cpu_bound23.log; head -n 30 gpurun_out/breakdown23.log; tail -n 40 gpurun_out/graph23.log | cut -c1-300
