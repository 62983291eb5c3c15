#!/bin/bash
mkdir -p gpurun_out; S=gpurun_out/summary25.txt; rm -f $S
timeout 600 python -m pytest tests -m gpu -q --tb=short 2>&1 > gpurun_out/tests25.log; echo "tests rc=$?" >> $S
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/tests25.log | cut -c1-300 | head -30 >> $S
KANTTS_B200_STREAMS=0 timeout 300 python scripts/breakdown.py > gpurun_out/breakdown25.log 2>&1; echo "breakdown rc=$?" >> $S
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench25.log 2>&1; echo "bench rc=$?" >> $S
timeout 300 python scripts/cpu_bound.py > gpurun_out/cpu_bound25.log 2>&1
timeout 900 python scripts/graph_probe.py > gpurun_out/graph_probe25.log 2>&1; echo "probe rc=$?" >> $S
cat $S; tail -n 1 gpurun_out/bench25.log | cut -c1-300; head -n 3 gpurun_out/cpu_bound25.log; head -n 8 gpurun_out/breakdown25.log; cat gpurun_out/graph_probe25.log | cut -c1-250
