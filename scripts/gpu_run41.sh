#!/bin/bash
mkdir -p gpurun_out; S=gpurun_out/summary41.txt; rm -f $S
timeout 600 python -m pytest tests -m gpu -q --tb=short 2>&1 > gpurun_out/tests41.log; echo "tests rc=$?" >> $S
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/tests41.log | cut -c1-300 | head -30 >> $S
timeout 300 python scripts/tc_trace.py > gpurun_out/tc_trace41.log 2>&1; echo "trace rc=$?" >> $S
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench41.log 2>&1; echo "bench rc=$?" >> $S
cat $S; tail -n 1 gpurun_out/bench41.log | cut -c1-330; grep -E "===|tile [0-1]:" gpurun_out/tc_trace41.log | cut -c1-200
