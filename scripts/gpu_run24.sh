#!/bin/bash
# producer groups / block-diagonal grouped tiles / C_in=1 kernels / upsample dgrad on tc: tests, breakdown, bench
mkdir -p gpurun_out; S=gpurun_out/summary24.txt; rm -f $S
timeout 600 python -m pytest tests -m gpu -q --tb=short 2>&1 > gpurun_out/tests24.log; echo "tests rc=$?" >> $S
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/tests24.log | cut -c1-300 | head -30 >> $S
KANTTS_B200_STREAMS=0 timeout 300 python scripts/breakdown.py > gpurun_out/breakdown24.log 2>&1; echo "breakdown rc=$?" >> $S
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench24.log 2>&1; echo "bench rc=$?" >> $S
timeout 300 python scripts/cpu_bound.py 2>&1 | head -n 3 > gpurun_out/cpu_bound24.log
cat $S; tail -n 1 gpurun_out/bench24.log | cut -c1-400; cat gpurun_out/cpu_bound24.log; head -n 12 gpurun_out/breakdown24.log
