#!/bin/bash
mkdir -p gpurun_out; S=gpurun_out/summary29.txt; rm -f $S
timeout 600 python -m pytest tests -m gpu -q --tb=short 2>&1 > gpurun_out/tests29.log; echo "tests rc=$?" >> $S
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/tests29.log | cut -c1-300 | head -30 >> $S
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench29.log 2>&1; echo "bench rc=$?" >> $S
timeout 600 python bench.py --launch eager --steps 10 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/bench29e.log 2>&1; echo "bench eager rc=$?" >> $S
timeout 300 python scripts/sambert_medium_debug.py > gpurun_out/sambert_dbg29.log 2>&1; echo "sambert dbg rc=$?" >> $S
cat $S; tail -n 1 gpurun_out/bench29.log | cut -c1-330; tail -n 1 gpurun_out/bench29e.log | cut -c1-330; grep -E "==|e-0[123]" gpurun_out/sambert_dbg29.log | head -50
