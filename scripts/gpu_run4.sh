#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
T=tests/test_gpu_parity.py
timeout 600 python -m pytest $T -m gpu -q -k "tcgen05_vs_oracle" > gpurun_out/t_tc.log 2>&1
echo "pytest tc rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests -m gpu -q -k "not tcgen05_vs_oracle" > gpurun_out/t_rest.log 2>&1
echo "pytest rest rc=$?" >> gpurun_out/summary.txt
timeout 600 python scripts/breakdown.py > gpurun_out/breakdown.log 2>&1
echo "breakdown rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_auto.log 2>&1
echo "bench auto rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; grep -E "passed|failed|^FAILED|^E  " gpurun_out/t_tc.log | cut -c1-300 | head -40; grep -E "passed|failed|^FAILED" gpurun_out/t_rest.log | head; head -n 40 gpurun_out/breakdown.log; tail -n 1 gpurun_out/bench_auto.log | cut -c1-400
