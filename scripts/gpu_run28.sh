#!/bin/bash
mkdir -p gpurun_out; S=gpurun_out/summary28.txt; rm -f $S
timeout 300 python -m pytest tests/test_gpu_graph.py -m gpu -q --tb=short 2>&1 | tail -n 15 > gpurun_out/graphtest28.log; echo "graph test rc=${PIPESTATUS[0]}" >> $S
timeout 600 python -X faulthandler bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench28.log 2>&1; echo "bench(auto) rc=$?" >> $S
cat $S; tail -n 3 gpurun_out/bench28.log | cut -c1-900; tail -n 6 gpurun_out/graphtest28.log | cut -c1-300
