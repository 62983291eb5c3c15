#!/bin/bash
mkdir -p gpurun_out; S=gpurun_out/summary27.txt; rm -f $S
timeout 300 python -X faulthandler bench.py --graph --steps 10 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/graph27.log 2>&1; echo "graph bench rc=$?" >> $S
KANTTS_B200_TEST_GRAPH=1 timeout 300 python -m pytest tests/test_gpu_graph.py -m gpu -q --tb=short 2>&1 | tail -n 15 > gpurun_out/graphtest27.log; echo "graph test rc=${PIPESTATUS[0]}" >> $S
cat $S; tail -n 12 gpurun_out/graph27.log | cut -c1-500; tail -n 6 gpurun_out/graphtest27.log | cut -c1-300
