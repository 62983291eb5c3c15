#!/bin/bash
mkdir -p gpurun_out; S=gpurun_out/summary30.txt; rm -f $S
timeout 600 python -m pytest tests -m gpu -q --tb=short 2>&1 > gpurun_out/tests30.log; echo "tests rc=$?" >> $S
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/tests30.log | cut -c1-300 | head -30 >> $S
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"conv_tc_kernel|wgrad_tc_kernel" --launch-skip 30 --launch-count 15 -o gpurun_out/r01_layers_run30 -f python scripts/layer_bench.py --iters 1 --only gen_128_128_k11,gen_32_32_k7,mpd_1024_1024_k5_p3 > gpurun_out/ncu30.log 2>&1; echo "ncu rc=$?" >> $S
cat $S; tail -n 12 gpurun_out/ncu30.log | cut -c1-200
