#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sambert.py -m gpu -q 2>&1 | tail -n 60 > gpurun_out/sambert_tests.log; echo "tests rc=${PIPESTATUS[0]}"
grep -E "passed|failed|FAILED|Error|assert" gpurun_out/sambert_tests.log | cut -c1-300 | head -20
timeout 300 python scripts/sambert_c4.py --steps 10 --prof > gpurun_out/sambert_c4.log 2>&1; echo "c4 rc=$?"
head -n 1 gpurun_out/sambert_c4.log; grep -E "kt::|cudnn|RNN_|Self CUDA time" gpurun_out/sambert_c4.log | cut -c1-70,130-240 | head -30
timeout 300 python scripts/layer_bench.py --iters 20 > gpurun_out/layer_bench.log 2>&1; echo "lb rc=$?"; cat gpurun_out/layer_bench.log | cut -c1-200
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"conv_tc_kernel|wgrad_tc_kernel" --launch-skip 12 --launch-count 3 -o gpurun_out/r01_mpd1024_p3 -f python scripts/layer_bench.py --iters 2 --only mpd_1024_1024_k5_p3 > gpurun_out/ncu_mpd.log 2>&1; echo "ncu rc=$?"; tail -n 5 gpurun_out/ncu_mpd.log
