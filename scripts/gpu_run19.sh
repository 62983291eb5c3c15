#!/bin/bash
# SAM-BERT first GPU pass: kernel + model parity, C4 timing.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sambert.py -m gpu -x -q 2>&1 | tail -n 40 > gpurun_out/sambert_tests.log; echo "tests rc=${PIPESTATUS[0]}"
tail -n 30 gpurun_out/sambert_tests.log | cut -c1-400
timeout 300 python scripts/sambert_c4.py --steps 10 --prof > gpurun_out/sambert_c4.log 2>&1; echo "c4 rc=$?"
head -n 45 gpurun_out/sambert_c4.log | cut -c1-260
