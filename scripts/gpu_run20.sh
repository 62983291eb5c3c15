#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_sambert.py -m gpu -q 2>&1 | tail -n 60 > gpurun_out/sambert_tests.log; echo "tests rc=${PIPESTATUS[0]}"
grep -E "passed|failed|FAILED|Error|assert" gpurun_out/sambert_tests.log | cut -c1-300 | head -40
