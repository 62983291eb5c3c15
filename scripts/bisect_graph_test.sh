#!/bin/bash
# development aid: which switch makes tests/test_gpu_graph.py fail?
for cfg in "KANTTS_B200_EPI_TMA=1 KANTTS_B200_FUSED_ADAM=1" "KANTTS_B200_EPI_TMA=0 KANTTS_B200_FUSED_ADAM=1" "KANTTS_B200_EPI_TMA=1 KANTTS_B200_FUSED_ADAM=0" "KANTTS_B200_EPI_TMA=0 KANTTS_B200_FUSED_ADAM=0"; do
  for rep in 1 2; do
    echo "== $cfg rep $rep"
    env $cfg timeout 300 python -m pytest tests/test_gpu_graph.py -m gpu -q -x 2>&1 | grep -E "passed|failed|AssertionError" | head -3
  done
done
