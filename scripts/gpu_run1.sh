#!/bin/bash
# first GPU run: FFMA parity, tcgen05 descriptor variants, model-level parity, bench (ffma / auto)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
T=tests/test_gpu_parity.py
timeout 900 python -m pytest $T -m gpu -q -k "ffma_vs_oracle or discriminators or mel_and" > gpurun_out/t1_ffma.log 2>&1
echo "t1 rc=$?" >> gpurun_out/summary.txt
KT_TC_BASE_OFFSET=1 timeout 300 python -m pytest $T -m gpu -q -k "tcgen05_vs_oracle" > gpurun_out/t2_tc_bo1.log 2>&1
echo "t2 (base_offset=1) rc=$?" >> gpurun_out/summary.txt
KT_TC_BASE_OFFSET=0 timeout 300 python -m pytest $T -m gpu -q -k "tcgen05_vs_oracle" > gpurun_out/t3_tc_bo0.log 2>&1
echo "t3 (base_offset=0) rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest $T -m gpu -q -k "generator_matches or gan_train_step or c1_full or full_size" > gpurun_out/t4_models.log 2>&1
echo "t4 rc=$?" >> gpurun_out/summary.txt
KANTTS_B200_PATH=ffma timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ffma.log 2>&1
echo "bench ffma rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_auto.log 2>&1
echo "bench auto rc=$?" >> gpurun_out/summary.txt
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $T -m gpu -q -k "ffma_vs_oracle and (causal_dilated or period_strided or deconv_causal or upsample_conv or msd_grouped_strided or cin1)" > gpurun_out/t5_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -5 gpurun_out/t1_ffma.log gpurun_out/t2_tc_bo1.log gpurun_out/t3_tc_bo0.log gpurun_out/t4_models.log
tail -2 gpurun_out/bench_ffma.log gpurun_out/bench_auto.log
