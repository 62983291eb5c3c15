#!/bin/bash
# Re-entry validation of HEAD: full GPU suite, bench line, ncu launch list, conv_tc traffic, graph-crash diagnosis.
mkdir -p gpurun_out; S=gpurun_out/summary22.txt; rm -f $S
t0=$(date +%s)
timeout 1300 python -m pytest tests -m gpu -q --durations=12 2>&1 | tail -n 70 > gpurun_out/tests22.log; echo "tests rc=${PIPESTATUS[0]} t=$(( $(date +%s)-t0 ))s" >> $S
grep -E "passed|failed|FAILED|Error" gpurun_out/tests22.log | cut -c1-300 | head -20 >> $S
t0=$(date +%s)
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench22.log 2>&1; echo "bench rc=$? t=$(( $(date +%s)-t0 ))s" >> $S
timeout 300 python scripts/cpu_bound.py 2>&1 | head -n 40 > gpurun_out/cpu_bound22.log
t0=$(date +%s)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches22.csv python bench.py --ncu --steps 1 > gpurun_out/ncu_list22.log 2>&1
echo "ncu list rc=$? t=$(( $(date +%s)-t0 ))s" >> $S
t0=$(date +%s)
timeout 300 python bench.py --graph --steps 5 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/graph22.log 2>&1; echo "graph rc=$? t=$(( $(date +%s)-t0 ))s" >> $S
tail -n 25 gpurun_out/graph22.log | cut -c1-400 >> $S
cat $S; tail -n 1 gpurun_out/bench22.log | cut -c1-1500; head -n 3 gpurun_out/cpu_bound22.log
