#!/bin/bash
# Round-1 final validation + profile evidence.
mkdir -p gpurun_out; S=gpurun_out/summary38.txt; rm -f $S
timeout 600 python -m pytest tests -m gpu -q --tb=short 2>&1 > gpurun_out/tests38.log; echo "tests rc=$?" >> $S
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/tests38.log | cut -c1-300 | head -30 >> $S
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke38.log 2>&1; echo "smoke rc=$?" >> $S
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench38.log 2>&1; echo "bench rc=$?" >> $S
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench38_ref.log 2>&1; echo "bench ref rc=$?" >> $S
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches38.csv python bench.py --ncu --steps 1 > gpurun_out/ncu_list38.log 2>&1; echo "ncu list rc=$?" >> $S
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc_kernel -s 200 -c 4 -o gpurun_out/r01_conv_tc_run38 -f python bench.py --ncu --steps 1 > gpurun_out/ncu_tc38.log 2>&1; echo "ncu conv_tc rc=$?" >> $S
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:wgrad_tc_kernel -s 60 -c 3 -o gpurun_out/r01_wgrad_tc_run38 -f python bench.py --ncu --steps 1 > gpurun_out/ncu_wg38.log 2>&1; echo "ncu wgrad rc=$?" >> $S
cat $S; tail -n 3 gpurun_out/smoke38.log; tail -n 1 gpurun_out/bench38.log | cut -c1-2500; tail -n 1 gpurun_out/bench38_ref.log | cut -c1-900
