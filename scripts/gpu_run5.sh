#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 300 python scripts/cpu_bound.py 2>&1 | head -n 4 > gpurun_out/cpu_bound.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.log 2>&1
echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python bench.py --ncu --steps 1 > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc_kernel -s 150 -c 4 -o gpurun_out/prof_conv_tc python bench.py --ncu --steps 1 > gpurun_out/ncu_tc.log 2>&1
echo "ncu conv_tc rc=$?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:wgrad_tc_kernel -s 40 -c 3 -o gpurun_out/prof_wgrad_tc python bench.py --ncu --steps 1 > gpurun_out/ncu_wg.log 2>&1
echo "ncu wgrad_tc rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt gpurun_out/cpu_bound.log; tail -n 1 gpurun_out/bench_final.log | cut -c1-600
