#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
T=tests/test_gpu_parity.py
timeout 900 python -m pytest $T -m gpu -q > gpurun_out/t_all.log 2>&1
echo "pytest gpu rc=$?" >> gpurun_out/summary.txt
timeout 120 python scripts/debug_msd.py > gpurun_out/debug_msd.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/summary.txt
KANTTS_B200_PATH=ffma timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ffma.log 2>&1
echo "bench ffma rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_auto.log 2>&1
echo "bench auto rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python bench.py --ncu --steps 1 > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc_kernel -s 40 -c 3 -o gpurun_out/prof_tc python bench.py --ncu --steps 1 > gpurun_out/ncu_tc.log 2>&1
echo "ncu tc rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -n 15 gpurun_out/t_all.log; cat gpurun_out/debug_msd.log | tail -n 40; tail -n 3 gpurun_out/smoke.log; tail -n 2 gpurun_out/bench_ffma.log gpurun_out/bench_auto.log
