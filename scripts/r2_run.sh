#!/bin/bash
# round-2 GPU run helper: $1 = run tag; remaining args select legs (tests unverified trace layers bench ncu_list)
TAG=$1; shift
mkdir -p gpurun_out; S=gpurun_out/summary_$TAG.txt; rm -f $S
for leg in "$@"; do
  case $leg in
    tests) timeout 900 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/tests_$TAG.log 2>&1; echo "tests rc=$?" >> $S
           grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/tests_$TAG.log | cut -c1-300 | head -30 >> $S ;;
    unverified) KANTTS_B200_TEST_UNVERIFIED=1 timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -q --tb=short > gpurun_out/unverified_$TAG.log 2>&1; echo "unverified rc=$?" >> $S
           grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/unverified_$TAG.log | cut -c1-300 | head -30 >> $S ;;
    trace) timeout 300 python scripts/tc_trace.py > gpurun_out/tc_trace_$TAG.log 2>&1; echo "trace rc=$?" >> $S ;;
    layers) KANTTS_B200_WGRAD_STREAMS=0 timeout 300 python scripts/layer_bench.py > gpurun_out/layers_$TAG.log 2>&1; echo "layers rc=$?" >> $S ;;
    smoke) timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> $S ;;
    bench) timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.log 2>&1; echo "bench rc=$?" >> $S ;;
    breakdown) timeout 600 python scripts/breakdown.py > gpurun_out/breakdown_$TAG.log 2>&1; echo "breakdown rc=$?" >> $S ;;
    ncu_list) timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --ncu --steps 1 > gpurun_out/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?" >> $S ;;
    ncu_layer) # full capture (source counters) of conv_tc on one layer: NCU_LAYER=name NCU_SKIP=n
           timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s ${NCU_SKIP:-6} -c 1 -f -o gpurun_out/ncu_${NCU_LAYER:-gen_128_128_k11}_$TAG python scripts/layer_bench.py --only ${NCU_LAYER:-gen_128_128_k11} --iters 2 > gpurun_out/ncu_layer_$TAG.log 2>&1; echo "ncu_layer rc=$?" >> $S ;;
    graderr) timeout 300 python scripts/grad_err_report.py > gpurun_out/graderr_$TAG.log 2>&1; echo "graderr rc=$?" >> $S ;;
    torchgpu) timeout 600 python bench.py --impl torch_gpu --steps 5 --warmup 3 > gpurun_out/torchgpu_$TAG.log 2>&1; echo "torchgpu rc=$?" >> $S ;;
    rbtest) timeout 300 python scripts/rb_test.py > gpurun_out/rbtest_$TAG.log 2>&1; echo "rbtest rc=$?" >> $S; tail -3 gpurun_out/rbtest_$TAG.log >> $S ;;
    diag) timeout 600 python scripts/diag_fullsize.py > gpurun_out/diag_$TAG.log 2>&1; echo "diag rc=$?" >> $S; grep -E "MSD|G grads|FFMA mel" gpurun_out/diag_$TAG.log >> $S ;;
    diaglayers) timeout 600 python scripts/diag_layers.py > gpurun_out/diaglayers_$TAG.log 2>&1; echo "diaglayers rc=$?" >> $S ;;
    diagmsd) timeout 600 python scripts/diag_layers.py msd > gpurun_out/diagmsd_$TAG.log 2>&1; echo "diagmsd rc=$?" >> $S ;;
    diaggan) timeout 900 python scripts/diag_ganstep.py > gpurun_out/diaggan_$TAG.log 2>&1; echo "diaggan rc=$?" >> $S ;;
    benchside) timeout 600 python bench.py --workload c1 --steps 20 --warmup 5 > gpurun_out/bench_c1_$TAG.log 2>&1; echo "bench c1 rc=$?" >> $S
               timeout 900 python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/bench_c4_$TAG.log 2>&1; echo "bench c4 rc=$?" >> $S ;;
    ncu_wgrad) for LN in ${NCU_WG_LAYERS:-mpd_1024_1024_k5_p3}; do
           timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tma_kernel -s 4 -c 1 -f -o gpurun_out/ncu_wgrad_${LN}_$TAG python scripts/layer_bench.py --only $LN --iters 2 > gpurun_out/ncu_wgrad_$TAG.log 2>&1; echo "ncu_wgrad $LN rc=$?" >> $S; done ;;
    ncu_rb) timeout 600 ncu --set full --clock-control none --import-source on -k regex:resblock_tc_kernel -s 2 -c 1 -f -o gpurun_out/ncu_resblock_$TAG env RB_ONLY_BIG=1 python scripts/rb_test.py > gpurun_out/ncu_rb_$TAG.log 2>&1; echo "ncu_rb rc=$?" >> $S ;;
    ablayers) for F in 0 16 32 48 64 80 96 112; do timeout 200 python scripts/layer_bench.py --flags $F --only ${AB_LAYERS:-gen_128_128_k11,mpd_1024_1024_k5_p3,msd_1024_1024_k5,gen_32_32_k7,mpd_128_512_k5s3_p5} >> gpurun_out/ablayers_$TAG.log 2>&1; done; echo "ablayers rc=$?" >> $S ;;
    ncu_layers) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ncu_layers_$TAG.csv python scripts/layer_bench.py --iters 1 ${NCU_LAYERS:+--only $NCU_LAYERS} > gpurun_out/ncu_layers_$TAG.log 2>&1; echo "ncu_layers rc=$?" >> $S ;;
    torchops) timeout 600 python scripts/torch_ops_profile.py > gpurun_out/torchops_$TAG.log 2>&1; echo "torchops rc=$?" >> $S ;;
    ablate) timeout 900 python scripts/ablate_step.py > gpurun_out/ablate_$TAG.log 2>&1; echo "ablate rc=$?" >> $S; grep -E "ms" gpurun_out/ablate_$TAG.log >> $S ;;
    *) echo "unknown leg $leg" >> $S ;;
  esac
done
cat $S
