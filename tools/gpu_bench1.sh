#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bf16 exit $?"; tail -c 3000 gpurun_out/bench_bf16.json
python bench.py --steps 5 --warmup 3 --precision bf16x3 --no-cpu-baseline > gpurun_out/bench_bf16x3.json 2> gpurun_out/bench_bf16x3.err; echo "x3 exit $?"
python bench.py --steps 3 --warmup 3 --precision fp32 --no-cpu-baseline > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; echo "fp32 exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 1 -c 3 -o gpurun_out/prof_r1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out | head -30
