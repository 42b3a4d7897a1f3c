#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/timeline.py --out gpurun_out/timeline10 > gpurun_out/timeline10.log 2>&1; echo "timeline exit $?"; tail -n 75 gpurun_out/timeline10.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel<1, 1, 256" -s 40 -c 2 -o gpurun_out/prof_k5_s10 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full10.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out/*.ncu-rep
