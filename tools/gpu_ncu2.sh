#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 9 -c 14 -o gpurun_out/prof_r1b python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out/prof_r1b.ncu-rep
