#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider --timeout=600 -m gpu"
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n 8 gpurun_out/$name.log; }
run lstm      $PT tests/test_gpu_decoder.py -k "context_lstm"
run ops       $PT tests/test_gpu_ops.py
run dec       $PT tests/test_gpu_decoder.py -k "not context_lstm"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench2_bf16.json 2> gpurun_out/bench2_bf16.err; echo "bf16 exit $?"; tail -c 2500 gpurun_out/bench2_bf16.json; tail -5 gpurun_out/bench2_bf16.err
python bench.py --steps 5 --warmup 3 --precision bf16x3 --no-cpu-baseline > gpurun_out/bench2_bf16x3.json 2> gpurun_out/bench2_bf16x3.err; echo "x3 exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch2.log 2>&1; echo "ncu launches exit $?"
