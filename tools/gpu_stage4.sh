#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider --timeout=600 -m gpu"
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n 14 gpurun_out/$name.log; }
run newtests  $PT tests/test_gpu_decoder.py tests/test_gpu_ops.py -k "spline_flow_step or conv_attention_module or soft_attention"
