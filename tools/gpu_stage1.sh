#!/bin/bash
# First GPU bring-up: staged so a trap in a tensor-core kernel cannot take the simpler stages down with it.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
PT="python -m pytest -q -p no:cacheprovider --timeout=600 -m gpu"
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n 12 gpurun_out/$name.log; }
run ops_simple   $PT tests/test_gpu_ops.py -k "not conv_rows and not wgrad_rows"
run gemm_fp32    $PT tests/test_gpu_ops.py -k "(conv_rows or wgrad_rows) and fp32"
run dec_fp32     $PT tests/test_gpu_decoder.py -k "fp32"
run gemm_bf16    $PT tests/test_gpu_ops.py -k "(conv_rows or wgrad_rows) and bf16 and not bf16x3"
run gemm_bf16x3  $PT tests/test_gpu_ops.py -k "(conv_rows or wgrad_rows) and bf16x3"
run dec_tc       $PT tests/test_gpu_decoder.py -k "not fp32"
run smoke        python __graft_entry__.py --smoke
