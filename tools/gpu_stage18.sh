#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest tests/test_gpu_graph.py -q -x -p no:cacheprovider --timeout=600 -m gpu -k fp32"
echo "=== default"; timeout 300 $PT > gpurun_out/g18_default.log 2>&1; echo "exit $?"; grep -A12 "graph gradients differ" gpurun_out/g18_default.log | head -30
echo "=== PRE_W=0"; RADMMM_B200_PRE_W=0 timeout 300 $PT > gpurun_out/g18_prew0.log 2>&1; echo "exit $?"; grep -A12 "graph gradients differ" gpurun_out/g18_prew0.log | head -20
echo "=== LANES=0"; RADMMM_B200_LANES=0 timeout 300 $PT > gpurun_out/g18_lanes0.log 2>&1; echo "exit $?"; grep -A12 "graph gradients differ" gpurun_out/g18_lanes0.log | head -20
echo "=== SIDE=0"; RADMMM_B200_SIDE_STREAM=0 timeout 300 $PT > gpurun_out/g18_side0.log 2>&1; echo "exit $?"; grep -A12 "graph gradients differ" gpurun_out/g18_side0.log | head -20
echo "=== ops"; timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -p no:cacheprovider --timeout=600 -m gpu > gpurun_out/g18_ops.log 2>&1; echo "exit $?"; tail -3 gpurun_out/g18_ops.log
