#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest tests/test_gpu_graph.py -q -x -p no:cacheprovider --timeout=600 -m gpu -k fp32"
for v in skip0 main rec; do
echo "=== PRE_W=$v"; RADMMM_B200_PRE_W=$v timeout 300 $PT > gpurun_out/g19_$v.log 2>&1; echo "exit $?"; grep -A10 "graph gradients differ" gpurun_out/g19_$v.log | head -12 | cut -c1-150
done
