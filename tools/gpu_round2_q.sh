#!/bin/bash
# Round 2, GPU call Q: 256-wide tiles for odd-width weight-grads
mkdir -p gpurun_out
echo "=== pytest ops + decoder"; timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_decoder.py tests/test_gpu_graph.py -q -p no:cacheprovider --timeout=900 -m gpu 2>&1 | tail -3
one() { python bench.py --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train ms', d['ms_per_step'], 'infer ms', d['infer']['ms_per_call'])"; }
echo "=== bench"; one; one
