#!/bin/bash
# Round 2, GPU call W: prefetch pipeline for the e2e path
mkdir -p gpurun_out
echo "=== pytest graph"; timeout 900 python -m pytest tests/test_gpu_graph.py -q -p no:cacheprovider --timeout=600 -m gpu 2>&1 | tail -3
echo "=== bench"; python bench.py --quick 2>gpurun_out/r2w.err | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'e2e value', d['e2e']['value'], 'h2d', d['e2e']['h2d_bytes_per_step'])"; tail -2 gpurun_out/r2w.err | cut -c1-200
