#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest graph"; timeout 600 python -m pytest tests/test_gpu_graph.py -q -x -p no:cacheprovider --timeout=300 -m gpu > gpurun_out/pytest26.log 2>&1; echo "exit $?"; tail -n 5 gpurun_out/pytest26.log | cut -c1-300
echo "=== default bench"; timeout 600 python bench.py > gpurun_out/bench26.json 2> gpurun_out/bench26.err; echo "exit $?"; tail -c 300 gpurun_out/bench26.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench26.json'))
print('train', d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], 'infer', d['infer'], 'eager', d['eager']['ms_per_step'])
PY
