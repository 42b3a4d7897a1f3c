#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 1200 python -m pytest tests -q -x -p no:cacheprovider --timeout=600 -m gpu > gpurun_out/pytest17.log 2>&1; echo "exit $?"; tail -n 6 gpurun_out/pytest17.log
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench17.json 2> gpurun_out/bench17.err; echo "exit $?"; tail -c 300 gpurun_out/bench17.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench17.json'))
print('train', d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], 'infer', d['infer']['ms_per_call'], 'eager', d['eager']['ms_per_step'], 'graph', d['graph']['captured'], d['graph']['error'])
for k,v in d['contraction_kernels_one_step'].items(): print(k, v)
PY
timeout 600 python tools/timeline.py --graph --out gpurun_out/timeline17 > gpurun_out/timeline17.log 2>&1; echo "timeline exit $?"; grep -E "^span|any-stream" gpurun_out/timeline17.log; grep " us " gpurun_out/timeline17.log | head -36
