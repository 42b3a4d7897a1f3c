#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest graph"; timeout 900 python -m pytest tests/test_gpu_graph.py -q -x -p no:cacheprovider --timeout=600 -m gpu > gpurun_out/pytest13.log 2>&1; echo "exit $?"; tail -n 15 gpurun_out/pytest13.log
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench13.json 2> gpurun_out/bench13.err; echo "exit $?"; tail -c 600 gpurun_out/bench13.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench13.json'))
print('train', d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], 'infer', d['infer']['ms_per_call'], 'roof', d['roofline']['frac'], 'launches', d['gpu_launches'])
print('graph', d['graph']); print('eager', d['eager'])
for k,v in d['contraction_kernels_one_step'].items(): print(k, v)
PY
