#!/bin/bash
# What to run on a B200 box after a change (under gpurun): GPU parity tests, smoke, the default bench.
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 1200 python -m pytest tests -q -x -p no:cacheprovider --timeout=600 -m gpu > gpurun_out/pytest.log 2>&1; echo "exit $?"; tail -n 5 gpurun_out/pytest.log
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?"; tail -n 4 gpurun_out/smoke.log
echo "=== bench"; timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit $?"; tail -c 300 gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench.json'))
print('train ms', d['ms_per_step'], 'frames/s', d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'infer ms', d['infer']['ms_per_call'],
      'eager ms', d['eager']['ms_per_step'], 'roofline frac', d['roofline']['frac'])
PY
