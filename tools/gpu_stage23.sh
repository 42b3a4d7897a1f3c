#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 1200 python -m pytest tests -q -x -p no:cacheprovider --timeout=600 -m gpu > gpurun_out/pytest23.log 2>&1; echo "exit $?"; tail -n 6 gpurun_out/pytest23.log
timeout 300 python tools/lstm_probe.py > gpurun_out/lstm_probe23.log 2>&1; echo "exit $?"; cat gpurun_out/lstm_probe23.log
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench23.json 2> gpurun_out/bench23.err; echo "exit $?"; tail -c 300 gpurun_out/bench23.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench23.json'))
print('train', d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], 'infer', d['infer']['ms_per_call'], 'eager', d['eager']['ms_per_step'], 'graph', d['graph']['captured'], d['graph']['error'])
PY
