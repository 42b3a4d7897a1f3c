#!/bin/bash
mkdir -p gpurun_out
echo "=== default bench (as the driver runs it)"; timeout 600 python bench.py > gpurun_out/bench25.json 2> gpurun_out/bench25.err; echo "exit $?"; tail -c 300 gpurun_out/bench25.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench25.json'))
print('train', d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], 'infer', d['infer']['ms_per_call'], 'eager', d['eager']['ms_per_step'], 'roof', d['roofline']['frac'], 'cpu', d['cpu_baseline'], 'clocks', d['clocks'])
PY
echo "=== coop build"; RADMMM_B200_LIB=$PWD/rad-mmm_b200/libradmmm_b200_coop.so timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench25_coop.json 2> gpurun_out/bench25_coop.err; echo "exit $?"; tail -c 300 gpurun_out/bench25_coop.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench25_coop.json'))
print('coop train', d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], 'infer', d['infer']['ms_per_call'], 'eager', d['eager']['ms_per_step'], 'roof', d['roofline']['frac'])
PY
