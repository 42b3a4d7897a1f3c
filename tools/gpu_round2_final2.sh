#!/bin/bash
# Round 2, last GPU call: full gpu suite, smoke, default bench on the final build
mkdir -p gpurun_out
echo "=== pytest gpu (all)"; timeout 1500 python -m pytest tests -q -p no:cacheprovider --timeout=900 -m gpu > gpurun_out/r2zz_pytest.log 2>&1; echo "exit $?"; tail -n 3 gpurun_out/r2zz_pytest.log; grep -n "AssertionError\|^FAILED\|Error" gpurun_out/r2zz_pytest.log | head -8
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench"; SECONDS=0; timeout 1500 python bench.py > gpurun_out/r2zz_bench.json 2> gpurun_out/r2zz_bench.err; echo "exit $? wall ${SECONDS}s"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2zz_bench.json'))
print('train ms', d['ms_per_step'], 'frames/s', d['value'], 'e2e ms', d['e2e']['ms_per_step'], d['e2e']['value'], 'infer ms', d['infer']['ms_per_call'], 'eager', d['eager']['ms_per_step'], 'frac', d['roofline']['frac'], d['step_tensor_roofline']['frac'])
print('joint', d['joint_training']['ms_per_step'], 'opt', d['optimizer']['optimizer_ms'], d['optimizer']['full_step']['ms_per_step'], 'x3', d['parity_mode']['ms_per_step'], 'cpu', d['cpu_baseline']['value'])
PY
