#!/bin/bash
# Round 2, GPU call N: attention backward rewrite (3 launches), full gpu suite, default bench
mkdir -p gpurun_out
echo "=== pytest attention"; timeout 900 python -m pytest tests/test_gpu_ops.py -q -p no:cacheprovider --timeout=600 -m gpu -k "attention" 2>&1 | tail -15
echo "=== pytest gpu (all)"; timeout 1500 python -m pytest tests -q -p no:cacheprovider --timeout=900 -m gpu > gpurun_out/r2n_pytest.log 2>&1; echo "exit $?"; tail -n 4 gpurun_out/r2n_pytest.log; grep -n "AssertionError\|^FAILED" gpurun_out/r2n_pytest.log | head -8
echo "=== bench"; timeout 1500 python bench.py > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "exit $?"; tail -c 300 gpurun_out/r2n_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2n_bench.json'))
print('train ms', d['ms_per_step'], 'frames/s', d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'infer ms', d['infer']['ms_per_call'])
print('joint', json.dumps(d.get('joint_training')))
for r in d.get('roofline_hbm', []): print(r['kernel'][:60].ljust(60), r['us'], r['achieved'], r['frac'])
print('frontend', d['frontend']['us'], d['frontend']['frames_per_s'])
PY
