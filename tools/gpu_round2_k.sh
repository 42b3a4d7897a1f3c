#!/bin/bash
# Round 2, GPU call K: config 4 (inference sigma sweep + reference infer on CPU and on the same GPU), full gpu test suite
mkdir -p gpurun_out
echo "=== config 4"; timeout 1500 python tools/sweep_configs.py --config4 > gpurun_out/r2k_config4.md 2> gpurun_out/r2k_config4.err; echo "exit $?"; cat gpurun_out/r2k_config4.md; tail -n 5 gpurun_out/r2k_config4.err
echo "=== pytest gpu (all)"; timeout 1500 python -m pytest tests -q -p no:cacheprovider --timeout=900 -m gpu > gpurun_out/r2k_pytest.log 2>&1; echo "exit $?"; tail -n 4 gpurun_out/r2k_pytest.log; grep -n "AssertionError\|^FAILED" gpurun_out/r2k_pytest.log | head -8
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== bench quick"; python bench.py --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train ms', d['ms_per_step'], 'infer ms', d['infer']['ms_per_call'], 'frac', d['roofline']['frac'])"
