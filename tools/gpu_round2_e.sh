#!/bin/bash
# Round 2, GPU call E: LSTM pointwise probes, fused bias column sums, bulk weight-norm backward, RAdam-in-graph test, bench
mkdir -p gpurun_out
echo "=== lstm cluster probe"; timeout 300 python tools/lstm_cluster_probe.py > gpurun_out/r2e_lstm_probe.txt 2>&1; echo "exit $?"; cat gpurun_out/r2e_lstm_probe.txt
echo "=== pytest gpu (all)"; timeout 1500 python -m pytest tests -q -p no:cacheprovider --timeout=900 -m gpu > gpurun_out/r2e_pytest.log 2>&1; echo "exit $?"; tail -n 8 gpurun_out/r2e_pytest.log
echo "=== bench"; timeout 900 python bench.py --quick > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "exit $?"; tail -c 600 gpurun_out/r2e_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2e_bench.json'))
print('train ms', d['ms_per_step'], 'frames/s', d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'infer ms', d['infer']['ms_per_call'],
      'eager ms', d['eager']['ms_per_step'], 'roofline frac', d['roofline']['frac'])
print(json.dumps(d['contraction_kernels_one_step']))
PY
echo "=== bench (old wn_bwd)"; RADMMM_B200_WNBWD_BULK=0 timeout 900 python bench.py --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train ms', d['ms_per_step'])"
echo "=== timeline graph"; timeout 300 python tools/timeline.py --graph > gpurun_out/r2e_timeline.txt 2>&1; echo "exit $?"; sed -n 50,90p gpurun_out/r2e_timeline.txt
