#!/bin/bash
# Round 2, GPU call C: LSTM phase timers + ncu, failing tests, grouped res-skip weight-grad, bench
mkdir -p gpurun_out
echo "=== lstm cluster probe"; timeout 300 python tools/lstm_cluster_probe.py > gpurun_out/r2c_lstm_probe.txt 2>&1; echo "exit $?"; cat gpurun_out/r2c_lstm_probe.txt
echo "=== lstm cluster probe B=32"; timeout 300 python tools/lstm_cluster_probe.py 32 > gpurun_out/r2c_lstm_probe32.txt 2>&1; echo "exit $?"; cat gpurun_out/r2c_lstm_probe32.txt
echo "=== pytest gpu (all)"; timeout 1500 python -m pytest tests -q -p no:cacheprovider --timeout=900 -m gpu > gpurun_out/r2c_pytest.log 2>&1; echo "exit $?"; tail -n 12 gpurun_out/r2c_pytest.log
echo "=== bench"; timeout 900 python bench.py --quick > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "exit $?"; tail -c 600 gpurun_out/r2c_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2c_bench.json'))
print('train ms', d['ms_per_step'], 'frames/s', d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'infer ms', d['infer']['ms_per_call'],
      'eager ms', d['eager']['ms_per_step'], 'roofline frac', d['roofline']['frac'])
print(json.dumps(d['contraction_kernels_one_step']))
PY
echo "=== ncu lstm fwd"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_cl_fwd -s 2 -c 1 -o gpurun_out/prof_lstm_cl_fwd python tools/lstm_cluster_probe.py > gpurun_out/r2c_ncu_lstm.log 2>&1; echo "exit $?"; tail -3 gpurun_out/r2c_ncu_lstm.log
