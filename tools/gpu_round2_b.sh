#!/bin/bash
# Round 2, GPU call B: cluster LSTM + branch-free epilogues + new helper kernels: targeted tests first, then everything.
mkdir -p gpurun_out
echo "=== lstm tests"; timeout 600 python -m pytest tests/test_gpu_decoder.py -q -x -p no:cacheprovider --timeout=300 -m gpu -k "lstm" > gpurun_out/r2b_lstm.log 2>&1; echo "exit $?"; tail -n 12 gpurun_out/r2b_lstm.log
echo "=== lstm A/B"; timeout 300 python tools/lstm_ab.py > gpurun_out/r2b_lstm_ab.txt 2>&1; echo "exit $?"; cat gpurun_out/r2b_lstm_ab.txt
echo "=== pytest gpu (all)"; timeout 1500 python -m pytest tests -q -p no:cacheprovider --timeout=900 -m gpu > gpurun_out/r2b_pytest.log 2>&1; echo "exit $?"; tail -n 25 gpurun_out/r2b_pytest.log
echo "=== gemm timeline"; timeout 300 python tools/gemm_timeline.py > gpurun_out/r2b_gemm_timeline.txt 2>&1; echo "exit $?"; head -40 gpurun_out/r2b_gemm_timeline.txt
echo "=== bench"; timeout 900 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "exit $?"; tail -c 600 gpurun_out/r2b_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2b_bench.json'))
print('train ms', d['ms_per_step'], 'frames/s', d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'infer ms', d['infer']['ms_per_call'],
      'eager ms', d['eager']['ms_per_step'], 'roofline frac', d['roofline']['frac'])
print('parity_mode', json.dumps(d.get('parity_mode'))[:700])
print('cpu', d.get('cpu_baseline'))
for r in d.get('roofline_hbm', []): print(r['kernel'][:40], r['us'], r['achieved'], r['frac'])
print(json.dumps(d['contraction_kernels_one_step']))
PY
echo "=== timeline graph"; timeout 300 python tools/timeline.py --graph > gpurun_out/r2b_timeline.txt 2>&1; echo "exit $?"; head -45 gpurun_out/r2b_timeline.txt
