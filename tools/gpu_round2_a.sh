#!/bin/bash
# Round 2, GPU call A: new parity tests, bench (new sections), in-kernel GEMM timeline, cluster-exchange probe, LSTM A/B.
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -q -x -p no:cacheprovider --timeout=900 -m gpu > gpurun_out/r2a_pytest.log 2>&1; echo "exit $?"; tail -n 15 gpurun_out/r2a_pytest.log
echo "=== gemm timeline"; timeout 300 python tools/gemm_timeline.py > gpurun_out/r2a_gemm_timeline.txt 2>&1; echo "exit $?"; cat gpurun_out/r2a_gemm_timeline.txt | head -60
echo "=== cluster probe"; (cd tools/probes && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/cluster_xchg cluster_xchg.cu && timeout 120 /tmp/cluster_xchg) > gpurun_out/r2a_cluster_probe.txt 2>&1; echo "exit $?"; cat gpurun_out/r2a_cluster_probe.txt
echo "=== lstm A/B"; timeout 300 python tools/lstm_ab.py > gpurun_out/r2a_lstm_ab.txt 2>&1; echo "exit $?"; cat gpurun_out/r2a_lstm_ab.txt
echo "=== bench"; timeout 900 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "exit $?"; tail -c 600 gpurun_out/r2a_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2a_bench.json'))
print('train ms', d['ms_per_step'], 'frames/s', d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'infer ms', d['infer']['ms_per_call'],
      'eager ms', d['eager']['ms_per_step'], 'roofline frac', d['roofline']['frac'])
print('parity_mode', json.dumps(d.get('parity_mode'))[:600])
print('cpu', d.get('cpu_baseline'))
for r in d.get('roofline_hbm', []): print(r)
print('frontend', d.get('frontend'))
print(json.dumps(d['contraction_kernels_one_step']))
PY
