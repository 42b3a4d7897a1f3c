#!/bin/bash
# Round 2, GPU call D: exchange probe (bulk copies), LSTM v2 (tables, bulk DSMEM, one cluster per chunk), fused RAdam
mkdir -p gpurun_out
echo "=== cluster probe"; (cd tools/probes && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/cluster_xchg cluster_xchg.cu && timeout 120 /tmp/cluster_xchg) > gpurun_out/r2d_cluster_probe.txt 2>&1; echo "exit $?"; cat gpurun_out/r2d_cluster_probe.txt
echo "=== lstm tests"; timeout 600 python -m pytest tests/test_gpu_decoder.py -q -x -p no:cacheprovider --timeout=300 -m gpu -k "lstm" > gpurun_out/r2d_lstm.log 2>&1; echo "exit $?"; tail -n 6 gpurun_out/r2d_lstm.log
echo "=== lstm cluster probe (bulk)"; timeout 300 python tools/lstm_cluster_probe.py > gpurun_out/r2d_lstm_probe.txt 2>&1; echo "exit $?"; cat gpurun_out/r2d_lstm_probe.txt
echo "=== lstm cluster probe (st.async)"; RADMMM_B200_LSTM_BULK=0 timeout 300 python tools/lstm_cluster_probe.py > gpurun_out/r2d_lstm_probe_nobulk.txt 2>&1; echo "exit $?"; head -3 gpurun_out/r2d_lstm_probe_nobulk.txt
echo "=== lstm cluster probe B=32"; timeout 300 python tools/lstm_cluster_probe.py 32 > gpurun_out/r2d_lstm_probe32.txt 2>&1; echo "exit $?"; head -3 gpurun_out/r2d_lstm_probe32.txt
echo "=== lstm cluster probe B=64"; timeout 300 python tools/lstm_cluster_probe.py 64 > gpurun_out/r2d_lstm_probe64.txt 2>&1; echo "exit $?"; head -3 gpurun_out/r2d_lstm_probe64.txt
echo "=== pytest gpu (all)"; timeout 1500 python -m pytest tests -q -p no:cacheprovider --timeout=900 -m gpu > gpurun_out/r2d_pytest.log 2>&1; echo "exit $?"; tail -n 12 gpurun_out/r2d_pytest.log
echo "=== bench"; timeout 900 python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "exit $?"; tail -c 600 gpurun_out/r2d_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2d_bench.json'))
print('train ms', d['ms_per_step'], 'frames/s', d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'infer ms', d['infer']['ms_per_call'],
      'eager ms', d['eager']['ms_per_step'], 'roofline frac', d['roofline']['frac'])
print('optimizer', json.dumps(d.get('optimizer')))
print('parity', json.dumps(d.get('parity_mode'))[:300])
for r in d.get('roofline_hbm', []): print(r['kernel'][:40], r['us'], r['achieved'], r['frac'])
print(json.dumps(d['contraction_kernels_one_step']))
PY
echo "=== timeline graph"; timeout 300 python tools/timeline.py --graph > gpurun_out/r2d_timeline.txt 2>&1; echo "exit $?"; sed -n 50,95p gpurun_out/r2d_timeline.txt
