#!/bin/bash
# Round 2, GPU call F: wgrad 256-wide tiles, narrow-output tiles, LSTM deferred stores / 2-step prefetch, PDL A/B, ncu inv1x1
mkdir -p gpurun_out
echo "=== lstm cluster probe"; timeout 300 python tools/lstm_cluster_probe.py > gpurun_out/r2f_lstm_probe.txt 2>&1; echo "exit $?"; head -14 gpurun_out/r2f_lstm_probe.txt
echo "=== pytest gpu (all)"; timeout 1500 python -m pytest tests -q -p no:cacheprovider --timeout=900 -m gpu > gpurun_out/r2f_pytest.log 2>&1; echo "exit $?"; tail -n 8 gpurun_out/r2f_pytest.log; grep -n "AssertionError" gpurun_out/r2f_pytest.log | head -5
echo "=== bench"; timeout 900 python bench.py --quick > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "exit $?"; tail -c 600 gpurun_out/r2f_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2f_bench.json'))
print('train ms', d['ms_per_step'], 'frames/s', d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'infer ms', d['infer']['ms_per_call'],
      'eager ms', d['eager']['ms_per_step'], 'roofline frac', d['roofline']['frac'])
print(json.dumps(d['contraction_kernels_one_step']))
PY
one() { python bench.py --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train ms', d['ms_per_step'], 'infer ms', d['infer']['ms_per_call'])"; }
echo "=== bench wgrad BN128 (old)"; RADMMM_B200_WGRAD_BN128=1 one
echo "=== bench PDL"; RADMMM_B200_PDL=1 one
echo "=== pytest graph+decoder with PDL"; RADMMM_B200_PDL=1 timeout 900 python -m pytest tests/test_gpu_graph.py tests/test_gpu_decoder.py -q -p no:cacheprovider --timeout=600 -m gpu -k "not lstm" 2>&1 | tail -3
echo "=== gemm timeline"; timeout 300 python tools/gemm_timeline.py > gpurun_out/r2f_gemm_timeline.txt 2>&1; echo "exit $?"; head -34 gpurun_out/r2f_gemm_timeline.txt
echo "=== ncu inv1x1"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:inv1x1_kernel -s 4 -c 1 -o gpurun_out/prof_inv1x1 python tools/gemm_timeline.py > gpurun_out/r2f_ncu_inv.log 2>&1; echo "exit $?"
echo "=== timeline graph"; timeout 300 python tools/timeline.py --graph > gpurun_out/r2f_timeline.txt 2>&1; echo "exit $?"; sed -n 40,75p gpurun_out/r2f_timeline.txt
