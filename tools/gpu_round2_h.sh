#!/bin/bash
# Round 2, GPU call H: MAS + attention CTC kernels, inv1x1 with 8 warps per CTA, scatter with batched 16-byte loads
mkdir -p gpurun_out
echo "=== pytest alignment"; timeout 900 python -m pytest tests/test_gpu_alignment.py -q -p no:cacheprovider --timeout=600 -m gpu > gpurun_out/r2h_pytest_align.log 2>&1; echo "exit $?"; tail -n 30 gpurun_out/r2h_pytest_align.log
echo "=== inv1x1 alone"; timeout 300 python tools/inv1x1_probe.py
echo "=== pytest gpu (rest)"; timeout 1500 python -m pytest tests -q -p no:cacheprovider --timeout=900 -m gpu --deselect tests/test_gpu_alignment.py > gpurun_out/r2h_pytest.log 2>&1; echo "exit $?"; tail -n 6 gpurun_out/r2h_pytest.log; grep -n "AssertionError\|^FAILED" gpurun_out/r2h_pytest.log | head -8
echo "=== bench"; timeout 1200 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "exit $?"; tail -c 600 gpurun_out/r2h_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2h_bench.json'))
print('train ms', d['ms_per_step'], 'frames/s', d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'infer ms', d['infer']['ms_per_call'],
      'eager ms', d['eager']['ms_per_step'], 'roofline frac', d['roofline']['frac'])
for r in d.get('roofline_hbm', []): print(r['kernel'][:60].ljust(60), r['us'], r['achieved'], r['frac'])
PY
echo "=== timeline graph"; timeout 300 python tools/timeline.py --graph > gpurun_out/r2h_timeline.txt 2>&1; echo "exit $?"; sed -n 40,60p gpurun_out/r2h_timeline.txt
