#!/bin/bash
# Round 2, GPU call G: balanced (equal K-block runs) wgrad walk A/B, linear splines, RAdam-in-graph test, inv1x1 alone + ncu,
# r2 evidence: ncu launch list + ncu --set full of the dominant kernel
mkdir -p gpurun_out
echo "=== pytest gpu (all)"; timeout 1500 python -m pytest tests -q -p no:cacheprovider --timeout=900 -m gpu > gpurun_out/r2g_pytest.log 2>&1; echo "exit $?"; tail -n 8 gpurun_out/r2g_pytest.log; grep -n "AssertionError" gpurun_out/r2g_pytest.log | head -5
echo "=== bench"; timeout 1200 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "exit $?"; tail -c 600 gpurun_out/r2g_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2g_bench.json'))
print('train ms', d['ms_per_step'], 'frames/s', d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'infer ms', d['infer']['ms_per_call'],
      'eager ms', d['eager']['ms_per_step'], 'roofline frac', d['roofline']['frac'])
print(json.dumps(d['contraction_kernels_one_step']))
for r in d.get('hbm_rooflines', []): print(r['kernel'][:40], r['us'], r['frac'])
PY
one() { python bench.py --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train ms', d['ms_per_step'], 'infer ms', d['infer']['ms_per_call'])"; }
echo "=== bench whole-tile wgrad (old)"; RADMMM_B200_WGRAD_BALANCED=0 one
echo "=== gemm timeline"; timeout 300 python tools/gemm_timeline.py > gpurun_out/r2g_gemm_timeline.txt 2>&1; echo "exit $?"; head -34 gpurun_out/r2g_gemm_timeline.txt
echo "=== inv1x1 alone"; timeout 300 python tools/inv1x1_probe.py
echo "=== ncu inv1x1"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:inv1x1_kernel -s 4 -c 1 -o gpurun_out/prof_inv1x1 python tools/inv1x1_probe.py > gpurun_out/r2g_ncu_inv.log 2>&1; echo "exit $?"
echo "=== timeline graph"; timeout 300 python tools/timeline.py --graph > gpurun_out/r2g_timeline.txt 2>&1; echo "exit $?"; sed -n 40,75p gpurun_out/r2g_timeline.txt
echo "=== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2g_launches.csv python bench.py --eager --steps 1 --warmup 1 --no-cpu-baseline --quick > gpurun_out/r2g_ncu_launches.log 2>&1; echo "exit $?"
echo "=== ncu full k5"; timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'gemm_tc_kernel<\(int\)1, \(int\)1, \(int\)256' -s 40 -c 1 -o gpurun_out/prof_k5_r2g python bench.py --eager --steps 1 --warmup 1 --no-cpu-baseline --quick > gpurun_out/r2g_ncu_k5.log 2>&1; echo "exit $?"
ls -la gpurun_out/*.ncu-rep
