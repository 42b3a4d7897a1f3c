#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 900 python -m pytest tests -q -x -p no:cacheprovider --timeout=600 -m gpu > gpurun_out/pytest11.log 2>&1; echo "exit $?"; tail -n 4 gpurun_out/pytest11.log
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench11.json 2> gpurun_out/bench11.err; echo "exit $?"; tail -c 400 gpurun_out/bench11.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench11.json'))
print('train', d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], 'infer', d['infer']['ms_per_call'], 'roof', d['roofline']['frac'], 'host', d.get('host_enqueue_ms_per_step'), 'launches', d['gpu_launches'])
for k,v in d['contraction_kernels_one_step'].items(): print(k, v)
PY
timeout 600 python tools/timeline.py --out gpurun_out/timeline11 > gpurun_out/timeline11.log 2>&1; echo "timeline exit $?"; head -n 40 gpurun_out/timeline11.log
echo "=== ncu full lstm"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_ -s 2 -c 2 -o gpurun_out/prof_lstm_s11 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_lstm11.log 2>&1; echo "exit $?"
echo "=== ncu full k5"; timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"gemm_tc_kernel<1, 1, 256, 0, 2>" -s 40 -c 1 -o gpurun_out/prof_k5_s11 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_k511.log 2>&1; echo "exit $?"; tail -3 gpurun_out/ncu_k511.log
ls -la gpurun_out/*.ncu-rep
