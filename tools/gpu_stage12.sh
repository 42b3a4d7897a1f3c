#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest lstm+graph"; timeout 900 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_graph.py -q -x -p no:cacheprovider --timeout=600 -m gpu > gpurun_out/pytest12.log 2>&1; echo "exit $?"; tail -n 15 gpurun_out/pytest12.log
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench12.json 2> gpurun_out/bench12.err; echo "exit $?"; tail -c 600 gpurun_out/bench12.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench12.json'))
print('train', d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], 'infer', d['infer']['ms_per_call'], 'roof', d['roofline']['frac'], 'launches', d['gpu_launches'])
print('graph', d['graph']); print('eager', d['eager'])
for k,v in d['contraction_kernels_one_step'].items(): print(k, v)
PY
echo "=== ncu full k5"; timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'gemm_tc_kernel<\(int\)1, \(int\)1, \(int\)256' -s 40 -c 1 -o gpurun_out/prof_k5_s12 python bench.py --eager --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_k512.log 2>&1; echo "exit $?"; tail -3 gpurun_out/ncu_k512.log
echo "=== ncu full lstm"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_ -s 2 -c 2 -o gpurun_out/prof_lstm_s12 python bench.py --eager --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_lstm12.log 2>&1; echo "exit $?"
ls -la gpurun_out/*.ncu-rep
