#!/bin/bash
# Round 2, final GPU call: full gpu suite, smoke, default bench, ncu --set full of the dominant kernel + launch list (final kernels)
mkdir -p gpurun_out
echo "=== pytest gpu (all)"; timeout 1500 python -m pytest tests -q -p no:cacheprovider --timeout=900 -m gpu > gpurun_out/r2z_pytest.log 2>&1; echo "exit $?"; tail -n 3 gpurun_out/r2z_pytest.log; grep -n "AssertionError\|^FAILED\|Error" gpurun_out/r2z_pytest.log | head -8
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench"; SECONDS=0; timeout 1500 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "exit $? wall ${SECONDS}s"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2z_bench.json'))
print('train ms', d['ms_per_step'], 'frames/s', d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'infer ms', d['infer']['ms_per_call'], 'eager', d['eager']['ms_per_step'], 'frac', d['roofline']['frac'], d['step_tensor_roofline']['frac'])
print('joint', d['joint_training']['ms_per_step'], 'opt', d['optimizer']['optimizer_ms'], d['optimizer']['full_step']['ms_per_step'], 'x3', d['parity_mode']['ms_per_step'], 'cpu', d['cpu_baseline']['value'])
PY
echo "=== ncu full k5"; timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'gemm_tc_kernel<\(int\)1, \(int\)1, \(int\)256' -s 40 -c 1 -o gpurun_out/prof_k5_r2z python bench.py --eager --steps 1 --warmup 1 --no-cpu-baseline --quick > gpurun_out/r2z_ncu_k5.log 2>&1; echo "exit $?"
echo "=== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --eager --steps 1 --warmup 1 --no-cpu-baseline --quick > gpurun_out/r2z_ncu_launches.log 2>&1; echo "exit $?"
ls -la gpurun_out/prof_k5_r2z.ncu-rep
