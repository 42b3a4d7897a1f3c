#!/bin/bash
# Round-1 final measurement set: smoke, default bench (graph + eager + cpu baseline), reference arm, precision modes,
# ncu launch list of one eager step, ncu --set full of the dominant kernel and of the LSTM kernels.
mkdir -p gpurun_out
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke24.log 2>&1; echo "exit $?"; tail -n 4 gpurun_out/smoke24.log
echo "=== bench"; timeout 900 python bench.py > gpurun_out/bench24.json 2> gpurun_out/bench24.err; echo "exit $?"; tail -c 300 gpurun_out/bench24.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench24.json'))
print('train', d['ms_per_step'], d['value'], 'e2e', d['e2e'], 'infer', d['infer']['ms_per_call'], 'eager', d['eager']['ms_per_step'], 'roof', d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'], 'clocks', d['clocks'])
PY
echo "=== ref arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench24_ref.json 2> gpurun_out/bench24_ref.err; echo "exit $?"; cut -c1-200 gpurun_out/bench24_ref.json
echo "=== x3"; timeout 600 python bench.py --precision bf16x3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench24_x3.json 2> gpurun_out/bench24_x3.err; echo "exit $?"
echo "=== fp32"; timeout 600 python bench.py --precision fp32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench24_fp32.json 2> gpurun_out/bench24_fp32.err; echo "exit $?"
python - <<'PY'
import json
for n in ('x3','fp32'):
    d=json.load(open(f'gpurun_out/bench24_{n}.json')); print(n, d['ms_per_step'], d['value'], d['eager']['ms_per_step'], d['roofline']['frac'])
PY
echo "=== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches24.csv python bench.py --eager --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu24.log 2>&1; echo "exit $?"
echo "=== ncu full k5"; timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'gemm_tc_kernel<\(int\)1, \(int\)1, \(int\)256' -s 40 -c 1 -o gpurun_out/prof_k5_s24 python bench.py --eager --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_k524.log 2>&1; echo "exit $?"
echo "=== ncu full lstm"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_ -s 2 -c 2 -o gpurun_out/prof_lstm_s24 python bench.py --eager --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_lstm24.log 2>&1; echo "exit $?"
ls -la gpurun_out/*s24.ncu-rep
