#!/bin/bash
# Re-entry checkpoint: whole GPU suite, smoke, default bench, reference arm, ncu launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi9.txt
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -q -x -p no:cacheprovider --timeout=900 -m gpu > gpurun_out/pytest9.log 2>&1; echo "exit $?"; tail -n 8 gpurun_out/pytest9.log
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke9.log 2>&1; echo "exit $?"; tail -n 5 gpurun_out/smoke9.log
echo "=== bench"; timeout 900 python bench.py > gpurun_out/bench9.json 2> gpurun_out/bench9.err; echo "exit $?"; tail -c 600 gpurun_out/bench9.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench9.json'))
print('train', d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], 'infer', d['infer']['ms_per_call'], 'roof', d['roofline']['frac'], 'cpu', d.get('cpu_baseline',{}).get('value'))
for k,v in d['contraction_kernels_one_step'].items(): print(k, v)
PY
echo "=== bench x3"; timeout 600 python bench.py --precision bf16x3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench9_x3.json 2> gpurun_out/bench9_x3.err; echo "exit $?"
echo "=== ref arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench9_ref.json 2> gpurun_out/bench9_ref.err; echo "exit $?"; cat gpurun_out/bench9_ref.json | cut -c1-400
echo "=== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches9.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu9.log 2>&1; echo "exit $?"
wc -l gpurun_out/launches9.csv
