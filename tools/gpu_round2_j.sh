#!/bin/bash
# Round 2, GPU call J: config 4 / 5 sweeps, never-split A/B, default bench with the reference on the same GPU
mkdir -p gpurun_out
one() { python bench.py --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train ms', d['ms_per_step'], 'infer ms', d['infer']['ms_per_call'])"; }
echo "=== bench default"; one
echo "=== bench never split"; RADMMM_B200_WGRAD_WHOLE=2 one
echo "=== bench default"; one
echo "=== bench never split"; RADMMM_B200_WGRAD_WHOLE=2 one
echo "=== full bench"; timeout 1500 python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "exit $?"; tail -c 400 gpurun_out/r2j_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2j_bench.json'))
print('train ms', d['ms_per_step'], 'frames/s', d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'infer ms', d['infer']['ms_per_call'])
print(json.dumps(d.get('reference_same_gpu')))
PY
echo "=== config 4"; timeout 1500 python tools/sweep_configs.py --config4 > gpurun_out/r2j_config4.md 2> gpurun_out/r2j_config4.err; echo "exit $?"; tail -n 30 gpurun_out/r2j_config4.md; tail -n 5 gpurun_out/r2j_config4.err
echo "=== config 5"; timeout 2400 python tools/sweep_configs.py --config5 --points full > gpurun_out/r2j_config5.md 2> gpurun_out/r2j_config5.err; echo "exit $?"; cat gpurun_out/r2j_config5.md; tail -n 5 gpurun_out/r2j_config5.err
