#!/bin/bash
# Round 2, GPU call T: default bench (final line for profiles/), wall time of the default run
mkdir -p gpurun_out
echo "=== bench"; SECONDS=0; timeout 1500 python bench.py > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "exit $? wall ${SECONDS}s"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2t_bench.json'))
print('train ms', d['ms_per_step'], 'frames/s', d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'infer ms', d['infer']['ms_per_call'], 'eager', d['eager']['ms_per_step'], 'frac', d['roofline']['frac'], d['step_tensor_roofline']['frac'])
print('joint', d['joint_training']['ms_per_step'], 'opt', d['optimizer']['optimizer_ms'], d['optimizer']['full_step']['ms_per_step'], 'x3', d['parity_mode']['ms_per_step'], 'cpu', d['cpu_baseline']['value'])
PY
echo "=== N=1 sanity of --config3 under torchrun-less multi flag"; python bench.py --quick --config3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('joint', d['joint_training']['ms_per_step'])"
