#!/bin/bash
# final multi-GPU sanity exactly as the driver launches it (N=2): our arm, then the reference arm
mkdir -p gpurun_out
N=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2z_n2.json 2> gpurun_out/r2z_n2.err; echo "ours exit $?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2z_n2.json').read().strip().splitlines()[-1])
print('n_gpus', d['n_gpus'], 'train ms', d['ms_per_step'], 'frames/s', d['value'], 'e2e', d['e2e']['value'], 'graph', d['graph'])
c = d.get('allreduce_check') or {}
print('allreduce_check', c.get('max_abs_diff'), c.get('worst_parameter'), c.get('local_run_to_run_max_abs_diff'))
print('keys', sorted(d.keys()))
PY
tail -n 2 gpurun_out/r2z_n2.err | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-250
