#!/bin/bash
# 2-GPU check: graph capture with the NCCL bucket all-reduce inside, weak scaling
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench27_n2.json 2> gpurun_out/bench27_n2.err; echo "N=2 exit $?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench27_n2.json').read().strip().splitlines()[-1])
    print('train', d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], 'eager', d['eager']['ms_per_step'], 'graph', d['graph']['captured'], d['graph']['error'])
except Exception as e: print('parse failed', e)
PY
tail -n 12 gpurun_out/bench27_n2.err | cut -c1-300
