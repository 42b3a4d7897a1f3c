#!/bin/bash
# Multi-GPU weak-scaling check (gpurun --gpus N -- 'bash tools/gpu_multi.sh N'): graph capture with the bucketed NCCL
# all-reduce inside.  Hard timeout: a rendezvous problem must not burn the GPU budget.
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "N=$N exit $?"
python - "$N" <<'PY'
import json, sys
d = json.loads(open(f'gpurun_out/bench_n{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('train ms', d['ms_per_step'], 'frames/s', d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'eager ms', d['eager']['ms_per_step'], 'graph', d['graph']['captured'], d['graph']['error'])
PY
tail -n 5 gpurun_out/bench_n$N.err | cut -c1-300
