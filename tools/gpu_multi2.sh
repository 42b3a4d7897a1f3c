#!/bin/bash
# Round 2 multi-GPU run (gpurun --gpus N -- 'bash tools/gpu_multi2.sh N'): bench at N (graph with the bucketed NCCL all-reduce
# inside, allreduce_check), then a torch.profiler timeline of one replay on rank 0.  Hard timeouts throughout.
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --config3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "N=$N exit $?"
python - "$N" <<'PY'
import json, sys
d = json.loads(open(f'gpurun_out/r2_bench_n{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('train ms', d['ms_per_step'], 'frames/s', d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'eager ms', d['eager']['ms_per_step'], 'graph', d['graph']['captured'], d['graph']['error'])
c = d.get('allreduce_check') or {}; print('allreduce_check', c.get('max_abs_diff'), c.get('worst_parameter'), c.get('local_run_to_run_max_abs_diff')); print('joint', json.dumps(d.get('joint_training')))
PY
tail -n 5 gpurun_out/r2_bench_n$N.err | cut -c1-300
echo "=== timeline N=$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/timeline.py --graph --out gpurun_out/timeline_n$N > gpurun_out/r2_timeline_n$N.txt 2>&1; echo "exit $?"
grep -A 14 "NCCL kernels" gpurun_out/r2_timeline_n$N.txt | head -20; grep "^span\|any-stream" gpurun_out/r2_timeline_n$N.txt
if [ "$N" = "2" ]; then
  echo "=== reference arm under torchrun (rank 0 only)"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
fi
