#!/bin/bash
# NCCL protocol / CTA-count variants for the in-graph bucketed all-reduce (gpurun --gpus N -- 'bash tools/gpu_nccl_variants.sh N')
N=${1:-2}
mkdir -p gpurun_out
run() {
  echo "=== $1"
  env $1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 --quick 2>gpurun_out/nccl_var.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('train ms', d['ms_per_step'], 'frames/s', round(d['value']), 'graph', d['graph']['captured'], 'check', d.get('allreduce_check', {}).get('max_abs_diff'))" || tail -3 gpurun_out/nccl_var.err | cut -c1-200
}
run "RADMMM_NOOP=1"
run "NCCL_PROTO=Simple NCCL_MAX_CTAS=16"
run "NCCL_PROTO=Simple"
run "NCCL_ALGO=NVLS"
