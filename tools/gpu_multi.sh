#!/bin/bash
# multi-GPU validation: N ranks on one box (torchrun), weak scaling
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "N=$N exit $?"
tail -c 1200 gpurun_out/bench_n$N.json; tail -8 gpurun_out/bench_n$N.err
