#!/bin/bash
# Round 2, GPU call M (2 GPUs): reducer stream-join fix (allreduce_check), overlapped joint training at N=2 and N=1
mkdir -p gpurun_out
echo "=== N=2 bench with config 3"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 --quick --config3 > gpurun_out/r2m_n2.json 2> gpurun_out/r2m_n2.err; echo "exit $?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2m_n2.json').read().strip().splitlines()[-1])
print('train ms', d['ms_per_step'], 'frames/s', d['value'])
c = d.get('allreduce_check') or {}
print('allreduce_check', c.get('max_abs_diff'), c.get('worst_parameter'), c.get('local_run_to_run_max_abs_diff'), c.get('per_bucket_max_abs_diff'))
print('joint', json.dumps(d.get('joint_training')))
PY
tail -n 3 gpurun_out/r2m_n2.err | cut -c1-300
echo "=== N=1 joint"; CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py --quick --config3 2> gpurun_out/r2m_n1.err | tail -1 > gpurun_out/r2m_n1.json; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2m_n1.json').read())
print('train ms', d['ms_per_step'], 'joint', json.dumps(d.get('joint_training')))
PY
tail -n 3 gpurun_out/r2m_n1.err | cut -c1-300
echo "=== ddp / graph tests on GPU"; CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests/test_gpu_graph.py tests/test_gpu_parity_full.py -q -p no:cacheprovider --timeout=600 -m gpu -k "graph or radam or standin or pool" 2>&1 | tail -3
