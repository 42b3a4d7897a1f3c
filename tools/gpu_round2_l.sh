#!/bin/bash
# Round 2, GPU call L (2 GPUs): which gradient differs in allreduce_check; config 3 at N=2; new attention / STFT kernels; DAP test
mkdir -p gpurun_out
echo "=== pytest (attention, stft, predictor, front end)"; timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity_full.py -q -p no:cacheprovider --timeout=600 -m gpu -k "attention or stft or predictor or front_end or encoder or conv_lstm" 2>&1 | tail -8
echo "=== N=2 bench with config 3"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 --quick --config3 > gpurun_out/r2l_n2.json 2> gpurun_out/r2l_n2.err; echo "exit $?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2l_n2.json').read().strip().splitlines()[-1])
print('train ms', d['ms_per_step'], 'frames/s', d['value'])
print('allreduce_check', json.dumps(d.get('allreduce_check')))
print('joint', json.dumps(d.get('joint_training')))
PY
tail -n 3 gpurun_out/r2l_n2.err | cut -c1-300
echo "=== N=1 full bench"; CUDA_VISIBLE_DEVICES=0 timeout 1500 python bench.py > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "exit $?"; tail -c 300 gpurun_out/r2l_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2l_bench.json'))
print('train ms', d['ms_per_step'], 'frames/s', d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'infer ms', d['infer']['ms_per_call'])
print('joint', json.dumps(d.get('joint_training')))
for r in d.get('roofline_hbm', []): print(r['kernel'][:60].ljust(60), r['us'], r['achieved'], r['frac'])
print('frontend', d['frontend']['us'], d['frontend']['frames_per_s'])
PY
