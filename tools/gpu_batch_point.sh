#!/bin/bash
# One point of the synthetic sweep (BASELINE config 5): bash tools/gpu_batch_point.sh <batch> [frames]
B=${1:-32}; T=${2:-800}
mkdir -p gpurun_out
timeout 240 python bench.py --batch $B --frames $T --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b${B}_t${T}.json 2> gpurun_out/bench_b${B}_t${T}.err; echo "exit $?"; tail -c 300 gpurun_out/bench_b${B}_t${T}.err
python - "$B" "$T" <<'PY'
import json, sys
d = json.load(open(f'gpurun_out/bench_b{sys.argv[1]}_t{sys.argv[2]}.json'))
print('train ms', d['ms_per_step'], 'frames/s', d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'infer ms', d['infer']['ms_per_call'], 'eager ms', d['eager']['ms_per_step'],
      'k5 frac', d['roofline']['frac'], 'k5 executed TF', d['roofline']['executed_tflops'], 'step frac', d['step_tensor_roofline']['frac'], 'valid frames', d['config']['valid_frames_per_gpu'])
PY
