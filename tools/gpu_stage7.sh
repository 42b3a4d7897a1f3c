#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider --timeout=600 -m gpu"
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n 6 gpurun_out/$name.log; }
run dec       $PT tests/test_gpu_decoder.py -k "fixture or infer or roundtrip"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench7.json 2> gpurun_out/bench7.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench7.json'))
print('train', d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'])
print('infer', d['infer'])
PY
tail -3 gpurun_out/bench7.err
