#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider --timeout=600 -m gpu"
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n 6 gpurun_out/$name.log; }
run dec       $PT tests/test_gpu_decoder.py -k "fixture or small"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench5_bf16.json 2> gpurun_out/bench5_bf16.err; echo "bf16 exit $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench5_bf16.json'))
print('fork  ', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'])
PY
RADMMM_B200_SIDE_STREAM=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench5_nofork.json 2>/dev/null; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench5_nofork.json'))
print('nofork', d['ms_per_step'], d['value'])
PY
tail -3 gpurun_out/bench5_bf16.err
