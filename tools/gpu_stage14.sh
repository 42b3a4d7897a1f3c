#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest lstm"; timeout 900 python -m pytest tests/test_gpu_decoder.py -q -x -p no:cacheprovider --timeout=600 -m gpu -k "lstm or fixture" > gpurun_out/pytest14.log 2>&1; echo "exit $?"; tail -n 3 gpurun_out/pytest14.log
echo "=== probe"; timeout 600 python tools/gemm_probe.py > gpurun_out/probe14.log 2>&1; echo "exit $?"; cat gpurun_out/probe14.log
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench14.json 2> gpurun_out/bench14.err; echo "exit $?"; tail -c 300 gpurun_out/bench14.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench14.json'))
print('train', d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], 'infer', d['infer']['ms_per_call'], 'eager', d['eager']['ms_per_step'])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:lstm_ -c 4 python bench.py --eager --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | grep -A3 "lstm_.*kernel" | grep -E "lstm_|duration" | head
