#!/bin/bash
mkdir -p gpurun_out
PT="python -m pytest -q -p no:cacheprovider --timeout=600 -m gpu"
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n 6 gpurun_out/$name.log; }
run gemm      $PT tests/test_gpu_ops.py -k "conv_rows or wgrad_rows"
run dec       $PT tests/test_gpu_decoder.py -k "fixture and (bf16x3 or bf16)"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench6_pair.json 2> gpurun_out/bench6_pair.err; echo "pair exit $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench6_pair.json'))
print('pair  ', d['ms_per_step'], d['value'], d['roofline']['per_launch_ms'], d['roofline']['executed_tflops'], d['contraction_ms_one_step'])
for k,v in d['contraction_kernels_one_step'].items(): print('   ',k, v)
PY
tail -3 gpurun_out/bench6_pair.err
RADMMM_B200_TC_PAIR=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench6_single.json 2>/dev/null; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench6_single.json'))
print('single', d['ms_per_step'], d['value'], d['roofline']['per_launch_ms'], d['roofline']['executed_tflops'], d['contraction_ms_one_step'])
PY
