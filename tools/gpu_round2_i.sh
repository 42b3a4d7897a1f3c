#!/bin/bash
# Round 2, GPU call I: batched encoders, batched front end, inv1x1 32x64 tiles, whole-tile wgrad A/B, hard-alignment bench
mkdir -p gpurun_out
echo "=== pytest new"; timeout 900 python -m pytest tests/test_gpu_alignment.py tests/test_gpu_parity_full.py -q -p no:cacheprovider --timeout=600 -m gpu -k "alignment or ctc or mas or encoder or conv_lstm or front_end" > gpurun_out/r2i_pytest_new.log 2>&1; echo "exit $?"; tail -n 40 gpurun_out/r2i_pytest_new.log
echo "=== inv1x1 alone"; timeout 300 python tools/inv1x1_probe.py
echo "=== pytest ops"; timeout 900 python -m pytest tests/test_gpu_ops.py -q -p no:cacheprovider --timeout=600 -m gpu 2>&1 | tail -3
echo "=== bench"; timeout 1200 python bench.py > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "exit $?"; tail -c 600 gpurun_out/r2i_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2i_bench.json'))
print('train ms', d['ms_per_step'], 'frames/s', d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'infer ms', d['infer']['ms_per_call'],
      'eager ms', d['eager']['ms_per_step'], 'roofline frac', d['roofline']['frac'])
print(json.dumps(d.get('hard_alignment')))
PY
one() { python bench.py --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train ms', d['ms_per_step'], 'infer ms', d['infer']['ms_per_call'])"; }
echo "=== bench whole-tile wgrad, no split"; RADMMM_B200_WGRAD_WHOLE=1 one
echo "=== bench split-2 wgrad"; RADMMM_B200_WGRAD_BALANCED=0 one
echo "=== bench default again"; one
