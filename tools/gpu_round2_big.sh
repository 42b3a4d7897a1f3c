#!/bin/bash
# final build at the large end of the config-5 sweep
for bt in "64 2048" "64 800"; do
  set -- $bt
  python bench.py --quick --batch $1 --frames $2 --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B=$1 T=$2', 'train ms', round(d['ms_per_step'],3), 'frames/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'step frac', round(d['step_tensor_roofline']['frac'],3), 'k5 frac', round(d['roofline']['frac'],3), 'infer ms', round(d['infer']['ms_per_call'],3), 'infer frames/s', round(d['infer']['value']))"
done
