#!/bin/bash
# Round 2, GPU call O: weight-norm backward without bank conflicts, front-end timing without host syncs, reducer regression test
mkdir -p gpurun_out
echo "=== pytest (flow step / decoder gradients, reducer regression)"; timeout 1200 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_graph.py tests/test_gpu_ops.py -q -p no:cacheprovider --timeout=900 -m gpu 2>&1 | tail -4
one() { python bench.py --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train ms', d['ms_per_step'], 'infer ms', d['infer']['ms_per_call'])"; }
echo "=== bench default"; one
echo "=== bench old wn_bwd (staged)"; RADMMM_B200_WNBWD_BULK=0 one
echo "=== bench default"; one
echo "=== frontend"; python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
import bench
pk = bench.peaks()
print(bench.frontend_bench(torch.device('cuda:0'), pk, 8, 800))
PY
echo "=== timeline graph"; timeout 300 python tools/timeline.py --graph > gpurun_out/r2o_timeline.txt 2>&1; echo "exit $?"; sed -n 38,62p gpurun_out/r2o_timeline.txt
