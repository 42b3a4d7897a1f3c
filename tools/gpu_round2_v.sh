#!/bin/bash
# Round 2, GPU call V: TMA store also in the dgrad (EPI_DH) epilogue
mkdir -p gpurun_out
echo "=== pytest ops + decoder + graph"; timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_decoder.py tests/test_gpu_graph.py tests/test_gpu_parity_full.py -q -p no:cacheprovider --timeout=900 -m gpu -x 2>&1 | tail -4
one() { python bench.py --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train ms', d['ms_per_step'], 'infer ms', d['infer']['ms_per_call'], 'k5 frac', d['roofline']['frac'])"; }
echo "=== bench"; one; one
echo "=== bench per-lane stores"; RADMMM_B200_TMA_STORE=0 one
