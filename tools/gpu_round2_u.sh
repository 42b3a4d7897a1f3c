#!/bin/bash
# Round 2, GPU call U: TMA-store epilogue (UTMASTG) -- parity tests, A/B
mkdir -p gpurun_out
echo "=== pytest ops + decoder + graph"; timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_decoder.py tests/test_gpu_graph.py -q -p no:cacheprovider --timeout=900 -m gpu -x 2>&1 | tail -6
one() { python bench.py --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train ms', d['ms_per_step'], 'infer ms', d['infer']['ms_per_call'], 'k5 frac', d['roofline']['frac'])"; }
echo "=== bench TMA store"; one
echo "=== bench per-lane stores"; RADMMM_B200_TMA_STORE=0 one
echo "=== bench TMA store"; one
echo "=== gemm timeline"; timeout 300 python tools/gemm_timeline.py > gpurun_out/r2u_gemm_timeline.txt 2>&1; echo "exit $?"; sed -n 1,13p gpurun_out/r2u_gemm_timeline.txt
