#!/bin/bash
# Round 2, GPU call X: LSTM backward tail on two streams
mkdir -p gpurun_out
echo "=== pytest lstm + graph + joint"; timeout 900 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_graph.py tests/test_gpu_parity_full.py -q -p no:cacheprovider --timeout=600 -m gpu -k "lstm or graph or joint or standin or radam or encoder or predictor or pool" 2>&1 | tail -3
one() { python bench.py --quick $1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'joint', (d.get('joint_training') or {}).get('ms_per_step'))"; }
echo "=== bench (tail on two streams)"; one --config3
echo "=== bench (one stream)"; RADMMM_B200_LSTM_TAIL_STREAM=0 one
echo "=== bench (tail on two streams)"; one
