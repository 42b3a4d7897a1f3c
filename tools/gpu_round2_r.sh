#!/bin/bash
# Round 2, GPU call R: cluster-of-4 LSTM geometry for hidden <= 132 (predictors), joint training again
mkdir -p gpurun_out
echo "=== pytest lstm + encoders"; timeout 1200 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_parity_full.py -q -p no:cacheprovider --timeout=900 -m gpu -k "lstm or encoder or conv_lstm or predictor or joint" 2>&1 | tail -6
echo "=== bench quick + config3"; python bench.py --quick --config3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train ms', d['ms_per_step'], 'joint', json.dumps(d.get('joint_training')))"
echo "=== same with 16-CTA clusters for the predictors"; RADMMM_B200_LSTM_CL4=0 python bench.py --quick --config3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train ms', d['ms_per_step'], 'joint', d['joint_training']['ms_per_step'])"
