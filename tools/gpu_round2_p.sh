#!/bin/bash
# Round 2, GPU call P: full gpu suite, smoke, default bench, timeline after the tail fixes (no dx, even wgrad tiles), STFT 2-in-1 FFT
mkdir -p gpurun_out
echo "=== pytest gpu (all)"; timeout 1500 python -m pytest tests -q -p no:cacheprovider --timeout=900 -m gpu > gpurun_out/r2p_pytest.log 2>&1; echo "exit $?"; tail -n 4 gpurun_out/r2p_pytest.log; grep -n "AssertionError\|^FAILED\|Error" gpurun_out/r2p_pytest.log | head -8
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== bench"; timeout 1500 python bench.py > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "exit $?"; tail -c 300 gpurun_out/r2p_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2p_bench.json'))
print('train ms', d['ms_per_step'], 'frames/s', d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'infer ms', d['infer']['ms_per_call'], 'eager', d['eager']['ms_per_step'], 'frac', d['roofline']['frac'], d['step_tensor_roofline']['frac'])
print('joint', d['joint_training']['ms_per_step'], 'opt', d['optimizer']['optimizer_ms'], d['optimizer']['full_step']['ms_per_step'], 'x3', d['parity_mode']['ms_per_step'], 'cpu', d['cpu_baseline']['value'])
print('frontend', d['frontend']['us'], d['frontend']['frames_per_s'])
PY
echo "=== timeline graph"; timeout 300 python tools/timeline.py --graph > gpurun_out/r2p_timeline.txt 2>&1; echo "exit $?"; head -3 gpurun_out/r2p_timeline.txt | tail -1
python - <<'PY'
import csv
rows=[(float(r['start_us']),float(r['dur_us']),r['stream'],r['name'][:60]) for r in csv.DictReader(open('gpurun_out/timeline.csv'))]
end=max(s+d for s,d,_,_ in rows)
l=[r for r in rows if 'lstm_cl_bwd' in r[3]][0]
print('span', end, 'lstm bwd end', l[0]+l[1], 'tail', end-(l[0]+l[1]))
for r in rows:
    if r[0] >= l[0]+l[1]-5: print(f"{r[0]:8.1f} {r[1]:7.1f} s{r[2]} {r[3]}")
PY
