#!/bin/bash
# leaf runs inside windowed passes (run assigned to the pass of its union box's entry time): tests, then config 3 A/B
mkdir -p gpurun_out
echo "== tests" | tee gpurun_out/r2q.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py -q -m gpu -x --timeout 600 -s -k "leaf_runs or pruning or detect_exact or quirks or stress or config1" 2>&1 | grep -v "^$" | tail -8 | tee -a gpurun_out/r2q.log
for runs in 0 1 0 1; do
  XP_LEAF_RUNS=$runs timeout 400 python bench.py --config 3 --no-cpu 2>gpurun_out/r2q.err | tail -1 > gpurun_out/r2q_c3_$runs.json
  python -c "
import json
d=json.loads(open('gpurun_out/r2q_c3_$runs.json').read())['detail']
print('config3 runs=$runs', 'frame %.2f' % d['ms_per_frame'], {k: round(v,2) for k,v in d['sweep_stage_ms'].items()}, 'cand %.2f' % d['candidates_per_sphere'], 'hit', round(d['hit_fraction'],6))" | tee -a gpurun_out/r2q.log
done
