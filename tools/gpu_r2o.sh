#!/bin/bash
# leaf runs (XP_OPT_LEAF_RUNS): tree + parity tests, then A/B on config 2 handed over as a soup, config 3 and config 4
mkdir -p gpurun_out
echo "== tests" | tee gpurun_out/r2o.log
timeout 900 python -m pytest tests/test_gpu_lbvh.py tests/test_gpu_parity.py -q -m gpu -x --timeout 600 \
  -k "tree or leaf_runs or broadphase_pairs_bit_exact or pair_buffer_overflow or pipeline_chunks or pruning or detect_exact" 2>&1 | tail -6 | tee -a gpurun_out/r2o.log
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
s=d['stage_ms']
print('$1', 'step %.3f ms' % d['ms_per_step'], 'walk %.3f narrow %.3f' % (s['traverse'], s['narrowphase']), 'tests/step', d.get('tests_per_step'), 'visits', d.get('walk_visits_per_step'))"; }
for so in "" build_ab/libxp_span64.so; do
for runs in 0 1; do
  [ -n "$so" ] && [ $runs = 0 ] && continue
  echo "== soup runs=$runs so=$so" | tee -a gpurun_out/r2o.log
  XP_PHYSICS_CUDA_SO=${so:+$PWD/$so} XP_LEAF_RUNS=$runs timeout 300 python bench.py --scene soup --steps 20 --warmup 3 --no-cpu --no-extras 2>gpurun_out/r2o.err > gpurun_out/r2o_soup_${runs}${so:+_span64}.json \
    && show "soup runs=$runs" < gpurun_out/r2o_soup_${runs}${so:+_span64}.json | tee -a gpurun_out/r2o.log || tail -3 gpurun_out/r2o.err
done
done
for runs in 0 1; do
  echo "== config 3 runs=$runs" | tee -a gpurun_out/r2o.log
  XP_LEAF_RUNS=$runs timeout 400 python bench.py --config 3 --no-cpu 2>gpurun_out/r2o.err | tail -1 > gpurun_out/r2o_c3_$runs.json
  python -c "
import json
d=json.loads(open('gpurun_out/r2o_c3_$runs.json').read())['detail']
print('config3 runs=$runs', 'frame %.2f' % d['ms_per_frame'], {k: round(v,2) for k,v in d['sweep_stage_ms'].items()}, 'cand %.1f' % d['candidates_per_sphere'], 'hit', round(d['hit_fraction'],6))" | tee -a gpurun_out/r2o.log
  echo "== config 4 runs=$runs" | tee -a gpurun_out/r2o.log
  XP_LEAF_RUNS=$runs timeout 400 python bench.py --config 4 --no-cpu 2>gpurun_out/r2o.err | tail -1 > gpurun_out/r2o_c4_$runs.json
  python -c "
import json
d=json.loads(open('gpurun_out/r2o_c4_$runs.json').read())
print('config4 runs=$runs', d['value'], d['unit'], 'ms', d['ms_per_step'], {k:v for k,v in d['detail'].items() if isinstance(v,(int,float))})" | tee -a gpurun_out/r2o.log
done
