#!/bin/bash
# 256-bit record stores (XP_STG256) and the scatter's store form (XP_SCATTER_COOP 1 = shared-memory staging, 2 = STG.256): A/B
mkdir -p gpurun_out
echo "== tests (default build)" | tee gpurun_out/r2s.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_lbvh.py tests/test_gpu_grid.py -q -m gpu -x --timeout 600 2>&1 | tail -3 | tee -a gpurun_out/r2s.log
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
s=d['stage_ms']
print('$1', 'step %.4f ms' % d['ms_per_step'], 'walk %.3f narrow %.3f sort %.4f' % (s['traverse'], s['narrowphase'], s['sphere_sort']), 'e2e %.4f' % d['e2e']['ms_per_step'])"; }
for so in "" build_ab/libxp_coop2.so build_ab/libxp_stg0.so ""; do
  XP_PHYSICS_CUDA_SO=${so:+$PWD/$so} timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-extras 2>gpurun_out/r2s.err > gpurun_out/r2s_tmp.json \
    && show "${so:-default}" < gpurun_out/r2s_tmp.json | tee -a gpurun_out/r2s.log || tail -3 gpurun_out/r2s.err
done
for so in "" build_ab/libxp_stg0.so; do
  XP_PHYSICS_CUDA_SO=${so:+$PWD/$so} timeout 400 python bench.py --config 3 --no-cpu 2>gpurun_out/r2s.err | tail -1 > gpurun_out/r2s_c3.json
  python -c "
import json
d=json.loads(open('gpurun_out/r2s_c3.json').read())['detail']
print('config3 ${so:-default}', 'frame %.2f build %.3f' % (d['ms_per_frame'], d['build_ms']), {k: round(v,3) for k,v in d['build_stage_ms'].items()})" | tee -a gpurun_out/r2s.log
done
