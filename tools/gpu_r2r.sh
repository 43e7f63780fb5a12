#!/bin/bash
# sphere scatter: cooperative 64 B record stores (XP_SCATTER_COOP) A/B + parity tests on the default build
mkdir -p gpurun_out
echo "== tests" | tee gpurun_out/r2r.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_grid.py -q -m gpu -x --timeout 600 2>&1 | tail -4 | tee -a gpurun_out/r2r.log
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
s=d['stage_ms']
print('$1', 'step %.4f ms' % d['ms_per_step'], 'value %.4g' % d['value'], 'walk %.3f narrow %.3f sort %.4f fin %.3f' % (s['traverse'], s['narrowphase'], s['sphere_sort'], s['finalize']), 'e2e %.4f' % d['e2e']['ms_per_step'])"; }
for rep in 1 2; do
for so in "" build_ab/libxp_coop0.so; do
  XP_PHYSICS_CUDA_SO=${so:+$PWD/$so} timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-extras 2>gpurun_out/r2r.err > gpurun_out/r2r_tmp.json \
    && show "coop=${so:-1(default)}" < gpurun_out/r2r_tmp.json | tee -a gpurun_out/r2r.log || tail -3 gpurun_out/r2r.err
done
done
