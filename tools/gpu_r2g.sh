#!/bin/bash
# round 2, call G: full suite with the entry-time cull, bench (heightmap + soup), window-size A/B, ncu of the build kernels
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
timeout 300 python -m pytest tests/test_gpu_parity.py -q -s -k "entry_time_cull" 2>&1 | grep -E "tests exhaustive|passed|failed"
for v in "--scene heightmap" "--scene soup" "--scene heightmap --prune 1" "--scene soup --prune 1"; do
  tag=$(echo $v | tr -d ' -')
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-extras $v > gpurun_out/bench_$tag.json 2>> gpurun_out/bench.err
  python tools/bench_summary.py "$v" < gpurun_out/bench_$tag.json
done
echo "== knobs"
sed -i 's/for rep in 1 2; do/for rep in 1; do/' tools/ab_variants.sh
AB_ARGS="--no-extras" bash tools/ab_variants.sh run 2>&1 | tee gpurun_out/ab.log
echo "== ncu: build kernels of config 3 (10M triangles)"
timeout 600 python bench.py --config 3 --no-cpu > gpurun_out/plain_c3.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none -k regex:"k_rs_|k_refit|k_pack|k_karras|k_bounds|k_morton_tris|k_scan" -s 40 -c 24 -f -o gpurun_out/prof_r02g_build \
    python bench.py --config 3 --no-cpu > gpurun_out/ncu_build.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/plain_c3.log | cut -c1-600
tail -5 gpurun_out/bench.err
