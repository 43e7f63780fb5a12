#!/bin/bash
# round 2, call C: grid walk v2 (distance slack + window cull) and the device-resident response loop
mkdir -p gpurun_out
echo "== grid + response tests" ; timeout 900 python -m pytest tests/test_gpu_grid.py tests/test_gpu_parity.py -x -q -s -k "grid or response" 2>&1 | tail -15 | tee gpurun_out/grid.log
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
for v in "--scene heightmap --arith exact" "--scene soup --arith exact" "--scene heightmap --arith fast"; do
  tag=$(echo $v | tr -d ' -')
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu $v > gpurun_out/bench_$tag.json 2>> gpurun_out/bench.err
  python tools/bench_summary.py "$v" < gpurun_out/bench_$tag.json
done
XP_RESPONSE_SYNC=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_syncloop.json 2>> gpurun_out/bench.err
python tools/bench_summary.py "sync response loop" < gpurun_out/bench_syncloop.json
echo "== knobs"
sed -i 's/for rep in 1 2; do/for rep in 1; do/' tools/ab_variants.sh
bash tools/ab_variants.sh run 2>&1 | tee gpurun_out/ab.log
echo "== ncu full: grid walk + narrowphase"
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_broadphase_grid|k_narrow_pairs" -s 8 -c 4 -f -o gpurun_out/prof_r02c \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"
tail -5 gpurun_out/bench.err
