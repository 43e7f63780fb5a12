#!/bin/bash
# round 2, call K: cell-parallel grid walk (k_broadphase_grid2): grid + parity tests, A/B against the lane-per-ray walk
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
sed -i 's/for rep in 1 2; do/for rep in 1; do/' tools/ab_variants.sh
echo "== A/B exact"; AB_ARGS="--no-extras" bash tools/ab_variants.sh run 2>&1 | tee gpurun_out/ab.log
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; python tools/bench_summary.py default < gpurun_out/bench.json; tail -3 gpurun_out/bench.err
