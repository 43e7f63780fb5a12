#!/bin/bash
# round 2, call I: ray setup fused into the sphere scatter, pair prefetch A/B
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
sed -i 's/for rep in 1 2; do/for rep in 1; do/' tools/ab_variants.sh
echo "== prefetch A/B exact"; AB_ARGS="--no-extras" bash tools/ab_variants.sh run 2>&1 | tee gpurun_out/ab.log
echo "== prefetch A/B fast"; AB_ARGS="--no-extras --arith fast" bash tools/ab_variants.sh run 2>&1 | tee -a gpurun_out/ab.log
echo "== prefetch A/B soup"; AB_ARGS="--no-extras --scene soup" bash tools/ab_variants.sh run 2>&1 | tee -a gpurun_out/ab.log
for v in "--scene heightmap" ; do
  tag=$(echo $v | tr -d ' -')
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-extras $v > gpurun_out/bench_$tag.json 2>> gpurun_out/bench.err
  python tools/bench_summary.py "$v" < gpurun_out/bench_$tag.json
done
tail -5 gpurun_out/bench.err
