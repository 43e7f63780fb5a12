#!/bin/bash
# round 2, call B: the height-grid walk -- parity, then bench (heightmap vs soup hand-over, exact vs fast), then knob A/B
mkdir -p gpurun_out
echo "== grid tests" ; timeout 900 python -m pytest tests/test_gpu_grid.py -x -q -s 2>&1 | tail -15 | tee gpurun_out/grid.log
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
for v in "--scene heightmap --arith exact" "--scene soup --arith exact" "--scene heightmap --arith fast" "--scene soup --arith fast"; do
  tag=$(echo $v | tr -d ' -')
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu $v > gpurun_out/bench_$tag.json 2>> gpurun_out/bench.err
  python tools/bench_summary.py "$v" < gpurun_out/bench_$tag.json
done
echo "== knobs"
sed -i 's/for rep in 1 2; do/for rep in 1; do/' tools/ab_variants.sh
bash tools/ab_variants.sh run 2>&1 | tee gpurun_out/ab.log
tail -5 gpurun_out/bench.err
