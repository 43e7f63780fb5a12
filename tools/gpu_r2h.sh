#!/bin/bash
# round 2, call H: after reverting the entry-time cull: suite, bench (heightmap/soup, exact/fast), LBVH lane-sort A/B on the soup
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
for v in "--scene heightmap" "--scene soup" "--scene heightmap --arith fast" "--scene soup --arith fast"; do
  tag=$(echo $v | tr -d ' -')
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-extras $v > gpurun_out/bench_$tag.json 2>> gpurun_out/bench.err
  python tools/bench_summary.py "$v" < gpurun_out/bench_$tag.json
done
echo "== LBVH lane-sort A/B (soup)"
sed -i 's/for rep in 1 2; do/for rep in 1; do/' tools/ab_variants.sh
AB_ARGS="--no-extras --scene soup" bash tools/ab_variants.sh run 2>&1 | tee gpurun_out/ab.log
timeout 600 python bench.py --config 3 --no-cpu > gpurun_out/c3.json 2>> gpurun_out/bench.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/c3.json").read().strip().splitlines()[-1])["detail"]
print("config3 ms/frame %.2f build %.3f" % (d["ms_per_frame"], d["build_ms"]), {k: round(v,3) for k,v in d["build_stage_ms"].items()})
PY
tail -5 gpurun_out/bench.err
