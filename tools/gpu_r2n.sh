#!/bin/bash
# round 2, call N: fused bottom-up tree build: suite, config 3 (build stages), soup bench
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --config 3 --no-cpu > gpurun_out/c3.json 2>> gpurun_out/bench.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/c3.json").read().strip().splitlines()[-1])["detail"]
print("config3 ms/frame %.2f build %.3f" % (d["ms_per_frame"], d["build_ms"]), {k: round(v,3) for k,v in d["build_stage_ms"].items()})
PY
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-extras --scene soup > gpurun_out/bench_soup.json 2>> gpurun_out/bench.err; python tools/bench_summary.py soup < gpurun_out/bench_soup.json
tail -5 gpurun_out/bench.err
