#!/bin/bash
# config-3 build stages for every build_ab/*.so
mkdir -p gpurun_out
for so in build_ab/*.so; do
  XP_PHYSICS_CUDA_SO=$PWD/$so timeout 600 python bench.py --config 3 --no-cpu > /tmp/c3.json 2>/tmp/c3.err || { echo "$so FAILED"; tail -3 /tmp/c3.err; continue; }
  python - "$so" <<'PY'
import json, sys
d=json.loads(open("/tmp/c3.json").read().strip().splitlines()[-1])["detail"]
print("%-40s frame %.2f build %.3f" % (sys.argv[1], d["ms_per_frame"], d["build_ms"]), {k: round(v,3) for k,v in d["build_stage_ms"].items()})
PY
done | tee gpurun_out/ab_c3.log
