#!/bin/bash
# round 2, call A: arithmetic self-checks, the GPU suite, and an A/B of the four narrowphase kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== arith" ; timeout 600 python -m pytest tests/test_gpu_arith.py -x -q -s 2>&1 | tail -15 | tee gpurun_out/arith.log
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
for a in exact exact_literal fast fast_literal; do
  echo "== bench --arith $a"
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --arith $a > gpurun_out/bench_$a.json 2>> gpurun_out/bench.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$a.json").read().strip().splitlines()[-1])
    print("$a", "ms/step", round(d["ms_per_step"],4), "value %.4g"%d["value"], "np_ms", round(d["stage_ms"]["narrowphase"],4), "bp_ms", round(d["stage_ms"]["traverse"],4), "roof", round(d["roofline"]["frac"],4), "e2e %.4g"%d["e2e"]["value"], "resolved %.4g"%d["resolved"]["value"])
except Exception as e:
    print("$a failed", e)
PY
done
tail -5 gpurun_out/bench.err
ls -la gpurun_out | tail -12
