#!/bin/bash
# round 2, call D: full suite, the new bench line (all blocks), narrowphase occupancy A/B
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
echo "== bench (default)"
/usr/bin/time -v timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "rc=$?"; grep -E "Elapsed|Maximum resident" gpurun_out/bench_default.err; tail -3 gpurun_out/bench_default.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
print("value %.4g ms/step %.4f e2e %.4g (%.3f ms, frac of pcie ceiling %.3f) resolved %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["frac_of_pcie_ceiling"], d["resolved"]["value"]))
print("stage", {k: round(v,3) for k,v in d["stage_ms"].items() if v>0})
print("e2e full_records %.4g single_call %.4g ceiling" % (d["e2e"]["full_records"]["value"], d["e2e"]["single_call"]["value"]), {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["e2e"]["pcie_ceiling"].items() if k!="note"})
print("arith", json.dumps(d.get("arith"))[:600])
print("sustained", d.get("sustained"))
for k,v in d.get("configs",{}).items(): print("config",k, json.dumps(v)[:700])
print("roofline frac", d["roofline"]["frac"], "clocks", d["clocks"])
PY
echo "== knobs"
sed -i 's/for rep in 1 2; do/for rep in 1; do/' tools/ab_variants.sh
AB_ARGS="--no-extras" bash tools/ab_variants.sh run 2>&1 | tee gpurun_out/ab.log
AB_ARGS="--no-extras --arith fast" bash tools/ab_variants.sh run 2>&1 | tee -a gpurun_out/ab.log
