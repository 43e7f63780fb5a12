#!/bin/bash
# round 2, call E: f4 on the device, then the full default bench line
mkdir -p gpurun_out
echo "== f4 tests" ; timeout 900 python -m pytest tests/test_gpu_fauerby.py -x -q 2>&1 | tail -15 | tee gpurun_out/f4.log
echo "== bench (default)"
SECONDS=0
timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "rc=$? wall=${SECONDS}s"; tail -3 gpurun_out/bench_default.err | cut -c1-400
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
print("value %.4g ms/step %.4f e2e %.4g (%.3f ms, frac of pcie ceiling %.3f) resolved %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["frac_of_pcie_ceiling"], d["resolved"]["value"]))
print("stage", {k: round(v,3) for k,v in d["stage_ms"].items() if v>0})
print("e2e full_records %.4g single_call %.4g ceiling" % (d["e2e"]["full_records"]["value"], d["e2e"]["single_call"]["value"]), {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["e2e"]["pcie_ceiling"].items() if k!="note"})
print("arith", json.dumps(d.get("arith"))[:700])
print("sustained", d.get("sustained"))
for k,v in d.get("configs",{}).items(): print("config",k, json.dumps(v)[:900])
print("roofline frac", d["roofline"]["frac"], "clocks", d["clocks"], "cpu", json.dumps(d["cpu_baseline"])[:300])
PY
