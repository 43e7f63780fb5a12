#!/bin/bash
# round 2, call F (N GPUs): the multi-GPU tests over every visible GPU, then the default bench line under torchrun
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
nvidia-smi topo -m 2>/dev/null | head -12 > gpurun_out/topo_$N.txt
echo "== multi-GPU tests ($N GPUs)"; timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -s 2>&1 | tail -12 | tee gpurun_out/multi_$N.log
echo "== bench N=$N"
nvidia-smi nvlink -gt d > gpurun_out/nvlink_before_$N.txt 2>&1      # per-link data counters (KiB) around the bench
SECONDS=0
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$? wall=${SECONDS}s"
nvidia-smi nvlink -gt d > gpurun_out/nvlink_after_$N.txt 2>&1
python - $N <<'PY' | tee gpurun_out/nvlink_delta_$N.txt
import re, sys
n = sys.argv[1]
def read(path):
    tot, gpu = {}, None
    for line in open(path, errors="ignore"):
        m = re.match(r"GPU (\d+):", line)
        if m:
            gpu = int(m.group(1)); tot.setdefault(gpu, [0, 0]); continue
        m = re.search(r"Data (Tx|Rx): (\d+) KiB", line)
        if m and gpu is not None:
            tot[gpu][0 if m.group(1) == "Tx" else 1] += int(m.group(2))
    return tot
a, b = read(f"gpurun_out/nvlink_before_{n}.txt"), read(f"gpurun_out/nvlink_after_{n}.txt")
print("# nvidia-smi nvlink -gt d, summed over a GPU's links, after - before the whole `bench.py --gpus N` run (all blocks of the line)")
if "N/A" in open(f"gpurun_out/nvlink_after_{n}.txt", errors="ignore").read() and not any(sum(v) for v in b.values()):
    print("the per-link data counters read N/A in this container (see the raw file)")
for g in sorted(b):
    tx, rx = b[g][0] - a.get(g, [0, 0])[0], b[g][1] - a.get(g, [0, 0])[1]
    print(f"GPU {g}: NVLink data Tx {tx / 1048576:.2f} GiB, Rx {rx / 1048576:.2f} GiB")
PY
grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/bench_n$N.err | tail -5 | cut -c1-400
python - $N <<'PY'
import json, sys
n=sys.argv[1]
d=json.loads(open(f"gpurun_out/bench_n{n}.json").read().strip().splitlines()[-1])
print("N", d["n_gpus"], "value %.4g ms/step %.4f e2e %.4g (%.3f ms, frac of pcie ceiling %.3f) resolved %.4g verified %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["frac_of_pcie_ceiling"], d["resolved"]["value"], d["exchange_verified"]))
print("exchange:", d["arm"]["exchange"][:120])
print("ceiling", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["e2e"]["pcie_ceiling"].items() if k!="note"}, "full_records %.4g" % d["e2e"]["full_records"]["value"])
print("sustained", d.get("sustained"))
for k,v in d.get("configs",{}).items(): print("config",k, json.dumps(v)[:600])
PY
