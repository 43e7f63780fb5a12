#!/bin/bash
# One gpurun call: smoke, GPU parity tests, bench, optionally (`stress`) the stress-parity tool, then (only if
# the plain bench exited 0 and neither `noncu` nor `stress` was given) the ncu
# launch list and one --set full capture of the dominant kernel.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== smoke" ; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
BRC=$?
tail -c 6000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "$1" = "variants" ] || [ "$2" = "variants" ]; then
  for v in "--arith fast" "--method wide_pairs" "--method wide_fused" "--method lbvh_fused --prune 1"; do
    echo "== bench $v"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu $v > "gpurun_out/bench_$(echo $v | tr -d ' -').json" 2>> gpurun_out/bench.err
    tail -c 1500 "gpurun_out/bench_$(echo $v | tr -d ' -').json"
  done
fi
if [ "$1" = "stress" ] || [ "$2" = "stress" ]; then
  echo "== stress parity (mid-size adversarial scenes against the oracle)"
  timeout 900 python tools/stress_parity.py --cases 18 --seed $RANDOM 2>&1 | tail -4 | tee gpurun_out/stress.log
  timeout 600 python tools/stress_parity.py --cases 12 --seed $RANDOM --far-exact 1 --kinds 2,4,5 2>&1 | tail -2 | tee -a gpurun_out/stress.log
fi
if [ $BRC -eq 0 ] && [ "$1" != "noncu" ] && [ "$1" != "stress" ]; then
  echo "== ncu launch list"
  timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 &&
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_list.log 2>&1
  echo "launch list rc=$?"
  echo "== ncu full (k_traverse)"
  timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain2.log 2>&1 &&
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"k_broadphase_grid|k_narrow_pairs" -s 4 -c 4 -f -o gpurun_out/prof_twostage \
      python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
  echo "full rc=$?"
fi
ls -la gpurun_out | tail -20
