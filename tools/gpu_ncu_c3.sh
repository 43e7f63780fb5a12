#!/bin/bash
# ncu --set full of the build kernels matching $1 during `bench.py --config 3` -> gpurun_out/prof_$2.ncu-rep
K="$1"; TAG="${2:-c3}"; SKIP="${3:-0}"; CNT="${4:-2}"
mkdir -p gpurun_out
timeout 600 python bench.py --config 3 --no-cpu > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$K" -s $SKIP -c $CNT -f -o gpurun_out/prof_$TAG \
    python bench.py --config 3 --no-cpu > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"
