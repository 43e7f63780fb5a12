#!/bin/bash
# One ncu --set full capture (with source) of the kernels matching $1 during a short bench run; extra bench args in $2.
#   gpurun -- 'bash tools/gpu_ncu.sh "k_broadphase_grid2" "--no-extras"'   -> gpurun_out/prof_$3.ncu-rep
K="$1"; ARGS="$2"; TAG="${3:-cap}"; SKIP="${4:-4}"; CNT="${5:-1}"
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu $ARGS > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$K" -s $SKIP -c $CNT -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu $ARGS > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/prof_$TAG.ncu-rep
