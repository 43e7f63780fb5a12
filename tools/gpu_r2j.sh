#!/bin/bash
# round 2, call J: deferred face queue in the narrowphase (XP_NP_DEFER): suite, A/B exact + fast
mkdir -p gpurun_out
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
sed -i 's/for rep in 1 2; do/for rep in 1; do/' tools/ab_variants.sh
echo "== defer A/B exact"; AB_ARGS="--no-extras" bash tools/ab_variants.sh run 2>&1 | tee gpurun_out/ab.log
echo "== defer A/B fast"; AB_ARGS="--no-extras --arith fast" bash tools/ab_variants.sh run 2>&1 | tee -a gpurun_out/ab.log
echo "== defer A/B soup"; AB_ARGS="--no-extras --scene soup" bash tools/ab_variants.sh run 2>&1 | tee -a gpurun_out/ab.log
