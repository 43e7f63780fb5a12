#!/bin/bash
# A/B of every build_ab/*.so (tools/ab_variants.sh build ...), bench args in $1, optional grid tests first ($2 = tests)
mkdir -p gpurun_out
sed -i 's/for rep in 1 2; do/for rep in 1; do/' tools/ab_variants.sh
if [ "$2" = "tests" ]; then timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 2>&1 | tail -4; fi
AB_ARGS="--no-extras $1" bash tools/ab_variants.sh run 2>&1 | tee gpurun_out/ab.log
