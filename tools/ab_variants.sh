#!/bin/bash
# A/B of compile-time knobs.  Here (no GPU):   tools/ab_variants.sh build NAME "-DXP_BP_SSTACK=8" ...
# builds build_ab/libxp_NAME.so per (NAME, FLAGS) pair; on the GPU box:  tools/ab_variants.sh run
# benches every build_ab/*.so (XP_PHYSICS_CUDA_SO override) and prints ms/step + stage times.
set -e
cd "$(dirname "$0")/.."
CS=xp-game-engine_b200/csrc
if [ "$1" = "build" ]; then
  shift; mkdir -p build_ab
  while [ $# -ge 2 ]; do
    name=$1; flags=$2; shift 2
    /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --prec-div=true --prec-sqrt=true \
      --ftz=false -Xcompiler -fPIC -Xcompiler -fvisibility=default -Xptxas -v $flags -shared \
      -o build_ab/libxp_$name.so $CS/xp_api.cu $CS/xp_build.cu $CS/xp_query.cu $CS/xp_debug.cu 2> build_ab/build_$name.log \
      || { tail -20 build_ab/build_$name.log; exit 1; }
    echo "built $name ($flags)"
  done
else
  for so in build_ab/*.so; do
    for rep in 1 2; do
      XP_PHYSICS_CUDA_SO=$PWD/$so timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu ${AB_ARGS} > /tmp/ab.json 2>/tmp/ab.err || { echo "$so FAILED"; tail -3 /tmp/ab.err; continue; }
      python - "$so" <<'EOF'
import json, sys
d = json.load(open('/tmp/ab.json'))
s = d['stage_ms']
print(f"{sys.argv[1]:44s} step {d['ms_per_step']:.3f} ms  trav {s['traverse']:.3f}  narrow {s['narrowphase']:.3f}  e2e {d['e2e']['ms_per_step']:.3f}  hits {d['e2e']['hits']}")
EOF
    done
  done
fi
