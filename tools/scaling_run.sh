#!/bin/bash
# One multi-GPU bench line: tools/scaling_run.sh N  ->  gpurun_out/scale_N.json (run under `gpurun --gpus N`)
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
echo "rc=$?"; tail -c 600 gpurun_out/scale_$N.json; tail -3 gpurun_out/scale_$N.err
