#!/bin/bash
# strong-scaling runs: bash scripts/scale.sh <workload> <N> [extra bench args]
W=$1; N=$2; shift; shift
mkdir -p gpurun_out/scale
if [ "$N" = "1" ]; then
  python bench.py --workload $W --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline "$@" 2>gpurun_out/scale/${W}_$N.err | tail -1 | tee gpurun_out/scale/${W}_$N.json
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --workload $W --gpus $N --steps 20 --warmup 5 --allgather "$@" 2>gpurun_out/scale/${W}_$N.err | tail -1 | tee gpurun_out/scale/${W}_$N.json
fi
