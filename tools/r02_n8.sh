#!/bin/bash
# N GPUs (default 8): the weak-scaling step (cfg2 per GPU) through the default route, per chunk count, with rank 0's timeline
mkdir -p gpurun_out
N=${1:-8}; shift
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}
  env $(echo $envs | tr ',' ' ') KMC_DIST_PROF=1 timeout 500 $TR --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --no-extra \
    > gpurun_out/r02_n${N}_$name.json 2> gpurun_out/r02_n${N}_$name.err; echo "$name rc=$?"
  python tools/show_bench.py gpurun_out/r02_n${N}_$name.json
  grep "kmc dist r0" gpurun_out/r02_n${N}_$name.err | sed -n 7,7p | cut -c1-200
done
