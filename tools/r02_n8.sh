#!/bin/bash
# N GPUs (default 8): the weak-scaling step through the default route, with the per-step timeline of rank 0
mkdir -p gpurun_out
N=${1:-8}; shift
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
env "$@" KMC_DIST_PROF=1 timeout 500 $TR --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --no-extra \
    > gpurun_out/r02_n${N}_default.json 2> gpurun_out/r02_n${N}_default.err; echo "rc=$?"
python - <<PY
import json
o=json.load(open("gpurun_out/r02_n${N}_default.json"))
print(round(o["value"],2), "Gk/s", round(o["ms_per_step"],2), "ms e2e", o["e2e"] and round(o["e2e"]["ms_per_step"],2), o["run"]["parallelism"], o["phases_ms"])
PY
grep "kmc dist r0" gpurun_out/r02_n${N}_default.err | sed -n 6,8p | cut -c1-330
