#!/bin/bash
# N GPUs (default 2): multi-rank parity against the oracle, the CLI at N ranks, then the weak-scaling step through each route
mkdir -p gpurun_out
N=${1:-2}; shift
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ -z "$SKIP_TESTS" ]; then
timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q > gpurun_out/r02_n${N}_multirank.log 2>&1; echo "multirank tests rc=$?"; tail -n 15 gpurun_out/r02_n${N}_multirank.log
fi
run() { # name env...
  name=$1; shift
  env "$@" KMC_DIST_PROF=1 timeout 400 $TR --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --no-extra \
    > gpurun_out/r02_n${N}_$name.json 2> gpurun_out/r02_n${N}_$name.err; echo "$name rc=$?"
  python - <<PY
import json
try:
    o=json.load(open("gpurun_out/r02_n${N}_$name.json"))
    print("$name", round(o["value"],2), "Gk/s", round(o["ms_per_step"],2), "ms e2e", o["e2e"] and round(o["e2e"]["ms_per_step"],2), o["run"]["parallelism"], o["phases_ms"])
except Exception as e: print("$name failed", e)
PY
  grep "kmc dist r0" gpurun_out/r02_n${N}_$name.err | sed -n 7,8p | cut -c1-200
}
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}
  run $name $(echo $envs | tr ',' ' ')
done
