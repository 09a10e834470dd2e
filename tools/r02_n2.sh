#!/bin/bash
# 2 GPUs: multi-rank parity against the oracle, then the weak-scaling step through each route (cfg2 per GPU)
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29533 tools/dist_parity.py --bases 8e6 > gpurun_out/r02_n${N}_parity.jsonl 2> gpurun_out/r02_n${N}_parity.err; echo "parity rc=$?"
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
}
run pipelined KMC_X=0
run hash KMC_DIST_PIPELINE=0
run range KMC_DIST_PARTITION=range
