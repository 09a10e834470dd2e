#!/bin/bash
# Opening experiment of the next round: build the prepared variants (profiles/r01_notes.md, "Prepared, compiled, not
# yet measured") next to the default build and print the gpurun commands that measure them.
#   tools/ab_round2.sh          # builds ab_libs/{cur,wide,pf,wide_pf,cap64,wide_cap64,tma}.so here (no GPU needed)
set -e
cd "$(dirname "$0")/.."
tools/ab_build.sh cur &
tools/ab_build.sh wide -DKMC_PART1_WIDE=1 &
tools/ab_build.sh cap64 -DKMC_FINE_CAP64=5632 -DKMC_FINE_TARGET64=4700 -DKMC_FINISH_MINB64=3 &
tools/ab_build.sh wide_cap64 -DKMC_PART1_WIDE=1 -DKMC_FINE_CAP64=5632 -DKMC_FINE_TARGET64=4700 -DKMC_FINISH_MINB64=3 &
tools/ab_build.sh tma -DKMC_ROUTE_TMA=1 &
tools/ab_build.sh pf -DKMC_PART1_PREFETCH=1 &
tools/ab_build.sh wide_pf -DKMC_PART1_WIDE=1 -DKMC_PART1_PREFETCH=1 &
wait
cat <<'MSG'

# 1 GPU: step time, per-phase times and digest equality of every variant (k = 21, 31, 19), ~40 s of box time
gpurun --timeout 240 -- 'python tools/ab.py --k 21,31,19 ab_libs/cur.so ab_libs/wide.so ab_libs/pf.so ab_libs/wide_pf.so ab_libs/cap64.so ab_libs/wide_cap64.so > gpurun_out/ab_r2.jsonl 2> gpurun_out/ab_r2.err; tail -n 3 gpurun_out/ab_r2.err'
python tools/ab_report.py gpurun_out/ab_r2.jsonl

# 1 GPU: the routing kernel with bulk stores against the oracle (kmc_route goes through the same kernel)
gpurun --timeout 300 -- 'KMC_LIB=$PWD/ab_libs/tma.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_fastpath.py -x -q -m gpu -k "route or dist or key_array" > gpurun_out/tma_tests.log 2>&1; tail -n 3 gpurun_out/tma_tests.log'

# 1 GPU: the low-cardinality multi-GPU route (kmc_table_route + kmc_ingest_pairs) with emulated ranks, vs the oracle
gpurun --timeout 200 -- 'python tools/check_combine.py 3 31 && python tools/check_combine.py 2 21 && python tools/check_combine.py 4 32'

# 2 GPUs: routing pass with and without bulk stores (phases_ms.route in the JSON line)
gpurun --gpus 2 --timeout 200 -- 'for v in cur tma; do KMC_LIB=$PWD/ab_libs/$v.so python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu > gpurun_out/n2_$v.json 2> gpurun_out/n2_$v.err; done'
MSG
