#!/bin/bash
# One gpurun call (1 GPU): compute-sanitizer passes over a small job of every kernel family, and the ncu counters of the
# hash strategy on low-cardinality input (BASELINE config 5's shape at 1e9 bases) — SURVEY.md §5, VERDICT r1 item 9.
mkdir -p gpurun_out
python tools/sanitize_job.py > gpurun_out/r02_sanitize_plain.log 2>&1; echo "plain rc=$?"
for tool in memcheck racecheck synccheck; do
  small=""; [ $tool = racecheck ] && small="--small"
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_job.py $small > gpurun_out/r02_sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; tail -n 3 gpurun_out/r02_sanitize_$tool.log
done
timeout 600 python bench.py --workload cfg5 --bases 1e9 --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/r02_cfg5_1e9_bench.json 2> gpurun_out/r02_cfg5_1e9_bench.err; echo "cfg5 bench rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hash_ -s 20 -c 8 -o gpurun_out/r02_hash \
  python bench.py --workload cfg5 --bases 1e9 --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/r02_ncu_hash.log 2>&1; echo "ncu hash rc=$?"
ls -la gpurun_out | tail -12
