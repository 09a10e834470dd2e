#!/bin/bash
# One gpurun call: GPU parity tests, a bench line, the ncu launch list of the same command, one ncu --set full capture
# of the partitioned-path kernels.  Outputs under gpurun_out/ (copied into profiles/ by hand after reading).
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest.log )
tail -n 5 gpurun_out/r02_gputest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r02_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/r02_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fast_ -s 16 -c 5 -o gpurun_out/r02_full \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/r02_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
