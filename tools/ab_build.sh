#!/bin/bash
# Build libkmc variants side by side for tools/ab.py:   tools/ab_build.sh NAME [-DKMC_...=...]...   → ab_libs/NAME.so
# (ab_libs/ is git-ignored through *.so but travels to the GPU box with gpurun.)
set -e
cd "$(dirname "$0")/.."
mkdir -p ab_libs
name=$1; shift
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared \
  -Xptxas -v "$@" -o ab_libs/$name.so k-mer-count_b200/csrc/kmc_api.cu > ab_libs/$name.ptxas 2>&1
grep -A2 "fast_finish_kernelIjE\|fast_part2_kernelImjE" ab_libs/$name.ptxas | grep -E "Used|spill" | sed 's/ptxas info    ://' | tr '\n' ' '; echo
