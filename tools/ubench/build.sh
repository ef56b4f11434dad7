#!/bin/bash
# Builds the dev microbenchmarks in place (binaries are git-ignored; they travel to the GPU box with gpurun).
set -e
cd "$(dirname "$0")"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -cudart shared -Xlinker -rpath=/usr/local/cuda/lib64"
for src in i8_peak pipe_rates; do
  [ -f $src.cu ] && nvcc $FLAGS -o $src $src.cu
done
ls -la i8_peak pipe_rates 2>/dev/null
