#!/bin/bash
# Multi-GPU bench under gpurun --gpus N: tools/gpu_multi.sh <tag> <N> [extra bench args]
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
tag=$1; n=$2; shift 2
I8IE_BENCH_VERBOSE=120 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
  --master-port 29511 bench.py --gpus $n "$@" > gpurun_out/${tag}_n$n.json 2> gpurun_out/${tag}_n$n.err
echo "bench N=$n rc=$?"
tail -c 600 gpurun_out/${tag}_n$n.err
python tools/bench_brief.py gpurun_out/${tag}_n$n.json
