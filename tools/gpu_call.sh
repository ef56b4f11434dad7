#!/bin/bash
# One gpurun session: everything writes under gpurun_out/. Usage: gpurun -- 'bash tools/gpu_call.sh <tag> <steps...>'
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
tag=$1; shift
for step in "$@"; do
  case $step in
    tests)   timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?" ;;
    ubench)  timeout 300 tools/ubench/i8_peak > gpurun_out/${tag}_i8_peak.jsonl 2> gpurun_out/${tag}_i8_peak.err; echo "ubench rc=$?" ;;
    bench)   timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?" ;;
    smoke)   timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" ;;
    *)       echo "unknown step $step" ;;
  esac
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.csv 2>&1
tail -5 gpurun_out/${tag}_tests.log 2>/dev/null
