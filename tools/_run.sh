cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for b in 16 64 100 125 128 200 250 500; do
  echo "batch $b cluster: $(python tools/layer_bench.py --batch $b --layers fc1,fc2 2>&1 | grep -E '^fc' | awk '{print $1, $3}' | tr '\n' ' ')"
  echo "batch $b old    : $(I8IE_NO_FC_CLUSTER=1 python tools/layer_bench.py --batch $b --layers fc1,fc2 2>&1 | grep -E '^fc' | awk '{print $1, $3}' | tr '\n' ' ')"
done
