cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for ns in 0 20 50 100 200; do echo "sleep $ns: $(I8IE_WAIT_SLEEP_NS=$ns python tools/stem_bench.py 2>&1 | tail -1)"; done
bash tools/gpu_ab.sh s2q "I8IE_WAIT_SLEEP_NS=0" "I8IE_WAIT_SLEEP_NS=50"
