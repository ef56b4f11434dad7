cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s2s_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/s2s_tests.log
