cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s2o_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/s2o_tests.log
python tools/layer_bench.py --layers fc1,fc2 2>&1 | tail -3
python tools/step_trace.py --batch 100 --steps 10 > gpurun_out/s2o_step_trace.txt 2>&1
bash tools/gpu_ab.sh s2o "I8IE_X=1"
