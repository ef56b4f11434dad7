cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_fp32.py tests/test_gpu_api.py tests/test_gpu_nets.py -m gpu -x -q 2>&1 | tail -15
