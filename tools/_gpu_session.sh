cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -k "stem" 2>&1 | tail -8
python tools/stem_bench.py 2>&1 | tail -1
I8IE_NO_STEM_FUSEQ=1 python tools/stem_bench.py 2>&1 | tail -1
