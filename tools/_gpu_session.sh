cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -x -q 2>&1 | tail -8
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo PAIR; timeout 300 python tools/layer_bench.py --batch 100 --layers conv2,conv3,conv4,conv5 2>&1 | grep conv
echo NOCLUSTER; I8IE_NO_CLUSTER=1 timeout 300 python tools/layer_bench.py --batch 100 --layers conv2,conv3,conv4,conv5 2>&1 | grep conv
echo PAIR b1000; timeout 300 python tools/layer_bench.py --batch 1000 --reps 5 --iters 5 --layers conv2,conv3,conv4,conv5 2>&1 | grep conv
echo NOCLUSTER b1000; I8IE_NO_CLUSTER=1 timeout 300 python tools/layer_bench.py --batch 1000 --reps 5 --iters 5 --layers conv2,conv3,conv4,conv5 2>&1 | grep conv
