cd $GRAFT_REPO_ROOT
for d in 0 16 32 48 64 96; do echo "DBG=$d $(I8IE_STEM2_DBG=$d python tools/layer_bench.py --layers conv1 --batch 100 2>&1 | grep conv1)"; done
