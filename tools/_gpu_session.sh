cd $GRAFT_REPO_ROOT
python tools/e2e_probe.py 2>&1 | tail -6
