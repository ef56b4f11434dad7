cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/step_trace.py 2>&1 | tail -19
python bench.py --no-cpu-baseline --no-hbm-kernels 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('b100', j['value'], j['ms_per_step'], j['e2e']['value'], j['gpu_launches'])"
