cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 rc=$?"
tail -2 gpurun_out/bench_n2.err
python - <<'PY'
import json
for f in ['gpurun_out/bench_n2.json']:
    j=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, j['value'], j['ms_per_step'], j['e2e']['value'], j['config']['per_gpu_batch'])
PY
python bench.py --batch 500 --steps 30 --no-cpu-baseline --no-hbm-kernels 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('b500 1gpu', j['value'], j['ms_per_step'])"
