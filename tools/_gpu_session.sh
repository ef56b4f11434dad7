cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
for d in 0 1 2 3 7; do echo "DBG=$d"; I8IE_STEM2_DBG=$d python tools/step_trace.py 2>&1 | tail -18 | sed -n 2,4p; done
python tools/step_trace.py > gpurun_out/step_trace_b100.txt 2>&1; cat gpurun_out/step_trace_b100.txt | tail -19
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
j=json.load(open('gpurun_out/bench.json'))
print(j['value'], j['ms_per_step'], j['e2e']['value'])
print(j['hbm_kernels'])
PY
