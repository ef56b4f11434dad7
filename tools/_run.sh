cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s2v_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/s2v_tests.log
bash tools/gpu_ab.sh s2v "I8IE_X=1" "I8IE_PDL_MASK=0" "I8IE_X=2"
for c in simple_conv fc_mnist; do for m in 0 90; do I8IE_PDL_MASK=$m timeout 200 python bench.py --config $c --no-hbm-kernels --no-cpu-baseline > gpurun_out/s2v_${c}_$m.json 2>/dev/null; python tools/bench_brief.py gpurun_out/s2v_${c}_$m.json | head -1; done; done
