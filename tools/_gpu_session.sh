cd $GRAFT_REPO_ROOT
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python tools/step_trace.py > gpurun_out/step_trace_b100.txt 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-hbm-kernels --profiler-range > gpurun_out/bench_small.json 2>gpurun_out/bench_small.err && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r01_step_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-hbm-kernels --profiler-range > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc_igemm2 -c 2 -o gpurun_out/r01_top_kernel -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-hbm-kernels --profiler-range > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log | cut -c1-200
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json | cut -c1-600
