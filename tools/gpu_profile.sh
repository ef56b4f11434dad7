#!/bin/bash
# Round evidence on one B200 (run under gpurun): bench line, ncu launch lists of the timed steps
# (cold caches = ncu default, and --cache-control none = the forward's own cache state) and one
# `--set full` capture of the tensor-core conv kernels. Outputs land in gpurun_out/.
set -u
cd "${GRAFT_REPO_ROOT:-.}"
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err || { echo "bench failed"; tail -5 $OUT/${TAG}_bench.err; exit 1; }
CMD="python bench.py --steps 3 --warmup 3 --profiler-range --no-cpu-baseline --no-hbm-kernels"
$CMD > $OUT/${TAG}_plain.log 2>&1 || { echo "plain profiling command failed"; exit 1; }
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu1.log 2>&1
ncu --metrics $M --clock-control none --cache-control none --profile-from-start off --csv --log-file $OUT/${TAG}_launches_warm.csv $CMD > $OUT/${TAG}_ncu2.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on --profile-from-start off \
    -k regex:"tc_igemm2_kernel|tc_stem2_kernel|tc_fc_cluster_kernel" -c 7 -o $OUT/${TAG}_prof_conv $CMD > $OUT/${TAG}_ncu3.log 2>&1
tail -2 $OUT/${TAG}_ncu3.log
head -c 600 $OUT/${TAG}_bench.json
