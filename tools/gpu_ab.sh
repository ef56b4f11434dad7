#!/bin/bash
# A/B runs of bench.py under different environment switches: tools/gpu_ab.sh <tag> "<ENV=..> <ENV=..>" ...
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
tag=$1; shift
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 300 python bench.py --no-hbm-kernels --no-cpu-baseline --steps 50 > gpurun_out/${tag}_ab$i.json 2> gpurun_out/${tag}_ab$i.err
  echo "variant $i [$envs] rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${tag}_ab$i.json'))
    print('  ms_per_step %.4f  img/s %.0f  parity %s  launches/step %d' % (d['ms_per_step'], d['value'], d['parity']['bit_identical'], d['gpu_launches']//d['steps']))
    print('  ' + '  '.join('%s %.1f' % (l['layer'], l['us']) for l in d['layers']))
except Exception as e:
    print('  ERR', e)
PY
done
