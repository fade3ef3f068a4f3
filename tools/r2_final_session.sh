#!/bin/bash
# final single-GPU session of round 2: GPU suite, smoke, the bench lines (tag = $1)
set -u
TAG=${1:-v11}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/r02_gpu_tests_$TAG.log
tail -3 $O/r02_gpu_tests_$TAG.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python bench.py > $O/r02_bench_$TAG.json 2> $O/r02_bench_$TAG.err
for c in rb vk; do
  timeout 300 python bench.py --config $c --steps 10 --warmup 3 > $O/r02_bench_${c}_$TAG.json 2> $O/r02_bench_${c}_$TAG.err
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference_$TAG.json 2>&1
python - $TAG <<'PY'
import json,sys
t=sys.argv[1]
for fn in ['r02_bench_%s.json'%t,'r02_bench_rb_%s.json'%t,'r02_bench_vk_%s.json'%t,'r02_bench_reference_%s.json'%t]:
  for l in open('gpurun_out/'+fn):
    if l.startswith('{'):
        d=json.loads(l); print(fn, d.get('steps'), d.get('ms_per_step'), d.get('value'), d.get('e2e'), (d.get('s5_one_gpu') or {}).get('ms_per_step'))
PY
