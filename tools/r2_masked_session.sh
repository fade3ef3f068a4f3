#!/bin/bash
# masked / stored-coefficient paths: parity tests, then the S3 / S4 bench lines (tag = $1)
set -u
TAG=${1:-x}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_multigrid.py tests/test_gpu_live_oracle.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -12 > $O/r02_${TAG}_tests.log
tail -4 $O/r02_${TAG}_tests.log
for c in rb vk; do
  timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu > $O/r02_bench_${c}_${TAG}.json 2> $O/r02_bench_${c}_${TAG}.err
  python - $O/r02_bench_${c}_${TAG}.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], d.get('ms_per_step'), d.get('value'))
        r=d['roofline']
        for k in (r.get('kernels') or d.get('kernels'))[:16]:
            print("   %-50s n=%5.1f iso=%s us/step=%.1f"%(k['kernel'],k['launches_per_step'],k.get('isolated_us') and round(k['isolated_us'],1),k['us_per_step']))
PY
done
