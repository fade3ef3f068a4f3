#!/bin/bash
# quick single-GPU session: multigrid parity tests, V-cycle by level, S1 bench line (tag = $1)
set -u
TAG=${1:-x}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_multigrid.py tests/test_gpu_fullsize.py tests/test_gpu_live_oracle.py tests/test_gpu_golden.py -x -q 2>&1 | tail -15 > $O/r02_${TAG}_tests.log
tail -4 $O/r02_${TAG}_tests.log
timeout 200 python tools/vcycle_by_level.py 4096 > $O/r02_${TAG}_vcycle.txt 2>&1
cat $O/r02_${TAG}_vcycle.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-s5 > $O/r02_bench_${TAG}.json 2> $O/r02_bench_${TAG}.err
python - $O/r02_bench_${TAG}.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(d.get('ms_per_step'), d.get('value'), d.get('vcycle_ms'))
        r=d['roofline']; print(r['kernel'], r['frac'], r['ms_per_launch'])
        for k in (r.get('kernels') or d.get('kernels'))[:14]:
            print("%-50s n=%5.1f iso=%s frac=%s us/step=%.1f"%(k['kernel'],k['launches_per_step'],k.get('isolated_us') and round(k['isolated_us'],1),k.get('frac_of_peak') and round(k['frac_of_peak'],3),k['us_per_step']))
PY
