#!/bin/bash
# A/B of programmatic dependent launch (F2D_PDL=0/1): GPU suite, V-cycle by level, bench lines.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/r02_pdl_tests.log
tail -3 $O/r02_pdl_tests.log
for p in 0 1; do
  F2D_PDL=$p timeout 200 python tools/vcycle_by_level.py 4096 > $O/r02_vcycle_pdl$p.txt 2>&1
  F2D_PDL=$p timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-s5 > $O/r02_bench_pdl$p.json 2> $O/r02_bench_pdl$p.err
  F2D_PDL=$p timeout 300 python bench.py --config rb --steps 10 --warmup 3 --no-cpu > $O/r02_bench_rb_pdl$p.json 2> $O/r02_bench_rb_pdl$p.err
  F2D_PDL=$p timeout 300 python bench.py --config vk --steps 10 --warmup 3 --no-cpu > $O/r02_bench_vk_pdl$p.json 2> $O/r02_bench_vk_pdl$p.err
done
paste $O/r02_vcycle_pdl0.txt $O/r02_vcycle_pdl1.txt | cut -c1-200
for f in $O/r02_bench_pdl0.json $O/r02_bench_pdl1.json $O/r02_bench_rb_pdl0.json $O/r02_bench_rb_pdl1.json $O/r02_bench_vk_pdl0.json $O/r02_bench_vk_pdl1.json; do
  python - $f <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], d.get('ms_per_step'), d.get('value'), d.get('vcycle_ms'))
PY
done
