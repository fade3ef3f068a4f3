set -u
O=gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r02_bench_v9.json 2> $O/r02_bench_v9.err
for c in rb vk; do timeout 300 python bench.py --config $c --steps 10 --warmup 3 > $O/r02_bench_${c}_v9.json 2> $O/r02_bench_${c}_v9.err; done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference_v9.json 2>&1
python - <<'PY'
import json
for fn in ['r02_bench_v9.json','r02_bench_rb_v9.json','r02_bench_vk_v9.json','r02_bench_reference_v9.json']:
  for l in open('gpurun_out/'+fn):
    if l.startswith('{'):
        d=json.loads(l); print(fn, d.get('ms_per_step'), d.get('value'), d.get('e2e'))
PY
