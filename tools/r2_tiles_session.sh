set -u
bash tools/r2_quick_session.sh tiles1
for c in rb vk; do
  timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_bench_${c}_tiles1.json 2> gpurun_out/r02_bench_${c}_tiles1.err
  python - gpurun_out/r02_bench_${c}_tiles1.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], d.get('ms_per_step'), d.get('value'))
        r=d['roofline']
        for k in (r.get('kernels') or d.get('kernels'))[:14]:
            print("   %-50s n=%5.1f iso=%s us/step=%.1f"%(k['kernel'],k['launches_per_step'],k.get('isolated_us') and round(k['isolated_us'],1),k['us_per_step']))
PY
done
F2D_SMALL_TILE_MAXCTAS=592 timeout 200 python tools/vcycle_by_level.py 4096 2>&1 | head -6
F2D_SMALL_TILE_MAXCTAS=0 timeout 200 python tools/vcycle_by_level.py 4096 2>&1 | head -6
