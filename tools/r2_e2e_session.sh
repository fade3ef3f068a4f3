set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_loop_output.py tests/test_gpu_live_oracle.py tests/test_gpu_zz_late.py -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-s5 > $O/r02_bench_e2e1.json 2> $O/r02_bench_e2e1.err
python - <<'PY'
import json
for l in open('gpurun_out/r02_bench_e2e1.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d.get('ms_per_step'), d.get('value'), d.get('e2e'))
PY
