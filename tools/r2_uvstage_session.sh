set -u
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py tests/test_gpu_live_oracle.py tests/test_gpu_loop_output.py tests/test_gpu_zz_late.py -x -q 2>&1 | tail -6
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-s5 > $O/r02_bench_uv1.json 2> $O/r02_bench_uv1.err
timeout 300 python bench.py --config vk --steps 10 --warmup 3 --no-cpu > $O/r02_bench_vk_uv1.json 2>/dev/null
python - <<'PY'
import json
for fn in ["r02_bench_uv1.json","r02_bench_vk_uv1.json"]:
  for l in open("gpurun_out/"+fn):
    if l.startswith("{"):
        d=json.loads(l); print(fn, d["ms_per_step"], d["value"], d["gpu_launches"])
        for k in d["kernels"]:
            if "orthogradient" in k["kernel"] or "map_vec" in k["kernel"]: print("  ",k["kernel"], k["launches_per_step"], round(k["avg_us"],1), k["frac_of_peak"] and round(k["frac_of_peak"],3))
PY
