"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`: the
reference's algorithm on the host cores through the oracle port) prints exactly one JSON
line with the keys the driver reads, and the roofline arithmetic of the GPU arm."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--n", "64",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "cell_updates_per_s" and d["unit"] == "cell-updates/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and abs(d["value"]-64*64/(d["ms_per_step"]*1e-3)) <= 1e-6*d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    e = d["e2e"]
    assert e["value"] == d["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--n", "64", "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=120,
                       env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_algorithmic_bytes_formula():
    sys.path.insert(0, REPO)
    import bench
    # B_alg(T, n_F) = 1114.2 + 144 T + 272.5 n_F
    assert abs(bench.b_alg(1, 2)-1803.2) < 1e-9
    assert abs(bench.b_alg(1, 4)-2348.2) < 1e-9
    assert abs(bench.b_alg(2, 2)-bench.b_alg(1, 2)-144.) < 1e-9
    peak, src = bench.measured_peaks()
    assert peak > 1000.


def test_arms_name_the_same_workload_and_kernel_bytes():
    sys.path.insert(0, REPO)
    import bench
    import argparse
    # N = 1: 4096^2 (S1); N > 1: 16384^2 strong scaling (S5) unless --weak
    a = argparse.Namespace(strong=False, weak=False, replicas=False, n=0)
    assert bench.resolve_workload(a, 1) == (4096, False)
    assert bench.resolve_workload(a, 8) == (16384, True)
    a.weak = True
    assert bench.resolve_workload(a, 8) == (4096, False)
    assert "16384x16384" in bench.workload_string(16384, 1, True)
    # algorithmic bytes per cell of the tagged kernels
    assert bench.kernel_bytes_per_cell("k_smooth2<mode1,input3> 4096x4096") == (26., 4096*4096)
    assert bench.kernel_bytes_per_cell("k_smooth2<mode2,input0,peer> 4096x512") == (25., 4096*512)
    assert bench.kernel_bytes_per_cell("k_resid_restrict<mode1> 2048x2048") == (18., 2048*2048)
    assert bench.kernel_bytes_per_cell("k_adv<upw1,order5,masked0> 4096x4096") == (32., 4096*4096)
    assert bench.kernel_bytes_per_cell("k_map_vec<3 in> 1000 doubles") == (32., 1000)
    assert bench.kernel_bytes_per_cell("k_mg_ctail<program0> 128x128") == (None, None)
