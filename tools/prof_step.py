"""Profiling driver: the bench workload (Euler freedecay n^2), a few warm steps, then ONE
step between cudaProfilerStart/Stop (run under `ncu --profile-from-start off`)."""
import os
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
import torch  # noqa: E402
import fluid2d_b200  # noqa: E402
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1
graphs = int(sys.argv[3]) if len(sys.argv) > 3 else 1   # 0: no CUDA graphs, so ncu lists the kernels of the solve
api = fluid2d_b200.api()
so = sys.stdout
sys.stdout = sys.stderr
f2d = bench.build_case(api, n, T, tempfile.mkdtemp())
f2d.model.diagnostics(f2d.model.var, 0.)
if not graphs:
    g = f2d.model.ope.gmg
    g.lib.mg_set_graphs(g.h, 0)
for _ in range(3):
    bench.loop_body(f2d)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
bench.loop_body(f2d)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
sys.stdout = so
print("done")
