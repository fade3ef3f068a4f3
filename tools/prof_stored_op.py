"""ncu driver for the stored-coefficient kernels on small levels: an obstacle hierarchy of
ny x nx cells, operator `kind` (f2d_mg_bench_op) on level `lev`, profiled between
cudaProfilerStart/Stop.   python tools/prof_stored_op.py ny nx lev kind"""
import ctypes
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import torch  # noqa: E402
import gpu_util as g  # noqa: E402
import test_gpu_multigrid as T  # noqa: E402
from fluid2d_b200 import _lib  # noqa: E402

ny, nx, lev, kind = [int(a) for a in sys.argv[1:5]]
lib = _lib.lib(strict=False)
ref, h, rng = T.make(lib, "obstacle", ny, nx)
s = g.stream()
print("modes", [lib.mg_level_matrix_mode(h, l) for l in range(lib.mg_nlevels(h))], file=sys.stderr)
lib.mg_fcycle(h, 0, s)
lib.mg_bench_op(h, kind, lev, 3, s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
lib.mg_bench_op(h, kind, lev, 20, s)
e1.record()
torch.cuda.synchronize()
print("kind %d lev %d: %.2f us per launch" % (kind, lev, 1e3*e0.elapsed_time(e1)/20), file=sys.stderr)
torch.cuda.cudart().cudaProfilerStart()
lib.mg_bench_op(h, kind, lev, 2, s)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
