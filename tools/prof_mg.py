"""Profiling driver (used under ncu on the GPU box): one multigrid hierarchy at n^2,
doubly periodic, a few V-cycles / F-cycles / smooths through the C ABI.
    python tools/prof_mg.py [n] [what]      what in {vcycle, fcycle, smooth, all}
"""
import ctypes
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from fluid2d_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
what = sys.argv[2] if len(sys.argv) > 2 else "all"
geom = sys.argv[3] if len(sys.argv) > 3 else "perio"
L = _lib.lib()
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
ptr = lambda t: ctypes.c_void_p(t.data_ptr())
cm = torch.ones((n+6, n+6), dtype=torch.float64, device="cuda")
cm[-1, :] = 0
cm[:, -1] = 0
if geom == "closed":
    cm[:3, :] = 0
    cm[-4:, :] = 0
    cm[:, :3] = 0
    cm[:, -4:] = 0
h = ctypes.c_void_p()
L.mg_create(ctypes.byref(h), ptr(cm), n+6, n+6, 1./n, 1./n, 8./9., 1., 0., s)
print("levels", L.mg_nlevels(h), "modes", [L.mg_level_matrix_mode(h, l) for l in range(L.mg_nlevels(h))])
if len(sys.argv) > 4:
    L.mg_set_graphs(h, int(sys.argv[4]))
x = torch.zeros((n+6, n+6), dtype=torch.float64, device="cuda")
b = torch.randn((n+6, n+6), dtype=torch.float64, device="cuda")
b -= b[3:-3, 3:-3].mean()
L.fill_halo(ptr(b), 3, n+6, n+6, s)


def timeit(f, reps=10):
    f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps


if os.environ.get("F2D_PROF_RANGE"):
    # under `ncu --profile-from-start off`: capture exactly one pass of the chosen cycle
    f = {"vcycle": lambda: L.mg_two_vcycle(h, ptr(x), ptr(b), s),
         "fcycle": lambda: L.mg_fcycle(h, 0, s),
         "smooth": lambda: L.mg_smooth(h, 0, ptr(x), ptr(b), 2, s)}[what]
    f()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    f()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    L.mg_destroy(h)
    sys.exit(0)
if what in ("smooth", "all"):
    ms = timeit(lambda: L.mg_smooth(h, 0, ptr(x), ptr(b), 2, s))/2
    print("smooth2 level0: %.4f ms  -> %.0f GB/s at 24 B/cell" % (ms, 24.*n*n/ms/1e6))
if what in ("vcycle", "all"):
    ms = timeit(lambda: L.mg_two_vcycle(h, ptr(x), ptr(b), s))/2
    print("V-cycle: %.4f ms  (139.3 B/cell -> %.0f GB/s)" % (ms, 139.3*n*n/ms/1e6))
if what in ("fcycle", "all"):
    nite, res = ctypes.c_int(), ctypes.c_double()
    x.zero_()
    t0 = time.time()
    L.mg_solve(h, ptr(x), ptr(b), 1e-11, 4, ctypes.byref(nite), ctypes.byref(res), s)
    torch.cuda.synchronize()
    ms = timeit(lambda: L.mg_fcycle(h, 0, s), reps=5)
    print("F-cycle: %.4f ms ; solve nite=%d res=%.2e" % (ms, nite.value, res.value))
L.mg_destroy(h)
