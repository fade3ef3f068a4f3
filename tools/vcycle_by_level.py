"""Cost of the multigrid by starting level, through the CUDA graphs the solver replays: V(lev) =
one V-cycle started at level lev of an n^2 doubly periodic hierarchy (so V(lev) - V(lev+1) is what
the three kernels of level lev cost per visit), and the F-cycle.   python tools/vcycle_by_level.py [n]"""
import ctypes
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from fluid2d_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
L = _lib.lib()
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
ptr = lambda t: ctypes.c_void_p(t.data_ptr())
cm = torch.ones((n+6, n+6), dtype=torch.float64, device="cuda")
cm[-1, :] = 0
cm[:, -1] = 0
h = ctypes.c_void_p()
L.mg_create(ctypes.byref(h), ptr(cm), n+6, n+6, 1./n, 1./n, 8./9., 1., 0., s)
nlev = L.mg_nlevels(h)
b = torch.randn((n+6, n+6), dtype=torch.float64, device="cuda")
L.fill_halo(ptr(b), 3, n+6, n+6, s)
L.copy(L.mg_level_ptr(h, 0, 3), ptr(b), (n+6)*(n+6)*8, s)
L.mg_fcycle(h, 0, s)      # gives every level a right-hand side


def timeit(f, reps=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return 1e3*e0.elapsed_time(e1)/reps


rows = []
for lev in range(nlev):
    us = timeit(lambda: L.mg_vcycle(h, lev, s))
    rows.append((lev, us))
for k, (lev, us) in enumerate(rows):
    size = n >> lev
    own = us-rows[k+1][1] if k+1 < len(rows) else us
    print("V-cycle from level %2d (%5d^2): %8.1f us   this level alone: %7.1f us" % (lev, size, us, own))
print("F-cycle from level 0: %.1f us" % timeit(lambda: L.mg_fcycle(h, 0, s), reps=5))
L.mg_destroy(h)
