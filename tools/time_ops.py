"""Isolated timing of the non-multigrid kernels of the Euler step at n^2 (CUDA events around 20
back-to-back launches through the C ABI): advection, orthogradient, celltocorner, the RK
combinations, the fused diagnostics.   python tools/time_ops.py [n]"""
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
shape = (n+6, n+6)
q, dq, u, v, psi, w = [torch.randn(shape, dtype=torch.float64, device="cuda") for _ in range(6)]
cst = (ctypes.c_double*5)(1./n, 1./n, 0.05, 3., 0.05)
peak = 6538.3


def timeit(name, f, bytes_per_cell, reps=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    us = 1e3*e0.elapsed_time(e1)/reps
    gbs = bytes_per_cell*n*n/(us*1e-6)/1e9
    print("%-34s %8.1f us  %6.0f GB/s  %.2f of peak" % (name, us, gbs, gbs/peak))


timeit("adv_upwind order 5, mask-free", lambda: L.adv_upwind(None, ptr(q), ptr(dq), ptr(u), ptr(v), None, None, cst, 3, 1, 5,
                                                            n+6, n+6, 1, s), 32.)
timeit("adv_upwind order 3, mask-free", lambda: L.adv_upwind(None, ptr(q), ptr(dq), ptr(u), ptr(v), None, None, cst, 3, 1, 3,
                                                            n+6, n+6, 1, s), 32.)
msk = torch.ones(shape, dtype=torch.int8, device="cuda")
timeit("adv_upwind order 5, masked", lambda: L.adv_upwind(ptr(msk), ptr(q), ptr(dq), ptr(u), ptr(v), None, None, cst, 3, 1, 5,
                                                         n+6, n+6, 1, s), 33.)
timeit("mask_orthogradient, mask-free", lambda: L.mask_orthogradient(None, None, ptr(psi), 1./n, 1./n, 3, ptr(u), ptr(v),
                                                                    n+6, n+6, s), 32.)
timeit("celltocorner", lambda: L.celltocorner(ptr(w), ptr(psi), n+6, n+6, s), 16.)
N = (n+6)*(n+6)
timeit("ts_xpay (2 in, 1 out)", lambda: L.ts_xpay(ptr(dq), ptr(q), 0.1, ptr(u), N, s), 24.)
timeit("ts_rk3ssp_final (4 in, 1 out)", lambda: L.ts_rk3ssp_final(ptr(q), 0.1, ptr(u), ptr(v), ptr(w), N, s), 40.)
