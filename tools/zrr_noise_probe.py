"""product-build (FMA) distance from the oracle hierarchy on a periodic white-noise problem,
with and without the fused descent kernel (F2D_MG_NO_ZRR), raw and with the mean (the null
space of the periodic operator) removed.   python tools/zrr_noise_probe.py [n]"""
import ctypes
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

if len(sys.argv) > 2 and sys.argv[2] == "child":
    import numpy as np
    import gpu_util as g
    import test_gpu_multigrid as T
    from fluid2d_b200 import _lib
    from oracle import kernels as K
    n = int(sys.argv[1])
    lib = _lib.lib(strict=False)
    ref, h, rng = T.make(lib, "perio", n, n)
    s = g.stream()
    shape = ref.msk[0].shape
    rhs = rng.standard_normal(shape)
    rhs[3:-3, 3:-3] -= rhs[3:-3, 3:-3].mean()
    K.fortran_multigrid.fillhalo(rhs, 3)
    psi0 = 0.01 * rng.standard_normal(shape)
    K.fortran_multigrid.fillhalo(psi0, 3)

    def err(a, b):
        i = (slice(3, -3), slice(3, -3))
        raw = np.linalg.norm(a[i]-b[i])/np.linalg.norm(b[i])
        a0, b0 = a[i]-a[i].mean(), b[i]-b[i].mean()
        return raw, np.linalg.norm(a0-b0)/np.linalg.norm(b0)
    pr = psi0.copy()
    d = g.dev(psi0)
    drhs = g.dev(rhs)
    for rep in range(2):
        ref.two_vcycle(pr, rhs)
        lib.mg_two_vcycle(h, g.ptr(d), g.ptr(drhs), s)
        print("  twoVcycle #%d  raw %.3e  mean removed %.3e" % ((rep,)+err(g.host(d), pr)))
    pr = psi0.copy()
    ref.solve(pr, rhs, maxite=3, tol=1e-11)
    d = g.dev(psi0)
    nite, res = ctypes.c_int(), ctypes.c_double()
    lib.mg_solve(h, g.ptr(d), g.ptr(drhs), 1e-11, 3, ctypes.byref(nite), ctypes.byref(res), s)
    print("  solve (3 F)    raw %.3e  mean removed %.3e" % err(g.host(d), pr))
else:
    n = sys.argv[1] if len(sys.argv) > 1 else "512"
    for nz in ("1", "0"):
        print("F2D_MG_NO_ZRR=%s, %s^2" % (nz, n))
        env = dict(os.environ, F2D_MG_NO_ZRR=nz)
        subprocess.call([sys.executable, os.path.abspath(__file__), n, "child"], env=env)
