"""Multigrid parity: hierarchy set-up (masks, Galerkin matrices), every per-level
operator, the V/F cycles, twoVcycle and solve of the CUDA library against the oracle's
restated hierarchy (oracle/model.py: MG), which is pinned against the reference's own
gmg Python (tests/test_oracle_golden.py).

Jacobi smoothing is order independent, so the -fmad=false build must reproduce the
oracle's fields BIT FOR BIT through whole cycles; only the residual norms (reduction
order) may differ in the last digits."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import kernels as K  # noqa: E402
from oracle import model as om  # noqa: E402


def corner_mask(msk):
    w = np.zeros(msk.shape)
    K.fortran_operators.celltocorner(msk * 1., w)
    w[w < 1.] = 0.
    return w


def cell_mask(kind, ny, nx, rng):
    m = np.ones((ny + 6, nx + 6), dtype=np.int8)
    if kind in ("closed", "obstacle", "xchannel"):
        m[:3, :] = 0
        m[-3:, :] = 0
    if kind in ("closed", "obstacle"):
        m[:, :3] = 0
        m[:, -3:] = 0
    if kind == "obstacle":
        yy, xx = np.mgrid[0:ny + 6, 0:nx + 6]
        m[(yy - ny * 0.4) ** 2 + (xx - nx * 0.3) ** 2 < (0.12 * min(nx, ny)) ** 2] = 0
        m[ny // 2:ny // 2 + 2, nx // 2:] = 0
    return m


CASES = [("perio", 64, 64), ("closed", 64, 64), ("obstacle", 64, 128), ("xchannel", 32, 128),
         ("perio", 16, 8), ("obstacle", 128, 128), ("perio", 256, 128), ("closed", 256, 256)]


@pytest.fixture(params=["strict", "product"])
def L(request):
    from fluid2d_b200 import _lib
    return _lib.lib(strict=request.param == "strict"), request.param == "strict"


def make(lib, kind, ny, nx, Rd=0., force_stored=False, notail=False, noctail=False, ctail_nc=None, ctail_mincells=None,
         noptail=False, ptail_maxn=None):
    import os
    import gpu_util as g
    os.environ["F2D_MG_FORCE_STORED"] = "1" if force_stored else "0"
    os.environ["F2D_MG_NO_TAIL"] = "1" if notail else "0"
    os.environ["F2D_MG_NO_CTAIL"] = "1" if noctail else "0"
    # cluster tail: CTAs per cluster (default 16) and the size below which a level is replicated
    # (0: distribute every level with >= 4 rows per CTA, so that small test grids exercise the
    # distributed-shared-memory paths too)
    os.environ["F2D_MG_NO_PTAIL"] = "1" if noptail else "0"     # all-fluid periodic hierarchies: interior-only tail
    for key, val in (("F2D_CTAIL_NC", ctail_nc), ("F2D_CTAIL_MINCELLS", ctail_mincells), ("F2D_PTAIL_MAXN", ptail_maxn)):
        if val is None:
            os.environ.pop(key, None)
        else:
            os.environ[key] = str(val)
    rng = np.random.default_rng(ny + nx)
    msk = cell_mask(kind, ny, nx, rng)
    cm = corner_mask(msk)
    dx, dy = 1. / nx, 1. / nx
    ref = om.MG(cm, nx, ny, dx, dy, Rd=(Rd if Rd > 0 else None))
    h = ctypes.c_void_p()
    lib.mg_create(ctypes.byref(h), g.ptr(g.keep(cm)), ny + 6, nx + 6, dx, dy, 8. / 9., 1., Rd, g.stream())
    return ref, h, rng


def level_array(lib, h, lev, which, shape, dtype=np.float64):
    import torch
    import gpu_util as g
    ny, nx = shape
    n = ny * nx * (5 if which == 1 else 1)
    p = lib.mg_level_ptr(h, lev, which)
    buf = torch.empty(n, dtype=torch.int8 if dtype == np.int8 else torch.float64, device="cuda")
    nbytes = n * (1 if dtype == np.int8 else 8)
    lib.copy(g.ptr(buf), p, nbytes, g.stream())
    a = buf.cpu().numpy()
    return a.reshape((5, ny, nx)) if which == 1 else a.reshape((ny, nx))


@pytest.mark.parametrize("kind,ny,nx", CASES)
@pytest.mark.parametrize("Rd", [0., 0.1])
def test_hierarchy_setup(L, kind, ny, nx, Rd):
    lib, strict = L
    ref, h, rng = make(lib, kind, ny, nx, Rd)
    try:
        assert lib.mg_nlevels(h) == ref.nlevs
        for lev in range(ref.nlevs):
            sy, sx = ctypes.c_int(), ctypes.c_int()
            lib.mg_level_shape(h, lev, ctypes.byref(sy), ctypes.byref(sx))
            assert (sy.value, sx.value) == ref.msk[lev].shape
            shape = ref.msk[lev].shape
            np.testing.assert_array_equal(level_array(lib, h, lev, 0, shape, np.int8), ref.msk[lev])
            A = level_array(lib, h, lev, 1, shape)
            Aref = np.moveaxis(ref.A[lev], 2, 0)
            if strict:
                np.testing.assert_array_equal(A, Aref)
            else:
                np.testing.assert_allclose(A, Aref, rtol=1e-14, atol=1e-14 * np.abs(Aref).max())
    finally:
        lib.mg_destroy(h)


@pytest.mark.parametrize("kind,ny,nx", CASES)
def test_level_operators(L, kind, ny, nx):
    import gpu_util as g
    lib, strict = L
    ref, h, rng = make(lib, kind, ny, nx)
    s = g.stream()
    try:
        for lev in range(ref.nlevs):
            shape = ref.msk[lev].shape
            x = rng.standard_normal(shape) * ref.msk[lev]
            b = rng.standard_normal(shape) * ref.msk[lev]
            K.fortran_multigrid.fillhalo(x, 3)
            K.fortran_multigrid.fillhalo(b, 3)
            for nite in (1, 3):
                xr = x.copy(); ref.smooth(lev, xr, b, nite)
                d = g.dev(x); lib.mg_smooth(h, lev, g.ptr(d), g.ptr(g.keep(b)), nite, s)
                g.check(g.host(d), xr, strict, what="smooth lev %d" % lev)
            rr = np.zeros(shape); ref.residual(lev, x, b, rr)
            d = g.dev(np.ones(shape)); lib.mg_residual(h, lev, g.ptr(g.keep(x)), g.ptr(g.keep(b)), g.ptr(d), s)
            g.check(g.host(d), rr, strict, what="residual lev %d" % lev)
            import torch
            out = torch.zeros(1, dtype=torch.float64, device="cuda")
            lib.mg_sumsq(h, lev, g.ptr(g.keep(x)), g.ptr(out), s)
            np.testing.assert_allclose(np.sqrt(g.host(out)[0]), ref.norm(lev, x), rtol=1e-13)
            if lev < ref.nlevs - 1:
                cshape = ref.msk[lev + 1].shape
                xc = np.ones(cshape); ref.down(lev, x, xc)
                d = g.dev(np.ones(cshape)); lib.mg_restrict(h, lev, g.ptr(g.keep(x)), g.ptr(d), s)
                g.check(g.host(d), xc, strict, what="restrict lev %d" % lev)
                c = rng.standard_normal(cshape)
                xf = np.ones(shape); ref.up(lev, c, xf)
                d = g.dev(np.ones(shape)); lib.mg_interpolate(h, lev, g.ptr(g.keep(c)), g.ptr(d), 0, s)
                g.check(g.host(d), xf, strict, what="interpolate lev %d" % lev)
                d = g.dev(x); lib.mg_interpolate(h, lev, g.ptr(g.keep(c)), g.ptr(d), 1, s)
                g.check(g.host(d), x + xf, strict, what="interpolate+add lev %d" % lev)
    finally:
        lib.mg_destroy(h)


def test_matrix_classes(L):
    """doubly periodic: every level is a constant stencil; walls: the finest level is
    stencil x mask products, Galerkin levels next to walls keep stored coefficients"""
    lib, strict = L
    ref, h, rng = make(lib, "perio", 64, 64)
    assert [lib.mg_level_matrix_mode(h, l) for l in range(ref.nlevs)] == [1]*ref.nlevs
    lib.mg_destroy(h)
    ref, h, rng = make(lib, "obstacle", 64, 128)
    modes = [lib.mg_level_matrix_mode(h, l) for l in range(ref.nlevs)]
    assert modes[0] == 2 and set(modes[1:]) <= {0, 2}
    lib.mg_destroy(h)
    ref, h, rng = make(lib, "perio", 64, 64, force_stored=True)
    assert [lib.mg_level_matrix_mode(h, l) for l in range(ref.nlevs)] == [0]*ref.nlevs
    lib.mg_destroy(h)


@pytest.mark.parametrize("kind,ny,nx", CASES)
@pytest.mark.parametrize("graphs,force_stored,notail,noctail,ctail_nc,ctail_mincells,noptail,ptail_maxn",
                         [(0, False, False, False, None, None, False, None), (1, False, False, False, None, None, False, None),
                          (1, True, False, False, None, None, False, None), (1, False, True, False, None, None, False, None),
                          (1, False, False, True, None, None, False, None), (1, True, False, True, None, None, False, None),
                          (1, False, False, False, 16, 0, False, None), (1, True, False, False, 16, 0, False, None),
                          (1, False, False, False, 8, 0, False, None), (1, True, False, False, 8, 0, False, None),
                          (1, False, False, False, 4, 0, False, None), (1, False, False, False, 2, 0, False, None),
                          # the general cluster tail on the periodic cases too; the periodic tail from 256^2
                          (1, False, False, False, None, None, True, None), (1, False, False, False, 16, 0, True, None),
                          (1, False, False, False, None, None, False, 256), (0, False, False, False, 8, 0, False, 256),
                          (1, False, False, False, 1, None, False, None)])
def test_cycles_and_solve(L, kind, ny, nx, graphs, force_stored, notail, noctail, ctail_nc, ctail_mincells, noptail,
                          ptail_maxn):
    """graphs on/off; matrix class forced to 'stored'; coarse levels by the cluster tail
    kernel (<= 256^2, 16 / 8 / 4 / 2 CTAs with ghost rows exchanged through distributed shared
    memory; small levels distributed too when ctail_mincells = 0), by the one-CTA tail kernel
    (<= 64^2), or by the per-level kernels"""
    import gpu_util as g
    lib, strict = L
    ref, h, rng = make(lib, kind, ny, nx, force_stored=force_stored, notail=notail, noctail=noctail,
                       ctail_nc=ctail_nc, ctail_mincells=ctail_mincells, noptail=noptail, ptail_maxn=ptail_maxn)
    s = g.stream()
    lib.mg_set_graphs(h, graphs)
    tol_cycle = 1e-12
    try:
        shape = ref.msk[0].shape
        rhs = rng.standard_normal(shape) * ref.msk[0]
        if kind == "perio":
            rhs[3:-3, 3:-3] -= rhs[3:-3, 3:-3].mean()
        K.fortran_multigrid.fillhalo(rhs, 3)
        psi0 = 0.01 * rng.standard_normal(shape) * ref.msk[0]
        K.fortran_multigrid.fillhalo(psi0, 3)
        # twoVcycle, called twice (the second call replays the cached graph)
        pr = psi0.copy()
        d = g.dev(psi0)
        drhs = g.dev(rhs)
        for rep in range(2):
            ref.two_vcycle(pr, rhs)
            lib.mg_two_vcycle(h, g.ptr(d), g.ptr(drhs), s)
            g.check(g.host(d), pr, strict, tol=tol_cycle, what="twoVcycle #%d" % rep)
        # solve
        pr = psi0.copy()
        nite_ref, res_ref = ref.solve(pr, rhs, maxite=4, tol=1e-11)
        d = g.dev(psi0)
        nite, res = ctypes.c_int(), ctypes.c_double()
        lib.mg_solve(h, g.ptr(d), g.ptr(drhs), 1e-11, 4, ctypes.byref(nite), ctypes.byref(res), s)
        assert nite.value == nite_ref
        g.check_res(res.value, res_ref)
        g.check(g.host(d), pr, strict, tol=1e-11, what="solve")
        # zero right-hand side: returns (0, 0.) without touching psi (hierarchy.py:161-164)
        d = g.dev(psi0)
        lib.mg_solve(h, g.ptr(d), g.ptr(g.keep(np.zeros(shape))), 1e-11, 4, ctypes.byref(nite), ctypes.byref(res), s)
        assert nite.value == 0 and res.value == 0.
        np.testing.assert_array_equal(g.host(d), psi0)
    finally:
        lib.mg_destroy(h)


def test_convergence_factor(L):
    """a working multigrid contracts the residual by > 10x per F-cycle on a smooth RHS"""
    import gpu_util as g
    lib, strict = L
    ref, h, rng = make(lib, "closed", 128, 128)
    try:
        shape = ref.msk[0].shape
        yy, xx = np.mgrid[0:shape[0], 0:shape[1]]
        rhs = np.sin(xx * 0.1) * np.cos(yy * 0.07) * ref.msk[0]
        d = g.dev(np.zeros(shape))
        nite, res = ctypes.c_int(), ctypes.c_double()
        lib.mg_solve(h, g.ptr(d), g.ptr(g.keep(rhs)), 1e-11, 4, ctypes.byref(nite), ctypes.byref(res), g.stream())
        assert res.value < 1e-6
    finally:
        lib.mg_destroy(h)


@pytest.mark.parametrize("ny,nx", [(512, 512), (256, 1024), (1024, 256)])
@pytest.mark.parametrize("graphs", [0, 1])
def test_fused_descent_levels(L, ny, nx, graphs):
    """periodic hierarchies deep enough for the one-kernel level descent
    (fused::k_zsmooth_resid_restrict: smooth from x = 0 + residual + restriction, levels below the
    first one of a cycle that are larger than the shared-memory tail): twoVcycle and solve against
    the oracle hierarchy, bit for bit on the strict build"""
    import gpu_util as g
    lib, strict = L
    ref, h, rng = make(lib, "perio", ny, nx)
    s = g.stream()
    lib.mg_set_graphs(h, graphs)
    try:
        shape = ref.msk[0].shape
        rhs = rng.standard_normal(shape)
        rhs[3:-3, 3:-3] -= rhs[3:-3, 3:-3].mean()
        K.fortran_multigrid.fillhalo(rhs, 3)
        psi0 = 0.01 * rng.standard_normal(shape)
        K.fortran_multigrid.fillhalo(psi0, 3)
        def demean(a):
            # product build (FMA contraction): rounding differences accumulate in the constant, the
            # null space of the periodic operator, which no sweep damps (measured at 512^2,
            # tools/zrr_noise_probe.py: 3e-11 raw after three F-cycles, 1e-14 with the mean taken
            # out, the same with and without the fused kernel); the model removes that mean
            # (reference core/operators.py:474-478, "to avoid the drift"), so the product build is compared without it
            if strict:
                return a
            a = a.copy()
            a -= a[3:-3, 3:-3].mean()
            return a
        pr = psi0.copy()
        d = g.dev(psi0)
        drhs = g.dev(rhs)
        for rep in range(2):
            ref.two_vcycle(pr, rhs)
            lib.mg_two_vcycle(h, g.ptr(d), g.ptr(drhs), s)
            g.check(demean(g.host(d)), demean(pr), strict, tol=1e-12, what="twoVcycle #%d" % rep)
        pr = psi0.copy()
        nite_ref, res_ref = ref.solve(pr, rhs, maxite=3, tol=1e-11)
        d = g.dev(psi0)
        nite, res = ctypes.c_int(), ctypes.c_double()
        lib.mg_solve(h, g.ptr(d), g.ptr(drhs), 1e-11, 3, ctypes.byref(nite), ctypes.byref(res), s)
        assert nite.value == nite_ref
        g.check_res(res.value, res_ref)
        g.check(demean(g.host(d)), demean(pr), strict, tol=1e-12, what="solve")
    finally:
        lib.mg_destroy(h)
