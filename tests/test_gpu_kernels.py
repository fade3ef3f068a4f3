"""Per-kernel parity: every C-ABI entry point against the CPU oracle (oracle/kernels.py)
on the same seeded inputs.

Each test runs twice: on libf2d_b200_strict.so (same sources, -fmad=false), which must
agree with the oracle BIT FOR BIT (index / mask / loop-range handling is exact), and on
the product libf2d_b200.so (FMA contraction allowed), which must agree to a relative L2
of 1e-13 per kernel (the contract is 1e-12 on a full step).  Reductions use a different
summation order than the Fortran loop: tolerance 1e-13 relative in both builds.
"""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import kernels as K  # noqa: E402


def _gpu():
    import gpu_util
    return gpu_util


SHAPES = [(10, 10), (22, 38), (38, 22), (70, 70), (40, 262), (150, 134)]


def rand_mask(rng, ny, nx, kind):
    if kind == "ones":
        return np.ones((ny, nx), dtype=np.int8)
    m = np.ones((ny, nx), dtype=np.int8)
    if kind in ("closed", "blobs"):
        m[:3, :] = 0
        m[-3:, :] = 0
        m[:, :3] = 0
        m[:, -3:] = 0
    if kind in ("random", "blobs"):
        m[rng.random((ny, nx)) < (0.25 if kind == "random" else 0.08)] = 0
    return m


@pytest.fixture(params=["strict", "product"])
def L(request):
    g = _gpu()
    from fluid2d_b200 import _lib
    return _lib.lib(strict=request.param == "strict"), request.param == "strict"


@pytest.mark.parametrize("shape", SHAPES)
def test_fill_halo(L, shape):
    g = _gpu()
    lib, strict = L
    rng = np.random.default_rng(1)
    x = rng.standard_normal(shape)
    ref = x.copy()
    K.fortran_multigrid.fillhalo(ref, 3)
    d = g.dev(x)
    lib.fill_halo(g.ptr(d), 3, shape[0], shape[1], g.stream())
    np.testing.assert_array_equal(g.host(d), ref)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("mkind", ["ones", "none", "closed", "random", "blobs"])
@pytest.mark.parametrize("order,method", [(5, 1), (5, 0), (3, 1), (1, 0), (4, 0), (6, 0), (2, 0)])
def test_advection(L, shape, mkind, order, method):
    g = _gpu()
    lib, strict = L
    ny, nx = shape
    rng = np.random.default_rng(ny * 1000 + nx + order)
    msk = rand_mask(rng, ny, nx, "ones" if mkind == "none" else mkind)
    q = rng.standard_normal(shape)
    u = rng.standard_normal(shape) * 0.3
    v = rng.standard_normal(shape) * 0.3
    y0 = rng.standard_normal(shape)
    cst = np.array([1. / 64, 1. / 48, 0.05, 0.9, 0.05])
    upw = order % 2 == 1
    for flx in (False, True):
        ref = y0.copy()
        fxr, fyr = y0.copy() * 2, y0.copy() * 3
        if flx:
            f = K.fortran_fluxes.adv_upwind if upw else K.fortran_fluxes.adv_centered
            f(msk, q, ref, u, v, fxr, fyr, cst, 3, method, order)
        else:
            f = K.fortran_advection.adv_upwind if upw else K.fortran_advection.adv_centered
            f(msk, q, ref, u, v, cst, 3, method, order)
        for fill in (0, 1):
            refh = ref.copy()
            if fill:
                K.fortran_multigrid.fillhalo(refh, 3)
            dq = g.dev(y0)
            dfx, dfy = (g.dev(y0 * 2), g.dev(y0 * 3)) if flx else (None, None)
            dm = None if mkind == "none" else g.dev(msk)
            fn = lib.adv_upwind if upw else lib.adv_centered
            cst_c = (ctypes.c_double * 5)(*cst)
            fn(g.ptr(dm), g.ptr(g.keep(q)), g.ptr(dq), g.ptr(g.keep(u)), g.ptr(g.keep(v)),
               g.ptr(dfx), g.ptr(dfy), cst_c, 3, method, order, ny, nx, fill, g.stream())
            g.check(g.host(dq), refh, strict, what="dq")
            if flx:
                g.check(g.host(dfx), fxr, strict, what="xflx")
                g.check(g.host(dfy), fyr, strict, what="yflx")


@pytest.mark.parametrize("shape", [(38, 70), (70, 134), (150, 134), (134, 262)])
@pytest.mark.parametrize("mkind", ["none", "random"])
@pytest.mark.parametrize("ntr", [1, 2, 5])
def test_advection_multi_and_fused_stage(L, shape, mkind, ntr):
    """f2d_adv_multi: the tracers of a model in one launch (operators.py:214-236), with and without
    the fused Runge-Kutta stage output xout = xbase + coef*dq (timescheme.py:172-176): against the
    oracle's adv_upwind + fillhalo per tracer and numpy's x + c*dx, halo included"""
    g = _gpu()
    lib, strict = L
    ny, nx = shape
    rng = np.random.default_rng(ny + nx + ntr)
    msk = rand_mask(rng, ny, nx, "ones" if mkind == "none" else mkind)
    u = rng.standard_normal(shape) * 0.3
    v = rng.standard_normal(shape) * 0.3
    cst = np.array([1. / 64, 1. / 48, 0.05, 0.9, 0.05])
    cst_c = (ctypes.c_double * 5)(*cst)
    qs = [rng.standard_normal(shape) for _ in range(ntr)]
    xbs = [rng.standard_normal(shape) for _ in range(ntr)]   # (its halo need not hold periodic images)
    refs = []
    for q in qs:
        ref = np.zeros(shape)
        K.fortran_advection.adv_upwind(msk, q, ref, u, v, cst, 3, 1, 5)
        K.fortran_multigrid.fillhalo(ref, 3)
        refs.append(ref)
    coef = 0.37
    dm = None if mkind == "none" else g.dev(msk)
    du, dv = g.keep(u), g.keep(v)
    for fused in (False, True):
        dqs = [g.dev(np.zeros(shape)) for _ in range(ntr)]
        xos = [g.dev(np.full(shape, 7.)) for _ in range(ntr)]
        dq_in = [g.keep(q) for q in qs]
        dxb = [g.keep(x) for x in xbs]
        arr = lambda ts: (ctypes.c_void_p * ntr)(*[t.data_ptr() for t in ts])
        lib.adv_multi(g.ptr(dm), arr(dq_in), arr(dqs), ntr, g.ptr(du), g.ptr(dv), cst_c, 3, 1, 1, 5,
                      arr(dxb) if fused else None, arr(xos) if fused else None, coef, ny, nx, 1, g.stream())
        for k in range(ntr):
            g.check(g.host(dqs[k]), refs[k], strict, what="dq[%d]" % k)
            if fused:
                # numpy: x + coef*dx (product rounded, then the sum); both builds must give exactly
                # that from the dq they computed
                got_dq = g.host(dqs[k])
                np.testing.assert_array_equal(g.host(xos[k]), xbs[k] + coef*got_dq)


def test_advection_umax_zero_and_bad_nh(L):
    g = _gpu()
    lib, strict = L
    from fluid2d_b200._lib import F2DError
    ny, nx = 38, 38
    rng = np.random.default_rng(5)
    msk = rand_mask(rng, ny, nx, "closed")
    q, u, v = (rng.standard_normal((ny, nx)) for _ in range(3))
    cst = np.array([0.1, 0.1, 0.05, 0.0, 0.05])     # umax = 0 before the first diagnostics
    ref = np.zeros((ny, nx))
    K.fortran_advection.adv_upwind(msk, q, ref, u, v, cst, 3, 1, 5)
    dq = g.dev(np.zeros((ny, nx)))
    cst_c = (ctypes.c_double * 5)(*cst)
    args = [g.ptr(g.keep(msk)), g.ptr(g.keep(q)), g.ptr(dq), g.ptr(g.keep(u)), g.ptr(g.keep(v)), None, None, cst_c]
    lib.adv_upwind(*args, 3, 1, 5, ny, nx, 0, g.stream())
    out = g.host(dq)
    assert np.isfinite(out).all()
    g.check(out, ref, strict)
    with pytest.raises(F2DError) as e:
        lib.adv_upwind(*args, 2, 1, 5, ny, nx, 0, g.stream())
    assert e.value.code == lib.ERR_NH
    with pytest.raises(F2DError):
        lib.adv_upwind(*args, 3, 1, 4, ny, nx, 0, g.stream())


@pytest.mark.parametrize("shape", SHAPES + [(23, 37), (40, 518), (35, 600)])
def test_mask_orthogradient(L, shape):
    """psi *= mskp; computeorthogradient (operators.py:481,493) in one pass: masked form,
    and the mask-free form (msk = mskp = NULL) of all-fluid domains"""
    g = _gpu()
    lib, strict = L
    ny, nx = shape
    rng = np.random.default_rng(11 + ny + nx)
    fo = K.fortran_operators
    s = g.stream()
    for mkind in ("ones", "closed", "random"):
        msk = rand_mask(rng, ny, nx, mkind)
        mskp = np.zeros(shape, dtype=np.int8)
        mskp[:-1, :-1] = msk[:-1, :-1] & msk[:-1, 1:] & msk[1:, :-1] & msk[1:, 1:]
        psi = rng.standard_normal(shape)
        u0, v0 = rng.standard_normal(shape), rng.standard_normal(shape)
        pr = psi*mskp
        ur, vr = u0.copy(), v0.copy()
        fo.computeorthogradient(msk, pr, 0.01, 0.02, 3, ur, vr)
        forms = [(g.ptr(g.keep(msk)), g.ptr(g.keep(mskp)))]
        if mkind == "ones" and nx % 2 == 0:
            forms.append((None, None))
        for pm, pmp in forms:
            dp, du, dv = g.dev(psi), g.dev(u0), g.dev(v0)
            lib.mask_orthogradient(pm, pmp, g.ptr(dp), 0.01, 0.02, 3, g.ptr(du), g.ptr(dv), ny, nx, s)
            np.testing.assert_array_equal(g.host(dp), pr)
            g.check(g.host(du), ur, strict, what="u")
            g.check(g.host(dv), vr, strict, what="v")


@pytest.mark.parametrize("shape", SHAPES + [(35, 601), (70, 262)])
def test_mask_orthogradient_with_stage_update(L, shape):
    """f2d_mask_orthogradient_stage: the orthogradient and, from the velocities it has just derived,
    the Runge-Kutta stage state of u, v over the WHOLE arrays (timescheme.py:172-180) -- one kernel
    on even nx, the orthogradient followed by f2d_ts_xpay(2) otherwise: either way exactly
    numpy's  ub + c*du  and  ub + c*(ue + du)  of the du, dv the call itself leaves"""
    g = _gpu()
    lib, strict = L
    ny, nx = shape
    rng = np.random.default_rng(17 + ny + nx)
    s = g.stream()
    c = 0.3712
    for mkind in ("ones", "random"):
        msk = rand_mask(rng, ny, nx, mkind)
        mskp = np.zeros(shape, dtype=np.int8)
        mskp[:-1, :-1] = msk[:-1, :-1] & msk[:-1, 1:] & msk[1:, :-1] & msk[1:, 1:]
        psi = rng.standard_normal(shape)
        u0, v0 = rng.standard_normal(shape), rng.standard_normal(shape)   # the ring keeps these
        ub, vb, ue, ve = (rng.standard_normal(shape) for _ in range(4))
        forms = [(g.ptr(g.keep(msk)), g.ptr(g.keep(mskp)))]
        if mkind == "ones" and nx % 2 == 0:
            forms.append((None, None))
        for pm, pmp in forms:
            # plain call: the reference result of this build for psi, u, v
            dp, du, dv = g.dev(psi), g.dev(u0), g.dev(v0)
            lib.mask_orthogradient(pm, pmp, g.ptr(dp), 0.01, 0.02, 3, g.ptr(du), g.ptr(dv), ny, nx, s)
            pr, ur, vr = g.host(dp), g.host(du), g.host(dv)
            for extra in (False, True):
                dp, du, dv = g.dev(psi), g.dev(u0), g.dev(v0)
                uo, vo = g.dev(np.full(shape, 7.)), g.dev(np.full(shape, 7.))
                lib.mask_orthogradient_stage(pm, pmp, g.ptr(dp), 0.01, 0.02, 3, g.ptr(du), g.ptr(dv),
                                             g.ptr(g.keep(ub)), g.ptr(g.keep(vb)),
                                             g.ptr(g.keep(ue)) if extra else None, g.ptr(g.keep(ve)) if extra else None,
                                             g.ptr(uo), g.ptr(vo), c, ny, nx, s)
                np.testing.assert_array_equal(g.host(dp), pr)
                np.testing.assert_array_equal(g.host(du), ur)
                np.testing.assert_array_equal(g.host(dv), vr)
                np.testing.assert_array_equal(g.host(uo), ub + c*((ue + ur) if extra else ur))
                np.testing.assert_array_equal(g.host(vo), vb + c*((ve + vr) if extra else vr))


@pytest.mark.parametrize("shape", SHAPES + [(23, 37), (35, 601)])
def test_stencil_operators(L, shape):
    g = _gpu()
    lib, strict = L
    ny, nx = shape
    rng = np.random.default_rng(7 + ny + nx)
    fo = K.fortran_operators
    s = g.stream()
    for mkind in ("ones", "closed", "random"):
        msk = rand_mask(rng, ny, nx, mkind)
        a = rng.standard_normal(shape)
        b = rng.standard_normal(shape)
        # celltocorner / cornertocell
        ref = b.copy(); fo.celltocorner(a, ref)
        d = g.dev(b); lib.celltocorner(g.ptr(g.keep(a)), g.ptr(d), ny, nx, s)
        g.check(g.host(d), ref, strict, what="celltocorner")
        ref = b.copy(); fo.cornertocell(a, ref)
        d = g.dev(b); lib.cornertocell(g.ptr(g.keep(a)), g.ptr(d), ny, nx, s)
        g.check(g.host(d), ref, strict, what="cornertocell")
        # orthogradient
        ur, vr = a.copy(), b.copy()
        psi = rng.standard_normal(shape)
        fo.computeorthogradient(msk, psi, 0.01, 0.02, 3, ur, vr)
        du, dv = g.dev(a), g.dev(b)
        lib.orthogradient(g.ptr(g.keep(msk)), g.ptr(g.keep(psi)), 0.01, 0.02, 3, g.ptr(du), g.ptr(dv), ny, nx, s)
        g.check(g.host(du), ur, strict, what="u")
        g.check(g.host(dv), vr, strict, what="v")
        # diffusion
        for fill in (0, 1):
            ref = b.copy(); fo.add_diffusion(msk, a, 0.01, 3, 3e-4, ref)
            if fill:
                K.fortran_multigrid.fillhalo(ref, 3)
            d = g.dev(b)
            lib.add_diffusion(g.ptr(g.keep(msk)), g.ptr(g.keep(a)), 0.01, 3, 3e-4, g.ptr(d), ny, nx, fill, s)
            g.check(g.host(d), ref, strict, what="diffusion")
        # torque (with the y *= msk of operators.py:311)
        for premask in (0, 1):
            ref = b.copy()
            if premask:
                ref *= msk
            fo.add_torque(msk, a, 0.01, 3, 9.81, ref)
            K.fortran_multigrid.fillhalo(ref, 3)
            d = g.dev(b)
            lib.add_torque(g.ptr(g.keep(msk)), g.ptr(g.keep(a)), 0.01, 3, 9.81, g.ptr(d), ny, nx, premask, 1, s)
            g.check(g.host(d), ref, strict, what="torque")
        # no-slip source (scatter in the Fortran, gather here)
        ref = b.copy(); fo.computenoslipsourceterm(msk, psi, ref, 0.01, 0.02, 3)
        d = g.dev(b)
        lib.noslip_source(g.ptr(g.keep(msk)), g.ptr(g.keep(psi)), g.ptr(d), 0.01, 0.02, 3, ny, nx, s)
        g.check(g.host(d), ref, strict, what="noslip")


@pytest.mark.parametrize("shape", SHAPES + [(518, 1030)])
def test_reductions(L, shape):
    g = _gpu()
    lib, strict = L
    import torch
    ny, nx = shape
    rng = np.random.default_rng(11 + ny)
    fd = K.fortran_diag
    s = g.stream()
    sc = g.scratch(lib)
    out = torch.zeros(8, dtype=torch.float64, device="cuda")
    tol = 1e-13
    for mkind in ("ones", "random"):
        msk = rand_mask(rng, ny, nx, mkind)
        x, y, u, v, psi, src = (rng.standard_normal(shape) for _ in range(6))
        dm, dx_, dy_, du, dv, dpsi, dsrc = (g.dev(t) for t in (msk, x, y, u, v, psi, src))

        def close(a, b):
            assert abs(a - b) <= tol * max(1., abs(b)) * np.sqrt(ny * nx), (a, b)

        lib.computedotprod(g.ptr(dm), g.ptr(dx_), g.ptr(dy_), 3, ny, nx, g.ptr(out), g.ptr(sc), s)
        close(g.host(out)[0], fd.computedotprod(msk, x, y, 3))
        lib.computemax(g.ptr(dm), g.ptr(dx_), 3, ny, nx, g.ptr(out), g.ptr(sc), s)
        assert g.host(out)[0] == fd.computemax(msk, x, 3)
        lib.computesum(g.ptr(dm), g.ptr(dx_), 3, ny, nx, g.ptr(out), g.ptr(sc), s)
        close(g.host(out)[0], fd.computesum(msk, x, 3))
        lib.computesumandnorm(g.ptr(dm), g.ptr(dx_), 3, ny, nx, g.ptr(out), g.ptr(sc), s)
        r = fd.computesumandnorm(msk, x, 3)
        close(g.host(out)[0], r[0]); close(g.host(out)[1], r[1])
        lib.computenormmaxu(g.ptr(dm), g.ptr(dx_), 3, ny, nx, g.ptr(out), g.ptr(sc), s)
        r = fd.computenormmaxu(msk, x, 3)
        close(g.host(out)[0], r[0]); assert g.host(out)[1] == r[1]
        lib.computekemaxu(g.ptr(dm), g.ptr(du), g.ptr(dv), 3, ny, nx, g.ptr(out), g.ptr(sc), s)
        r = fd.computekemaxu(msk, u, v, 3)
        close(g.host(out)[0], r[0]); assert g.host(out)[1] == r[1]
        lib.computekemaxuv(g.ptr(dm), g.ptr(du), g.ptr(dv), 3, ny, nx, g.ptr(out), g.ptr(sc), s)
        r = fd.computekemaxuv(msk, u, v, 3)
        close(g.host(out)[0], r[0]); close(g.host(out)[1], r[1]); close(g.host(out)[2], r[2])
        lib.computekewithpsi(g.ptr(dm), g.ptr(dx_), g.ptr(dpsi), 3, ny, nx, g.ptr(out), g.ptr(sc), s)
        close(g.host(out)[0], fd.computekewithpsi(msk, x, psi, 3))
        lib.computenorm(g.ptr(dm), g.ptr(dx_), 3, ny, nx, g.ptr(out), g.ptr(sc), s)
        close(g.host(out)[0], K.fortran_multigrid.computenorm(msk, x, 3))
        lib.domain_sum(g.ptr(dx_), 3, ny, nx, g.ptr(out), g.ptr(sc), s)
        close(g.host(out)[0], np.sum(x[3:-3, 3:-3]))
        # fused Euler diagnostics
        xr, yr = np.meshgrid(np.arange(nx) * 0.01, np.arange(ny) * 0.02)
        lib.diag_euler(g.ptr(dm), g.ptr(du), g.ptr(dv), g.ptr(dx_), g.ptr(dpsi), g.ptr(dsrc),
                       g.ptr(g.keep(xr)), g.ptr(g.keep(yr)), 3, ny, nx, g.ptr(out), g.ptr(sc), s)
        o = g.host(out)
        ke, maxu = fd.computekemaxu(msk, u, v, 3)
        z, z2 = fd.computesumandnorm(msk, x, 3)
        assert o[0] == maxu
        for a, b in zip(o[1:], [ke, z, z2, fd.computedotprod(msk, x, xr, 3), fd.computedotprod(msk, x, yr, 3),
                                fd.computesum(msk, psi, 3), fd.computedotprod(msk, x, src, 3)]):
            close(a, b)


def test_timescheme_combinations_bitexact(L):
    """the numpy expressions of timescheme.py, evaluated without FMA: exact in BOTH builds"""
    g = _gpu()
    lib, strict = L
    rng = np.random.default_rng(3)
    n = 5 * 38 * 46
    x, a, b, d = (rng.standard_normal(n) for _ in range(4))
    dt = 0.0371
    s = g.stream()

    def run(fn, first, *rest):
        t = g.dev(first)
        fn(g.ptr(t), *[g.ptr(g.keep(r)) if isinstance(r, np.ndarray) else r for r in rest], n, s)
        return g.host(t)

    np.testing.assert_array_equal(run(lib.ts_axpy, x, dt, a), x + dt * a)
    np.testing.assert_array_equal(run(lib.ts_xpay, x * 0, x, dt, a), x + dt * a)
    np.testing.assert_array_equal(run(lib.ts_xpay2, x * 0, x, 0.25 * dt, a, b), x + (0.25 * dt) * (a + b))
    np.testing.assert_array_equal(run(lib.ts_rk3ssp_final, x, dt / 6., a, b, d), x + (dt / 6.) * (a + b + 4 * d))
    np.testing.assert_array_equal(run(lib.ts_ab2, x, 1.6 * dt, a, 0.6 * dt, b), x + ((1.6 * dt) * a - (0.6 * dt) * b))
    np.testing.assert_array_equal(run(lib.ts_ab3, x, 23 * dt / 12., a, 16 * dt / 12., b, 5 * dt / 12., d),
                                  x + ((23 * dt / 12.) * a - (16 * dt / 12.) * b + (5 * dt / 12.) * d))
    np.testing.assert_array_equal(run(lib.ts_set_xpay, x * 0, b, 2 * dt, a), b + (2 * dt) * a)
    np.testing.assert_array_equal(run(lib.ts_asselin, x, 0.1, a, b), x + 0.1 * (a + b - 2 * x))
    np.testing.assert_array_equal(run(lib.ts_am3, x, a, b), (1. / 12.) * (5. * x + 8. * a - b))
    m = (rng.random(n) < 0.5).astype(np.int8)
    np.testing.assert_array_equal(run(lib.mul_field, x, a), x * a)
    np.testing.assert_array_equal(run(lib.mul_mask, x, m), x * m)
    np.testing.assert_array_equal(run(lib.scale, x, 1. / dt), x * (1. / dt))
    np.testing.assert_array_equal(run(lib.add_scaled, x, -1., a), x - a)
    np.testing.assert_array_equal(run(lib.add_scaled_mask, x, dt, m), x + dt * m)
    np.testing.assert_array_equal(run(lib.set_sum, x, a, -0.3, b), a + (-0.3) * b)
    sc = np.array([1.2345])
    np.testing.assert_array_equal(run(lib.sub_devscalar, x, sc, 7.), x - sc[0] / 7.)
    np.testing.assert_array_equal(run(lib.sub_devscalar_mask, x, sc, 7., m), x - (sc[0] / 7.) * m)
