"""GPU parity of the cases written after the last GPU session of round 1 (cases.LATE): the
BoussinesqTS model (experiments/doublediffusion) with the damped-Jacobi and with the line
(tridiagonal) relaxation of the multigrid, and the QG model with its bottom-torque /
ageostrophic diagnostics.  Same check and same tolerances as tests/test_gpu_golden.py
(fixtures from the reference's own Python); the file name makes it run last.

The line relaxation is also compared operator by operator with the oracle: Grid.smooth on
every level of a masked hierarchy, then whole V- and F-cycles and the full solve.
"""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import cases  # noqa: E402
from test_gpu_golden import test_case_matches_reference_run as check_case  # noqa: E402


# Order: what runs on kernels already green on a B200 first, the kernels that have never run on
# one last, so that a fault in a new kernel cannot take the earlier tests with it.
NEW_KERNELS = {"si_32_flx": "thermal-wind kernels", "dbldiff_32_tridiag": "line relaxation"}


@pytest.mark.parametrize("name", sorted(cases.LATE-set(NEW_KERNELS), reverse=True))
def test_late_case_matches_reference_run(name):
    check_case(name)


from test_gpu_multigrid import cell_mask, corner_mask, level_array  # noqa: E402


@pytest.fixture(params=["strict", "product"])
def L(request):
    from fluid2d_b200 import _lib
    return _lib.lib(strict=request.param == "strict"), request.param == "strict"


@pytest.mark.parametrize("ny,nx", [(38, 70), (22, 22), (70, 134)])
def test_thermalwind_kernels_against_numpy(L, ny, nx):
    """f2d_extrapolate_bry / f2d_tw_torque / f2d_tw_coriolis / f2d_jacobian / f2d_negative_part
    against the numpy expressions of the reference (operators.py:330-394, thermalwind.py:86-98,
    136-137) as tests/emu_device.py restates them: bit-exact on both builds (the kernels use
    round-to-nearest multiplies and adds, no FMA, like numpy)."""
    import gpu_util as g
    import emu_device
    lib, strict = L
    emu = emu_device.EmuLib()
    rng = np.random.default_rng(ny*nx)
    s = g.stream()
    msk = np.ones((ny, nx), dtype=np.int8)
    msk[:3, :] = 0
    msk[-3:, :] = 0
    msk[:, :3] = 0
    msk[:, -3:] = 0
    msk[ny//2, nx//3] = 0
    d_msk = g.keep(msk)
    P = lambda a: a.ctypes.data   # noqa: E731
    dx, dy, grav, f0 = 1./37, 1./19, 9.81, 0.137
    for axis in (0, 1):
        a = rng.standard_normal((ny, nx))
        d = g.keep(a)
        lib.extrapolate_bry(g.ptr(d), 3, ny, nx, axis, s)
        emu.extrapolate_bry(P(a), 3, ny, nx, axis, None)
        np.testing.assert_array_equal(g.host(d), a, err_msg="extrapolate axis %d" % axis)
    b, V, u = (rng.standard_normal((ny, nx)) for _ in range(3))
    y = rng.standard_normal((ny, nx))
    d_y = g.keep(y)
    lib.tw_torque(g.ptr(d_msk), g.ptr(g.keep(b)), g.ptr(g.keep(V)), dx, dy, grav, f0, g.ptr(d_y), ny, nx, s)
    emu.tw_torque(P(msk), P(b), P(V), dx, dy, grav, f0, P(y), ny, nx, None)
    np.testing.assert_array_equal(g.host(d_y)[1:-1, 1:-1], y[1:-1, 1:-1], err_msg="tw_torque")
    y = rng.standard_normal((ny, nx))
    d_y = g.keep(y)
    lib.tw_coriolis(g.ptr(d_msk), g.ptr(g.keep(u)), f0, g.ptr(d_y), ny, nx, s)
    emu.tw_coriolis(P(msk), P(u), f0, P(y), ny, nx, None)
    np.testing.assert_array_equal(g.host(d_y)[:, 1:], y[:, 1:], err_msg="tw_coriolis")
    out = rng.standard_normal((ny, nx))
    d_out = g.keep(out)
    lib.jacobian(g.ptr(d_msk), g.ptr(g.keep(b)), g.ptr(g.keep(V)), dx, dy, g.ptr(d_out), ny, nx, s)
    emu.jacobian(P(msk), P(b), P(V), dx, dy, P(out), ny, nx, None)
    np.testing.assert_array_equal(g.host(d_out), out, err_msg="jacobian")
    neg = np.empty(ny*nx)
    d_neg = g.keep(neg)
    lib.negative_part(g.ptr(d_neg), g.ptr(g.keep(b)), ny*nx, s)
    emu.negative_part(P(neg), P(b), ny*nx, None)
    np.testing.assert_array_equal(g.host(d_neg), neg, err_msg="negative_part")


def test_thermalwind_case_matches_reference_run():
    check_case("si_32_flx")


@pytest.mark.parametrize("kind,ny,nx,dx,dy", [("xchannel", 32, 16, 1./16, 1./16),
                                              ("closed", 64, 32, 1./32, 1./32),
                                              ("closed", 32, 128, 1./128, 1./128),
                                              ("xchannel", 128, 64, 1./64, 1./64),
                                              ("xchannel", 32, 64, 1./8, 1./64)])
def test_line_relaxation_against_oracle(L, kind, ny, nx, dx, dy):
    """f2d_mg_set_relaxation(1): Grid.smooth per level, V-cycle, F-cycle, solve and twoVcycle
    against oracle/model.py:MG with relaxation='tridiagonal' (smoothtridiag,
    fortran_multigrid.f90:215-317).  The last case has flat cells (dy/dx = 1/8 <= 0.2): the
    5-point operator, and f2d_mg_create selects the line relaxation by itself.
    Bit-exact on the -fmad=false build; 1e-11 of the field's maximum on the product build."""
    import gpu_util as g
    from oracle import model as om
    lib, strict = L
    rng = np.random.default_rng(ny+nx)
    cm = corner_mask(cell_mask(kind, ny, nx, rng))
    ref = om.MG(cm, nx, ny, dx, dy, relaxation='tridiagonal')
    assert ref.relaxation == 'tridiagonal'
    s = g.stream()
    h = ctypes.c_void_p()
    lib.mg_create(ctypes.byref(h), g.ptr(g.keep(cm)), ny+6, nx+6, dx, dy, 8./9., 1., 0., s)
    lib.mg_set_relaxation(h, 1)

    def close(a, b, what):
        if strict:
            np.testing.assert_array_equal(a, b, err_msg=what)
        else:
            scale = max(np.abs(b).max(), 1e-300)
            assert np.abs(a-b).max() <= 1e-11*scale, "%s: %.3e" % (what, np.abs(a-b).max()/scale)

    try:
        assert lib.mg_nlevels(h) == ref.nlevs
        for lev in range(ref.nlevs):
            shape = ref.msk[lev].shape
            x = rng.standard_normal(shape)*ref.msk[lev]
            b = rng.standard_normal(shape)*ref.msk[lev]
            dxd, dbd = g.keep(x), g.keep(b)
            lib.mg_smooth(h, lev, g.ptr(dxd), g.ptr(dbd), 1, s)
            ref.smooth(lev, x, b, 1)
            close(g.host(dxd), x, "smooth level %d" % lev)
        shape = ref.msk[0].shape
        nbytes = shape[0]*shape[1]*8
        b0 = rng.standard_normal(shape)*ref.msk[0]
        om.fm.fillhalo(b0, 3)
        px = ctypes.c_void_p(lib.mg_level_ptr(h, 0, 2))
        pb = ctypes.c_void_p(lib.mg_level_ptr(h, 0, 3))
        for cyc in ("vcycle", "fcycle"):
            ref.x[0][:] = 0.
            ref.b[0][:] = b0
            getattr(ref, cyc)(0)
            lib.zero(px, nbytes, s)
            lib.copy(pb, g.ptr(g.keep(b0)), nbytes, s)
            getattr(lib, "mg_"+cyc)(h, 0, s)
            close(level_array(lib, h, 0, 2, shape), ref.x[0], cyc)
        psi = np.zeros(shape)
        nite_ref, res_ref = ref.solve(psi, b0.copy(), maxite=4, tol=1e-11)
        d_psi, d_b0 = g.keep(np.zeros(shape)), g.keep(b0)
        nite, res = ctypes.c_int(), ctypes.c_double()
        lib.mg_solve(h, g.ptr(d_psi), g.ptr(d_b0), 1e-11, 4, ctypes.byref(nite), ctypes.byref(res), s)
        assert nite.value == nite_ref
        g.check_res(res.value, res_ref)
        close(g.host(d_psi), psi, "solve")
        ref.two_vcycle(psi, b0.copy())
        lib.mg_two_vcycle(h, g.ptr(d_psi), g.ptr(d_b0), s)
        close(g.host(d_psi), psi, "twoVcycle")
    finally:
        lib.mg_destroy(h)


def test_line_relaxation_case_matches_reference_run():
    check_case("dbldiff_32_tridiag")
