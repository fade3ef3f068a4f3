"""The kernels written without GPU access (line relaxation, thermal-wind operators), compiled
for the host from their own source text and run with one thread (tests/kernel_hostsim.py),
against the oracle / the reference's numpy expressions: bit-exact.  Functional check of the
kernel bodies only; device behaviour is the business of tests/test_gpu_zz_late.py."""
import ctypes

import numpy as np
import pytest

import kernel_hostsim
import emu_device
from oracle import model as om
from test_gpu_multigrid import cell_mask, corner_mask

c_dp = ctypes.POINTER(ctypes.c_double)
c_bp = ctypes.POINTER(ctypes.c_int8)


def dp(a):
    return a.ctypes.data_as(c_dp)


def bp(a):
    return a.ctypes.data_as(c_bp)


@pytest.fixture(scope="module")
def sim():
    return kernel_hostsim.build()


@pytest.mark.parametrize("kind,ny,nx,dx,dy", [("xchannel", 32, 16, 1./16, 1./16), ("closed", 64, 32, 1./32, 1./32),
                                              ("closed", 32, 128, 1./128, 1./128), ("xchannel", 32, 64, 1./8, 1./64)])
def test_line_relaxation_kernel_body(sim, kind, ny, nx, dx, dy):
    rng = np.random.default_rng(ny+nx)
    cm = corner_mask(cell_mask(kind, ny, nx, rng))
    ref = om.MG(cm, nx, ny, dx, dy, relaxation='tridiagonal')
    for lev in range(ref.nlevs):
        shape = ref.msk[lev].shape
        x = rng.standard_normal(shape)*ref.msk[lev]
        b = rng.standard_normal(shape)*ref.msk[lev]
        planes = np.ascontiguousarray(np.moveaxis(ref.A[lev], 2, 0))    # the library's layout: 5 planes
        xs, x0 = x.copy(), x.copy()
        sim.sim_smooth_tridiag(bp(ref.msk[lev]), dp(planes), dp(xs), dp(b), shape[0], shape[1])
        om.fm.smoothtridiag(ref.msk[lev], ref.A[lev], x, b)
        np.testing.assert_array_equal(xs, x, err_msg="level %d" % lev)
        assert not np.array_equal(xs, x0)


@pytest.mark.parametrize("ny,nx", [(38, 70), (22, 22)])
def test_thermalwind_kernel_bodies(sim, ny, nx):
    emu = emu_device.EmuLib()
    rng = np.random.default_rng(ny*nx)
    msk = np.ones((ny, nx), dtype=np.int8)
    msk[:3, :] = 0
    msk[-3:, :] = 0
    msk[:, :3] = 0
    msk[:, -3:] = 0
    msk[ny//2, nx//3] = 0
    P = lambda a: a.ctypes.data   # noqa: E731
    dx, dy, grav, f0 = 1./37, 1./19, 9.81, 0.137
    D = ctypes.c_double
    for axis in (0, 1):
        a = rng.standard_normal((ny, nx))
        s_ = a.copy()
        sim.sim_extrapolate_bry(dp(s_), 3, ny, nx, axis)
        emu.extrapolate_bry(P(a), 3, ny, nx, axis, None)
        np.testing.assert_array_equal(s_, a)
    b, V, u = (rng.standard_normal((ny, nx)) for _ in range(3))
    y = rng.standard_normal((ny, nx))
    ys = y.copy()
    sim.sim_tw_torque(bp(msk), dp(b), dp(V), D(dx), D(dy), D(grav), D(f0), dp(ys), ny, nx)
    emu.tw_torque(P(msk), P(b), P(V), dx, dy, grav, f0, P(y), ny, nx, None)
    np.testing.assert_array_equal(ys[1:-1, 1:-1], y[1:-1, 1:-1])
    y = rng.standard_normal((ny, nx))
    ys = y.copy()
    sim.sim_tw_coriolis(bp(msk), dp(u), D(f0), dp(ys), ny, nx)
    emu.tw_coriolis(P(msk), P(u), f0, P(y), ny, nx, None)
    np.testing.assert_array_equal(ys[:, 1:], y[:, 1:])
    out = rng.standard_normal((ny, nx))
    outs = out.copy()
    sim.sim_jacobian(bp(msk), dp(b), dp(V), D(dx), D(dy), dp(outs), ny, nx)
    emu.jacobian(P(msk), P(b), P(V), dx, dy, P(out), ny, nx, None)
    np.testing.assert_array_equal(outs, out)
