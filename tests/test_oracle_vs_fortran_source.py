"""oracle/f2d_oracle.c (the C restatement every parity test leans on) against the reference's
Fortran SOURCE executed statement by statement (oracle/fortran_source.py: no Fortran compiler
exists in this image, so the five .f90 files are translated mechanically to Python with the
language's typing rules modelled -- single-precision literals, conversion on assignment,
integer division, 1-based arrays).  Bit for bit, on small arrays with random masks, for every
routine the hot path uses.  Build container only: skipped where /root/reference is absent.
"""
import numpy as np
import pytest

from oracle import fortran_source as F
from oracle import kernels as K

pytestmark = pytest.mark.skipif(not F.available(), reason="/root/reference is not present on this machine")

fa, ff, fo, fd, fm = (K.fortran_advection, K.fortran_fluxes, K.fortran_operators, K.fortran_diag,
                      K.fortran_multigrid)
SHAPES = [(14, 18), (17, 13)]


def fields(rng, shape, kind):
    ny, nx = shape
    msk = np.ones(shape, dtype=np.int8)
    if kind == "closed":
        msk[:3, :] = 0
        msk[-3:, :] = 0
        msk[:, :3] = 0
        msk[:, -3:] = 0
    elif kind == "random":
        msk = (rng.random(shape) > 0.25).astype(np.int8)
    elif kind == "blob":
        msk[:3, :] = 0
        msk[-3:, :] = 0
        msk[ny//2-1:ny//2+2, nx//2-2:nx//2+1] = 0
    return msk


def eq(a, b, what):
    np.testing.assert_array_equal(np.asarray(a), np.asarray(b), err_msg=what)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("kind", ["ones", "closed", "random", "blob"])
@pytest.mark.parametrize("order,method", [(5, 1), (3, 0), (1, 1), (5, 0), (2, 0), (4, 0), (6, 0)])
def test_advection_and_fluxes(shape, kind, order, method):
    rng = np.random.default_rng(order*10+method+shape[0])
    msk = fields(rng, shape, kind)
    x, u, v = (rng.standard_normal(shape) for _ in range(3))
    cst = np.array([0.1, 0.07, 0.05, float(np.abs(u).max()), 0.05 if order % 2 else 0.])
    name = "adv_upwind" if order % 2 else "adv_centered"
    y0 = rng.standard_normal(shape)
    yc, yf = y0.copy(), y0.copy()
    getattr(fa, name)(msk, x, yc, u, v, cst, 3, method, order)
    F.call("fortran_advection", name, msk=msk, x=x, y=yf, u=u, v=v, cst=cst, nh=3, method=method, order=order)
    eq(yc, yf, name)
    yc, yf = y0.copy(), y0.copy()
    xc, xf, zc, zf = (np.zeros(shape) for _ in range(4))
    getattr(ff, name)(msk, x, yc, u, v, xc, zc, cst, 3, method, order)
    F.call("fortran_fluxes", name, msk=msk, x=x, y=yf, u=u, v=v, xflx=xf, yflx=zf, cst=cst, nh=3, method=method,
           order=order)
    eq(yc, yf, name+" (fluxes file)")
    eq(xc, xf, "xflx")
    eq(zc, zf, "yflx")


def test_advection_with_zero_umax_and_bad_halo():
    rng = np.random.default_rng(3)
    shape = (14, 14)
    msk = fields(rng, shape, "closed")
    x, u, v = (rng.standard_normal(shape) for _ in range(3))
    cst = np.array([0.1, 0.1, 0.05, 0., 0.05])
    yc, yf = np.zeros(shape), np.zeros(shape)
    fa.adv_upwind(msk, x, yc, u, v, cst, 3, 1, 5)
    F.call("fortran_advection", "adv_upwind", msk=msk, x=x, y=yf, u=u, v=v, cst=cst, nh=3, method=1, order=5)
    eq(yc, yf, "umax = 0")
    with pytest.raises(F.FortranStop):
        F.call("fortran_advection", "adv_upwind", msk=msk, x=x, y=yf, u=u, v=v, cst=cst, nh=2, method=1, order=5)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("kind", ["ones", "closed", "random", "blob"])
def test_operators(shape, kind):
    rng = np.random.default_rng(shape[1])
    msk = fields(rng, shape, kind)
    a, b = rng.standard_normal(shape), rng.standard_normal(shape)
    dx, dy = 0.13, 0.09
    uc, vc, uf, vf = (rng.standard_normal(shape) for _ in range(4))
    uf[:], vf[:] = uc, vc
    fo.computeorthogradient(msk, a, dx, dy, 3, uc, vc)
    F.call("fortran_operators", "computeorthogradient", msk=msk, psi=a, dx=dx, dy=dy, nh=3, u=uf, v=vf)
    eq(uc, uf, "orthogradient u")
    eq(vc, vf, "orthogradient v")
    for name, args in (("celltocorner", ("xr", "xp")), ("cornertocell", ("xp", "xr"))):
        oc, of = b.copy(), b.copy()
        getattr(fo, name)(a, oc)
        F.call("fortran_operators", name, **{args[0]: a, args[1]: of})
        eq(oc, of, name)
    oc, of = b.copy(), b.copy()
    fo.add_diffusion(msk, a, dx, 3, 0.37, oc)
    F.call("fortran_operators", "add_diffusion", msk=msk, trac=a, dx=dx, nh=3, kdiff=0.37, dtrac=of)
    eq(oc, of, "add_diffusion")
    oc, of = b.copy(), b.copy()
    fo.add_torque(msk, a, dx, 3, 9.81, oc)
    F.call("fortran_operators", "add_torque", msk=msk, buoy=a, dx=dx, nh=3, gravity=9.81, domega=of)
    eq(oc, of, "add_torque")
    oc, of = b.copy(), b.copy()
    tc = fo.computenoslipsourceterm(msk, a, oc, dx, dy, 3)
    out = F.call("fortran_operators", "computenoslipsourceterm", msk=msk, x=a, y=of, dx=dx, dy=dy, nh=3)
    eq(oc, of, "noslip source")
    assert tc == out["total"]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("kind", ["ones", "closed", "random"])
def test_diagnostics(shape, kind):
    rng = np.random.default_rng(shape[0]+7)
    msk = fields(rng, shape, kind)
    a, b = rng.standard_normal(shape), rng.standard_normal(shape)
    assert fd.computedotprod(msk, a, b, 3) == F.call("fortran_diag", "computedotprod", msk=msk, x=a, y=b, nh=3)["z"]
    assert fd.computemax(msk, a, 3) == F.call("fortran_diag", "computemax", msk=msk, x=a, nh=3)["y"]
    assert fd.computesum(msk, a, 3) == F.call("fortran_diag", "computesum", msk=msk, x=a, nh=3)["y"]
    r = F.call("fortran_diag", "computesumandnorm", msk=msk, x=a, nh=3)
    assert tuple(fd.computesumandnorm(msk, a, 3)) == (r["y"], r["y2"])
    r = F.call("fortran_diag", "computenormmaxu", msk=msk, x=a, nh=3)
    assert tuple(fd.computenormmaxu(msk, a, 3)) == (r["y"], r["ymax"])
    r = F.call("fortran_diag", "computekemaxu", msk=msk, u=a, v=b, nh=3)
    assert tuple(fd.computekemaxu(msk, a, b, 3)) == (r["ke"], r["maxu"])
    r = F.call("fortran_diag", "computekemaxuv", msk=msk, u=a, v=b, nh=3)
    assert tuple(fd.computekemaxuv(msk, a, b, 3)) == (r["ke"], r["maxu"], r["maxv"])
    assert fd.computekewithpsi(msk, a, b, 3) == F.call("fortran_diag", "computekewithpsi", msk=msk, omega=a, psi=b,
                                                       nh=3)["ke"]
    assert fm.computenorm(msk, a, 3) == F.call("fortran_multigrid", "computenorm", msk=msk, x=a, nh=3)["y"]
    assert fm.computeinner(msk, a, b, 3) == F.call("fortran_multigrid", "computeinner", msk=msk, x=a, y=b, nh=3)["z"]


@pytest.mark.parametrize("shape", [(14, 18), (22, 14)])
@pytest.mark.parametrize("kind", ["ones", "closed", "random"])
def test_multigrid_kernels(shape, kind):
    rng = np.random.default_rng(shape[0]*3+1)
    ny, nx = shape
    msk = fields(rng, shape, kind)
    A = rng.standard_normal(shape+(5,))
    A[:, :, 4] = -3.-rng.random(shape)
    x, b = rng.standard_normal(shape), rng.standard_normal(shape)
    xc, xf = x.copy(), x.copy()
    fm.smoothtwicewitha(msk, A, xc, b, 8./9.)
    F.call("fortran_multigrid", "smoothtwicewitha", msk=msk, a=A, x=xf, b=b, coef=8./9., yo=np.zeros((3, nx)))
    eq(xc[2:-2, 2:-2], xf[2:-2, 2:-2], "smoothtwicewithA (valid range 3..m-2)")
    xc, xf = x.copy(), x.copy()
    fm.smoothtridiag(msk, A, xc, b)
    F.call("fortran_multigrid", "smoothtridiag", msk=msk, a=A, x=xf, b=b)
    eq(xc, xf, "smoothtridiag")
    rc, rf = rng.standard_normal(shape), None
    rf = rc.copy()
    fm.computeresidualwitha(msk, A, x, b, rc)
    F.call("fortran_multigrid", "computeresidualwitha", msk=msk, a=A, x=x, b=b, y=rf)
    eq(rc, rf, "residual")
    hc, hf = x.copy(), x.copy()
    fm.fillhalo(hc, 3)
    F.call("fortran_multigrid", "fillhalo", x=hf, nh=3)
    eq(hc, hf, "fillhalo")
    # transfer operators between this grid (fine) and the next coarser one
    m2, n2 = (ny-6)//2+6, (nx-6)//2+6
    msk2 = fields(rng, (m2, n2), kind)
    x2c = rng.standard_normal((m2, n2))
    x2f = x2c.copy()
    fm.restrict(msk2, x, 3, x2c)
    F.call("fortran_multigrid", "restrict", msk2=msk2, x1=x, nh=3, x2=x2f)
    eq(x2c, x2f, "restrict")
    x1c = rng.standard_normal(shape)
    x1f = x1c.copy()
    coarse = rng.standard_normal((m2, n2))
    fm.interpolate(msk, msk2, coarse, 3, x1c)
    F.call("fortran_multigrid", "interpolate", msk1=msk, msk2=msk2, x2=coarse, nh=3, x1=x1f)
    eq(x1c, x1f, "interpolate")
    A9 = rng.standard_normal(shape+(9,))
    Ac = fm.coarsenmatrix(A9, msk, msk2, 3)
    Af = np.zeros((m2, n2, 9))
    F.call("fortran_multigrid", "coarsenmatrix", afine=A9, acoarse=Af, msk1=msk, msk2=msk2, nh=3)
    eq(Ac, Af, "coarsenmatrix")
