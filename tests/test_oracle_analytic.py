"""Analytic known-answer tests of the CPU oracle's kernel arithmetic (oracle/f2d_oracle.c).

The reference ships no golden vectors and its Fortran cannot be compiled here (DESIGN.md
section 2), so the transcription of the kernels is checked against the mathematics the
Fortran implements -- weights, index offsets, signs and loop ranges:

  * the 1/3/5-point upwind and 2/4/6-point centred face interpolants are the finite-volume
    reconstructions (cell averages -> face value) that are exact for polynomials of the
    matching degree, and the tendency is the flux divergence
    (fortran_advection.f90:36-44, 149-155, 204-211);
  * flux splitting: minmax is plain upwinding, the parabola replaces |u| below
    aparab*umax and joins it continuously (fortran_advection.f90:77-86);
  * celltocorner / cornertocell average, computeorthogradient is (-d/dy, d/dx) of psi
    (fortran_operators.f90:2-64, 102-122);
  * the finest multigrid operator is a consistent 9-point Laplacian, the residual of an
    exact solution vanishes, one smoothing is the damped Jacobi formula, restriction is
    full weighting and interpolation is bilinear (fortran_multigrid.f90, level.py:261-302);
  * computekemaxu / computesumandnorm on uniform fields (fortran_diag.f90).
"""
import numpy as np
import pytest

from oracle import kernels as K
from oracle import model as om

NH = 3
fa, fo, fd, fm = K.fortran_advection, K.fortran_operators, K.fortran_diag, K.fortran_multigrid


def grid(ny, nx, dx):
    """cell-centre coordinates of an [ny, nx] array with halo 3 (x = (i - nh + 0.5) dx)"""
    x = (np.arange(nx)-NH+0.5)*dx
    y = (np.arange(ny)-NH+0.5)*dx
    return np.meshgrid(x, y)


@pytest.mark.parametrize("order", [1, 3, 5, 2, 4, 6])
@pytest.mark.parametrize("direction", ["x", "y"])
@pytest.mark.parametrize("sign", [1., -1.])
def test_face_interpolants_are_exact_for_polynomials(order, direction, sign):
    ny, nx, dx = 30, 34, 0.125
    xx, yy = grid(ny, nx, dx)
    s = xx if direction == "x" else yy
    # upwind order p uses p points: exact for degree p-1; centred order p: degree p-1 too
    deg = order-1
    coef = np.array([0.7, -1.3, 0.9, 0.4, -0.6, 0.35])[:deg+1]
    # finite-volume reconstruction: q holds the CELL AVERAGES of the polynomial, the face value
    # the interpolant returns is its point value at the face
    prim = lambda z: sum(c*z**(k+1)/(k+1) for k, c in enumerate(coef))
    q = (prim(s+0.5*dx)-prim(s-0.5*dx))/dx
    U = 0.8*sign
    u = np.full((ny, nx), U if direction == "x" else 0.)
    v = np.full((ny, nx), 0. if direction == "x" else U)
    msk = np.ones((ny, nx), dtype=np.int8)
    dq = np.zeros((ny, nx))
    cst = np.array([dx, dx, 0.05, abs(U), 0.])      # aparab = 0: pure upwinding
    adv = fa.adv_upwind if order % 2 else fa.adv_centered
    adv(msk, q, dq, u, v, cst, NH, 0, order)
    # exact flux divergence: -U (q(face+) - q(face-))/dx with the polynomial evaluated at the faces
    sp, sm = s+0.5*dx, s-0.5*dx
    exact = -U*(sum(c*sp**k for k, c in enumerate(coef))-sum(c*sm**k for k, c in enumerate(coef)))/dx
    inner = (slice(NH, -NH), slice(NH, -NH))
    # the float32 weights of the Fortran (default real(4) literals) limit the agreement to ~1e-7
    np.testing.assert_allclose(dq[inner], exact[inner], rtol=0, atol=3e-7*np.abs(q).max()/dx)
    # outside the interior the routine writes nothing
    assert not dq[:NH].any() and not dq[:, :NH].any() and not dq[-NH:].any() and not dq[:, -NH:].any()


def test_flux_splitting_minmax_and_parabolic():
    ny, nx, dx = 12, 40, 0.1
    rng = np.random.default_rng(0)
    q = rng.standard_normal((ny, nx))
    msk = np.ones((ny, nx), dtype=np.int8)
    v = np.zeros((ny, nx))
    umax, aparab = 2., 0.25
    u1 = aparab*umax
    # first order: face value is the upwind cell; flux = up*q_i + um*q_{i+1}
    for method in (0, 1):
        for U in (1.5, -1.5, 0.2, -0.3, 0.):
            u = np.full((ny, nx), U)
            dq = np.zeros((ny, nx))
            fa.adv_upwind(msk, q, dq, u, v, np.array([dx, dx, 0.05, umax, aparab]), NH, method, 1)
            def split(w):
                ww = abs(w)
                if method == 1 and ww < u1:
                    ww = w*w/(2*u1)+0.5*u1          # the parabola aa*u**2 + bb
                return 0.5*(w+ww), 0.5*(w-ww)
            up, um = split(U)
            vp, vm = split(0.)                       # v = 0 still diffuses under the parabola
            fx = up*q[:, :-1]+um*q[:, 1:]            # flux through the east face of cell i
            fy = vp*q[:-1, :]+vm*q[1:, :]            # flux through the north face of cell j
            exact = (-(fx[1:-1, 1:]-fx[1:-1, :-1])/dx-(fy[1:, 1:-1]-fy[:-1, 1:-1])/dx)
            np.testing.assert_allclose(dq[NH:-NH, NH:-NH], exact[NH-1:-(NH-1), NH-1:-(NH-1)], rtol=1e-13, atol=1e-13)
    # the parabola joins |u| continuously at u1
    assert abs((u1*u1/(2*u1)+0.5*u1)-u1) < 1e-15


def test_mask_lowers_the_order_and_blocks_the_flux():
    """a solid cell two cells downstream forces the 3-point interpolant; a face next to a
    solid cell carries no flux (fortran_advection.f90:63-70)"""
    ny, nx, dx = 12, 30, 0.1
    xx, _ = grid(ny, nx, dx)
    q = xx**2+dx*dx/12.            # cell averages of x**2: exact for the 3- and the 5-point reconstruction
    msk = np.ones((ny, nx), dtype=np.int8)
    msk[:, 15] = 0
    u = np.full((ny, nx), 1.)
    v = np.zeros((ny, nx))
    dq = np.zeros((ny, nx))
    fa.adv_upwind(msk, q, dq, u, v, np.array([dx, dx, 0.05, 1., 0.]), NH, 0, 5)
    exact = -((xx+0.5*dx)**2-(xx-0.5*dx)**2)/dx
    # cells whose two faces are at least one cell away from the wall: still exact (3rd order window)
    for i in (11, 12, 18, 19):
        np.testing.assert_allclose(dq[NH:-NH, i], exact[NH:-NH, i], atol=2e-6)
    # cell 14: its east face touches the solid cell -> no outflow, only the inflow through the west face
    west = (xx[0, 14]-0.5*dx)**2
    np.testing.assert_allclose(dq[NH:-NH, 14], +west/dx, atol=2e-6)


def test_celltocorner_cornertocell_orthogradient():
    ny, nx, dx, dy = 20, 26, 0.1, 0.2
    x = (np.arange(nx)-NH+0.5)*dx
    y = (np.arange(ny)-NH+0.5)*dy
    xx, yy = np.meshgrid(x, y)
    f = 2.*xx-3.*yy+1.
    c = np.full((ny, nx), np.nan)
    fo.celltocorner(f, c)
    # the corner (j,i) is the upper-right corner of cell (j,i): linear functions are averaged exactly
    np.testing.assert_allclose(c[:-1, :-1], 2.*(xx+0.5*dx)[:-1, :-1]-3.*(yy+0.5*dy)[:-1, :-1]+1., atol=1e-13)
    assert np.isnan(c[-1]).all() and np.isnan(c[:, -1]).all()
    back = np.full((ny, nx), np.nan)
    cc = 2.*(xx+0.5*dx)-3.*(yy+0.5*dy)+1.
    fo.cornertocell(cc, back)
    np.testing.assert_allclose(back[1:, 1:], f[1:, 1:], atol=1e-13)
    # psi at corners: u = -dpsi/dy on the east face, v = +dpsi/dx on the north face
    psi = 0.5*(xx+0.5*dx)-0.25*(yy+0.5*dy)
    msk = np.ones((ny, nx), dtype=np.int8)
    u, v = np.zeros((ny, nx)), np.zeros((ny, nx))
    fo.computeorthogradient(msk, psi, dx, dy, NH, u, v)
    np.testing.assert_allclose(u[1:-1, 1:-1], 0.25, atol=1e-12)
    np.testing.assert_allclose(v[1:-1, 1:-1], 0.5, atol=1e-12)
    msk[:, 10] = 0
    fo.computeorthogradient(msk, psi, dx, dy, NH, u, v)
    assert not u[1:-1, 9].any() and not u[1:-1, 10].any() and not v[1:-1, 10].any()


def periodic_mg(n, dx):
    cm = np.ones((n+2*NH, n+2*NH))
    cm[-1, :] = 0
    cm[:, -1] = 0
    return om.MG(cm, n, n, dx, dx)


def test_multigrid_operator_is_a_consistent_laplacian():
    n, dx = 32, 1./32
    mg = periodic_mg(n, dx)
    k = 2*np.pi
    x1 = (np.arange(n+2*NH)-NH+1.)*dx          # corner coordinates
    xx, yy = np.meshgrid(x1, x1)
    x = np.sin(k*xx)*np.cos(k*yy)               # periodic on the unit square
    b = np.zeros_like(x)
    r = np.zeros_like(x)
    mg.residual(0, x, b, r)                     # r = b - A x = -Laplacian(x)
    lap = -2*k*k*x
    inner = (slice(NH, -NH), slice(NH, -NH))
    err = np.abs(-r[inner]-lap[inner]).max()/np.abs(lap).max()
    assert err < 3*(k*dx)**2                    # second-order truncation error
    # finest stencil of level.py:261-302 for dx == dy: [[.25,.5,.25],[.5,-3,.5],[.25,.5,.25]]/(dx*dy)
    A = mg.A[0]
    np.testing.assert_allclose(A[10, 10, :5]*dx*dx, [0.25, 0.5, 0.25, 0.5, -3.])
    # residual of an exact solution is zero; a constant is in the null space
    one = np.ones_like(x)
    mg.residual(0, one, b, r)
    assert np.abs(r[inner]).max() < 1e-9
    # one application of smooth = two damped-Jacobi sweeps with omega = 8/9
    rng = np.random.default_rng(3)
    x = rng.standard_normal(x.shape)
    fm.fillhalo(x, NH)
    b = rng.standard_normal(x.shape)
    fm.fillhalo(b, NH)
    omega = 8./9.

    def jacobi(z):
        nb = (0.25*(np.roll(np.roll(z, 1, 0), 1, 1)+np.roll(np.roll(z, 1, 0), -1, 1)
                    + np.roll(np.roll(z, -1, 0), 1, 1)+np.roll(np.roll(z, -1, 0), -1, 1))
              + 0.5*(np.roll(z, 1, 0)+np.roll(z, -1, 0)+np.roll(z, 1, 1)+np.roll(z, -1, 1)))/(dx*dx)
        return (1-omega)*z+(omega/(3./(dx*dx)))*(nb-b)
    zi = x[inner]
    bi = b[inner]
    b_save, b = b, bi           # jacobi() reads b: the interior block, like z
    expect = jacobi(jacobi(zi))
    b = b_save
    y = x.copy()
    mg.smooth(0, y, b, 1)
    np.testing.assert_allclose(y[inner], expect, rtol=1e-11, atol=1e-11*np.abs(expect).max())


def test_restriction_and_interpolation_weights():
    n, dx = 32, 1./32
    mg = periodic_mg(n, dx)
    fshape, cshape = mg.msk[0].shape, mg.msk[1].shape
    # full weighting, centred on fine point 2*jc-2 (0-based; j1 = 2*j2-3 in the Fortran), preserves linear functions
    jf, if_ = np.meshgrid(np.arange(fshape[0]), np.arange(fshape[1]), indexing="ij")
    fine = 1.+0.5*jf-0.25*if_
    coarse = np.zeros(cshape)
    mg.down(0, fine, coarse)
    jc, ic = np.meshgrid(np.arange(cshape[0]), np.arange(cshape[1]), indexing="ij")
    expect = 1.+0.5*(2*jc-2)-0.25*(2*ic-2)
    inner = (slice(NH, -NH), slice(NH, -NH))
    np.testing.assert_allclose(coarse[inner][1:-1, 1:-1], expect[inner][1:-1, 1:-1], atol=1e-12)
    # a fine delta spreads with weights 1/4, 1/8, 1/16
    fine = np.zeros(fshape)
    fine[2*8-2, 2*9-2] = 1.
    mg.down(0, fine, coarse)
    assert coarse[8, 9] == 0.25
    fine[:] = 0.
    fine[2*8-1, 2*9-2] = 1.       # between coarse rows 8 and 9
    mg.down(0, fine, coarse)
    assert coarse[8, 9] == 0.125 and coarse[9, 9] == 0.125
    fine[:] = 0.
    fine[2*8-1, 2*9-1] = 1.
    mg.down(0, fine, coarse)
    assert coarse[8, 9] == 0.0625 and coarse[9, 10] == 0.0625 and coarse[8, 10] == 0.0625
    # bilinear interpolation reproduces linear functions of the coarse index
    c = 2.+0.5*jc-0.75*ic
    out = np.zeros(fshape)
    mg.up(0, c, out)
    expect = 2.+0.5*((jf+2)/2.)-0.75*((if_+2)/2.)
    np.testing.assert_allclose(out[inner], expect[inner], atol=1e-12)


def test_diagnostics_on_uniform_fields():
    ny, nx = 20, 24
    msk = np.ones((ny, nx), dtype=np.int8)
    u = np.full((ny, nx), 0.6)
    v = np.full((ny, nx), -0.8)
    ke, maxu = fd.computekemaxu(msk, u, v, NH)
    ncell = (ny-2*NH)*(nx-2*NH)
    np.testing.assert_allclose(ke, 0.5*(0.36+0.64)*ncell, rtol=1e-13)
    np.testing.assert_allclose(maxu, 1.4, rtol=1e-13)       # |u| + |v| of the cell-centred velocity
    w = np.full((ny, nx), 2.)
    z, z2 = fd.computesumandnorm(msk, w, NH)
    assert z == 2.*ncell and z2 == 4.*ncell
    msk[5, 5] = 0
    z, z2 = fd.computesumandnorm(msk, w, NH)
    assert z == 2.*(ncell-1)


def test_diffusion_is_the_masked_five_point_laplacian():
    """fortran_operators.f90:125-156: dtrac += K/dx^2 * sum of (neighbour - centre) over FLUID
    neighbours (homogeneous Neumann at walls), fluid cells of rows/cols 2..m-1 only"""
    ny, nx, dx, K0 = 18, 22, 0.25, 0.3
    xx, yy = grid(ny, nx, dx)
    t = xx**2+2.*yy**2                          # Laplacian = 2 + 4 = 6 exactly for the 5-point formula
    msk = np.ones((ny, nx), dtype=np.int8)
    d = np.full((ny, nx), 0.5)
    fo.add_diffusion(msk, t, dx, NH, K0, d)
    np.testing.assert_allclose(d[1:-1, 1:-1], 0.5+6.*K0, rtol=1e-12)
    assert (d[0] == 0.5).all() and (d[:, 0] == 0.5).all() and (d[-1] == 0.5).all() and (d[:, -1] == 0.5).all()
    # a wall on the east side removes that neighbour's contribution (no flux through the wall)
    msk[:, 12] = 0
    d[:] = 0.
    fo.add_diffusion(msk, t, dx, NH, K0, d)
    j, i = 8, 11
    east = t[j, i+1]-t[j, i]
    np.testing.assert_allclose(d[j, i], K0/dx**2*(6.*dx**2-east), rtol=1e-12)
    assert not d[:, 12].any()                   # solid cells get no tendency


def test_torque_is_the_centred_x_derivative_of_buoyancy():
    """fortran_operators.f90:330-381: domega += g/(2dx)*(b(i+1)-b(i-1)) on interior cells whose
    two x-neighbour pairs are all fluid"""
    ny, nx, dx, g = 14, 20, 0.2, 9.81
    xx, yy = grid(ny, nx, dx)
    b = 3.*xx-yy+0.5*xx*yy                      # db/dx = 3 + 0.5 y
    msk = np.ones((ny, nx), dtype=np.int8)
    dw = np.zeros((ny, nx))
    fo.add_torque(msk, b, dx, NH, g, dw)
    np.testing.assert_allclose(dw[NH:-NH, NH:-NH], (g*(3.+0.5*yy))[NH:-NH, NH:-NH], rtol=1e-12)
    assert not dw[:NH].any() and not dw[:, :NH].any()
    msk[:, 9] = 0
    dw[:] = 0.
    fo.add_torque(msk, b, dx, NH, g, dw)
    assert not dw[NH:-NH, 8:11].any()           # the wall column and both cells next to it
    np.testing.assert_allclose(dw[NH:-NH, 11], (g*(3.+0.5*yy))[NH:-NH, 11], rtol=1e-12)


def test_noslip_source_is_the_tangential_velocity_along_walls():
    """fortran_operators.f90:221-277: psi at corners; at a wall face the cell-averaged
    tangential velocity / (grid step) enters the FLUID cell next to the wall, with the sign
    of the vorticity that would cancel it"""
    ny, nx, dx, dy = 16, 18, 0.1, 0.1
    x = (np.arange(nx)-NH+1.)*dx                # corner coordinates
    y = (np.arange(ny)-NH+1.)*dy
    xx, yy = np.meshgrid(x, y)
    msk = np.ones((ny, nx), dtype=np.int8)
    msk[:NH, :] = 0                             # a wall below row NH (south wall)
    # uniform flow u = U along x: psi = -U*y  (u = -dpsi/dy)
    U = 0.7
    psi = -U*yy
    src = np.full((ny, nx), np.nan)
    fo.computenoslipsourceterm(msk, psi, src, dx, dy, NH)
    # first fluid row j = NH: the south face is a wall; u along it gives a vortex sheet of
    # strength u/dy, spread on the cell:  y(j,i) += u with u = -(psi(j,i)+psi(j,i-1)-psi(j-2,i)-psi(j-2,i-1))/(2dxdy)
    expect = -(psi[NH, 5]+psi[NH, 4]-psi[NH-2, 5]-psi[NH-2, 4])/(2*dx*dy)
    np.testing.assert_allclose(src[NH, NH:-NH], expect, rtol=1e-12)
    np.testing.assert_allclose(expect, 2.*U/dx, rtol=1e-12)   # two corner pairs, two rows apart: 4 U dy / (2 dx dy)
    assert not src[NH+1:-NH, NH:-NH].any()      # no wall, no source
    # the gather form the CUDA kernel uses gives the same field (scatter in the Fortran)
    tot = 0.
    for i in range(NH, nx-NH+1):
        tot += src[NH, i]
    assert tot != 0.


def test_remaining_reductions():
    ny, nx = 16, 20
    rng = np.random.default_rng(2)
    msk = (rng.random((ny, nx)) > 0.2).astype(np.int8)
    a, b = rng.standard_normal((ny, nx)), rng.standard_normal((ny, nx))
    inner = (slice(NH, -NH), slice(NH, -NH))
    m = msk[inner] != 0
    np.testing.assert_allclose(fd.computedotprod(msk, a, b, NH), (a[inner]*b[inner])[m].sum(), rtol=1e-13)
    np.testing.assert_allclose(fd.computesum(msk, a, NH), a[inner][m].sum(), rtol=1e-13)
    z, z2 = fd.computesumandnorm(msk, a, NH)
    np.testing.assert_allclose([z, z2], [a[inner][m].sum(), (a[inner][m]**2).sum()], rtol=1e-13)
    # multigrid norm: sum of squares over the interior where the mask is fluid
    np.testing.assert_allclose(fm.computenorm(msk, a, NH), (a[inner][m]**2).sum(), rtol=1e-13)
    # ke: quarter sum of the four face contributions of each cell, maxu from cell-centred speeds
    u, v = a, b
    ke, maxu = fd.computekemaxu(np.ones((ny, nx), dtype=np.int8), u, v, NH)
    ui, um = u[inner], u[NH:-NH, NH-1:-NH-1]
    vi, vm = v[inner], v[NH-1:-NH-1, NH:-NH]
    np.testing.assert_allclose(ke, 0.25*(ui**2+um**2+vi**2+vm**2).sum(), rtol=1e-13)
    np.testing.assert_allclose(maxu, 0.5*(np.abs(ui+um)+np.abs(vi+vm)).max(), rtol=1e-13)
