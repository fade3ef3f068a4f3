"""CPU ORACLE, part 2: restated orchestration (test infrastructure, NOT the product).

oracle/kernels.py restates the reference's Fortran; this file restates, in numpy, the
Python that drives it, so that a whole time step can be reproduced where the
reference itself is not available (the GPU box has no /root/reference):

  Param / Grid             core/param.py, core/defaults.json, core/grid.py:12-140
  Island                   core/island.py:6-43
  MG (Gridinfo, set-up,    core/gmg/level.py:24-117,227-335,340-496
      V/F cycles, solve)   core/gmg/hierarchy.py:23-218          (single rank)
  Ops                      core/operators.py:12-186,214-314,421-498
  Stepper                  core/timescheme.py:78-201
  EulerModel/BoussinesqModel  core/euler.py:19-223, core/boussinesq.py:17-153
  Fluid2d (set_dt, ...)    core/fluid2d.py:20-145,351-397
  Fluxes                   core/fluxes.py:6-213

It is PINNED: tests/test_oracle_golden.py requires that, for the cases of
tests/golden/cases.py it supports, it reproduces bit for bit the fixtures that
tests/golden/make_golden.py produced by running the reference's own Python on the
same oracle kernels.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this module.  Not covered (the goldens of those cases are pinned by
the reference run alone): the QG model, LFAM3 history (`xb`), dye/age tracers,
customized steps, multi-rank decomposition.
"""
import numpy as np

from . import kernels as K

fa = K.fortran_advection
fo = K.fortran_operators
fd = K.fortran_diag
fm = K.fortran_multigrid
ff = K.fortran_fluxes

NH = 3

DEFAULTS = dict(
    modelname='advection', expname='myexp', timestepping='RK3_SSP', order=5, aparab=0.05,
    flux_splitting_method='parabolic', relaxation='default', nh=3, adaptable_dt=True,
    dt=0.1, cfl=0.5, dtmax=5.0, rescaledtime='none', ninterrestart=1, nx=128, ny=128,
    Lx=1.0, Ly=1.0, geometry='disc', isisland=False, mpi=0, myrank=0, npx=1, npy=1,
    plot_interactive=True, var_to_save='vorticity', list_diag='all', nprint=20,
    freq_his=1.0, diag_fluxes=False, exacthistime=True, freq_diag=1.0, hydroepsilon=1.0,
    diffusion=False, customized=False, Kdiff=0.0, noslip=False, forcing=False,
    forcing_module='embedded', decay=True, enforce_momentum=False, spongelayer=False,
    datadir='~/data/fluid2d', beta=0., Rd=1., gravity=1.)


class Param(object):
    def __init__(self, defaultfile=None):
        for k, v in DEFAULTS.items():
            setattr(self, k, v)


class Island(object):
    """island.py: boundary psi values and the RHS correction they imply"""

    def __init__(self, grid):
        self.grid = grid
        self.items = []
        self.rhsp = np.zeros((grid.nyl, grid.nxl))
        self.psi = np.zeros((grid.nyl, grid.nxl))

    def add(self, idx, psi0):
        self.items.append((idx, psi0))

    def finalize(self):
        g = self.grid
        corner = np.zeros((g.nyl, g.nxl))
        cells = np.zeros((g.nyl, g.nxl))
        for idx, psi0 in self.items:
            cells[:, :] = 1.
            cells[idx] = 0.
            fo.celltocorner(cells, corner)
            inside = np.zeros((g.nyl, g.nxl), dtype=np.int8)
            inside[corner == 1] = 1
            inside = 1-inside
            nb = (np.roll(inside, -1, axis=1)+np.roll(inside, -1, axis=0)
                  + np.roll(inside, +1, axis=1)+np.roll(inside, +1, axis=0))
            z = nb*psi0/(g.dx*g.dy)
            self.rhsp[nb > 0] = z[nb > 0]
            self.psi[inside == 1] = psi0


class Grid(object):
    def __init__(self, param):
        param.myrank = 0
        param.nx = int(param.nx)
        param.ny = int(param.ny)
        self.nx, self.ny, self.nh = param.nx, param.ny, param.nh
        self.Lx, self.Ly, self.geometry = param.Lx, param.Ly, param.geometry
        self.npx = self.npy = 1
        self.i0 = self.j0 = 0
        nh = self.nh
        self.nxl = self.nx+2*nh
        self.nyl = self.ny+2*nh
        self.dx = self.Lx/self.nx
        self.dy = self.Ly/self.ny
        self.x1d = (np.arange(self.nxl)+0.5-nh)*self.dx
        self.y1d = (np.arange(self.nyl)+0.5-nh)*self.dy
        self.xr, self.yr = np.meshgrid(self.x1d, self.y1d)
        msk = np.ones((self.nyl, self.nxl), dtype=np.int8)
        if self.geometry in ('xperio', 'xchannel', 'closed', 'disc'):
            msk[-nh:, :] = 0
            msk[:nh, :] = 0
        if self.geometry in ('yperio', 'ychannel', 'closed', 'disc'):
            msk[:, -nh:] = 0
            msk[:, :nh] = 0
        if self.geometry == 'disc':
            r = np.sqrt((self.xr/self.Lx-0.5)**2 + (self.yr/self.Ly-0.5)**2)
            msk[r >= 0.5] = 0
        self.msk = msk
        self.msknoslip = msk.copy()
        self.finalize_msk()
        if param.isisland:
            self.island = Island(self)

    def domain_integration(self, z2d):
        nh = self.nh
        return np.array([np.sum(z2d[nh:-nh, nh:-nh])*1.])

    def finalize_msk(self):
        msk = self.msk
        self.area = self.domain_integration(msk)
        x0 = self.domain_integration(self.xr*msk) / self.area
        y0 = self.domain_integration(self.yr*msk) / self.area
        self.xr0 = (self.xr - x0)*msk
        self.yr0 = (self.yr - y0)*msk
        self.x2 = self.domain_integration((self.xr0)**2*msk) / self.area
        self.y2 = self.domain_integration((self.yr0)**2*msk) / self.area
        self.x0, self.y0 = x0, y0

    def fill_halo(self, x):
        fm.fillhalo(x, self.nh)


# ---------------------------------------------------------------------------
# multigrid
# ---------------------------------------------------------------------------
def level_sizes(n, m, n0=4):
    """single-rank Gridinfo (level.py:24-117): halve until a dimension is <= n0"""
    sizes = []
    lev = 0
    while True:
        if lev > 0:
            n, m = n//2, m//2
        sizes.append((m, n))
        lev += 1
        if n <= n0 or m <= n0:
            return sizes
        if lev > 20:
            raise RuntimeError('too many levels')


class MG(object):
    npre = 1
    npost = 1
    ndeepest = 16
    nvcyc = 1

    def __init__(self, cornermask, n, m, dx, dy, omega=8./9., hydroepsilon=1., Rd=None, relaxation='default'):
        self.nh = NH
        self.omega = omega
        # level.py:153-163: the line (tridiagonal) relaxation is taken when asked for
        # (param.relaxation) or when the cells are flat (hydroepsilon*dy/dx <= 0.2)
        self.relaxation = 'tridiagonal' if hydroepsilon*dy/dx <= .2 else relaxation
        self.sizes = level_sizes(n, m)
        self.nlevs = len(self.sizes)
        self.msk, self.A, self.x, self.b, self.r = [], [], [], [], []
        nh = self.nh
        A9 = None
        for lev, (ml, nl) in enumerate(self.sizes):
            mv, nv = ml+2*nh, nl+2*nh
            if lev == 0:
                msk = cornermask.astype(np.int8)
                A9 = self._finest_matrix(msk, dx, dy, hydroepsilon)
            else:
                prev = self.msk[lev-1]
                w = np.ones((mv, nv))
                fm.restrict(np.ones((mv, nv), dtype=np.int8), prev*1.0, nh, w)
                fm.fillhalo(w, nh)
                msk = np.ones((mv, nv), dtype=np.int8)
                msk[w <= 0.5] = 0
                A9 = fm.coarsenmatrix(A9, prev, msk, nh)
                for k in range(9):
                    d = np.ascontiguousarray(A9[:, :, k])
                    fm.fillhalo(d, nh)
                    A9[:, :, k] = d
            self.msk.append(msk)
            self.A.append(np.ascontiguousarray(A9[:, :, :5]))
            self.x.append(np.zeros((mv, nv)))
            self.b.append(np.zeros((mv, nv)))
            self.r.append(np.zeros((mv, nv)))
        if Rd is not None:   # Helmholtz operator of the QG model, hierarchy.py:80-84
            for lev in range(self.nlevs):
                d = self.A[lev][:, :, 4]
                d[d != 0.] -= 1./(Rd**2)

    def _finest_matrix(self, msk, dx, dy, hydroepsilon):
        mv, nv = msk.shape
        bx = dy/dx*hydroepsilon
        by = dx/dy
        a = -2*(bx+by)
        st = np.array([[0., by, 0.], [bx, a, bx], [0., by, 0.]])
        if dx == dy and hydroepsilon == 1:
            a, b, c = -6./2, 1./2, 0.5/2
            st = np.array([[c, b, c], [b, a, b], [c, b, c]])
        coef = 1./(dx*dy)
        A9 = np.zeros((mv, nv, 9))
        for k in range(9):
            i, j = (k % 3)-1, (k//3)-1
            d = np.zeros((mv, nv))
            d[1:-1, 1:-1] = (st[j+1, i+1]*coef * msk[1+j:mv-1+j, 1+i:nv-1+i]
                             * msk[1:-1, 1:-1])
            fm.fillhalo(d, self.nh)
            A9[:, :, k] = d
        return A9

    # -- per-level operators (level.py:340-496) -----------------------------
    def smooth(self, lev, x, b, nite):
        if self.relaxation == 'tridiagonal':     # level.py:340-349
            for _ in range(3*nite):
                fm.smoothtridiag(self.msk[lev], self.A[lev], x, b)
                fm.fillhalo(x, self.nh)
            return
        for _ in range(nite):
            fm.smoothtwicewitha(self.msk[lev], self.A[lev], x, b, self.omega)
            fm.fillhalo(x, self.nh)

    def residual(self, lev, x, b, r):
        fm.computeresidualwitha(self.msk[lev], self.A[lev], x, b, r)
        fm.fillhalo(r, self.nh)

    def norm(self, lev, x):
        return np.sqrt(fm.computenorm(self.msk[lev], x, self.nh))

    def down(self, lev, xf, xc):
        fm.restrict(self.msk[lev+1], xf, self.nh, xc)
        fm.fillhalo(xc, self.nh)

    def up(self, lev, xc, xf):
        fm.interpolate(self.msk[lev], self.msk[lev+1], xc, self.nh, xf)

    # -- cycles (hierarchy.py:98-218) ---------------------------------------
    def vcycle(self, lev1):
        last = self.nlevs-1
        x, b, r = self.x, self.b, self.r
        for lev in range(lev1, last):
            if lev > lev1:
                x[lev][:, :] = 0.
            self.smooth(lev, x[lev], b[lev], self.npre)
            self.residual(lev, x[lev], b[lev], r[lev])
            self.down(lev, r[lev], b[lev+1])
        x[last][:, :] = 0.
        self.smooth(last, x[last], b[last], self.ndeepest)
        for lev in range(last-1, lev1-1, -1):
            self.up(lev, x[lev+1], r[lev])
            x[lev] += r[lev]
            self.smooth(lev, x[lev], b[lev], self.npost)

    def fcycle(self, lev1):
        last = self.nlevs-1
        x, b = self.x, self.b
        for lev in range(lev1, last):
            self.down(lev, b[lev], b[lev+1])
        x[last][:, :] = 0.
        self.smooth(last, x[last], b[last], self.ndeepest)
        for lev in range(last-1, lev1-1, -1):
            self.up(lev, x[lev+1], x[lev])
            for _ in range(self.nvcyc):
                self.vcycle(lev)

    def solve(self, x, b, maxite=4, tol=1e-11):
        self.residual(0, x, b, self.b[0])
        normb = self.norm(0, b)
        if not normb > 0:
            return 0, 0.
        res0 = self.norm(0, self.b[0])/normb
        res = res0
        nite = 0
        ndiv = 0
        while nite < maxite and res0 > tol:
            self.fcycle(0)
            x += self.x[0]
            self.residual(0, x, b, self.b[0])
            res = self.norm(0, self.b[0])/normb
            conv = res0/res
            res0 = res
            nite += 1
            if conv < 1:
                ndiv += 1
            if ndiv > 4:
                raise RuntimeError('solver is not converging')
        return nite, res

    def two_vcycle(self, x, b):
        self.x[0][:, :] = x
        self.b[0][:, :] = b
        for _ in range(2):
            self.residual(0, self.x[0], self.b[0], self.r[0])
            self.vcycle(0)
        x[:, :] = self.x[0]
        return 1, 0.


# ---------------------------------------------------------------------------
# operators
# ---------------------------------------------------------------------------
class Ops(object):
    def __init__(self, param, grid, varnames, tracers, whosetspsi, qg=False):
        self.p, self.g = param, grid
        self.varnames, self.tracers, self.whosetspsi = varnames, tracers, whosetspsi
        self.nh = grid.nh
        self.msk = grid.msk
        self.dx, self.dy = grid.dx, grid.dy
        shape = (grid.nyl, grid.nxl)
        self.work = np.zeros(shape)
        fo.celltocorner(self.msk*1., self.work)
        self.work[self.work < 1.] = 0.
        self.mskp = self.msk*0
        self.mskp[self.work == 1.] = 1
        self.gmg = MG(self.work, param.nx, param.ny, grid.dx, grid.dy,
                      hydroepsilon=param.hydroepsilon, Rd=(param.Rd if qg else None),
                      relaxation=getattr(param, 'relaxation', 'default'))
        self.fill_halo = grid.fill_halo
        # boundary mask of the no-slip source (operators.py:157-186)
        ns = grid.msknoslip
        z = (np.roll(ns, -1, axis=1)+np.roll(ns, -1, axis=0)
             + np.roll(ns, +1, axis=1)+np.roll(ns, +1, axis=0)-4*ns)
        z = z*ns
        self.mskbc = self.msk*0
        self.mskbc[z < 0] = 1
        self.mskbc *= ns
        self.fill_halo(self.mskbc)
        self.bcarea = grid.domain_integration(self.mskbc)
        self.x2bc = grid.domain_integration((grid.xr0)**2*self.mskbc*ns)
        self.y2bc = grid.domain_integration((grid.yr0)**2*self.mskbc*ns)
        self.cst = np.zeros(5)
        self.cst[0], self.cst[1], self.cst[2] = grid.dx, grid.dy, 0.05
        if param.order % 2 == 0:
            self.adv = fa.adv_centered
        else:
            self.adv = fa.adv_upwind
            self.cst[4] = param.aparab
        methods = ['minmax', 'parabolic']
        fs = param.flux_splitting_method
        self.fs_method = methods.index(fs) if fs in methods else 1
        K_ = param.Kdiff
        self.Kdiff = K_ if isinstance(K_, dict) else {t: K_ for t in tracers}
        self.first_time = True
        self.rhsp = None
        self.psi = None

    def ix(self, name):
        return self.varnames.index(name)

    def rhs_adv(self, x, t, dxdt):
        u, v = x[self.ix('u')], x[self.ix('v')]
        for trac in self.tracers:
            k = self.ix(trac)
            self.adv(self.msk, x[k], dxdt[k], u, v, self.cst, self.nh, self.fs_method,
                     self.p.order)
            self.fill_halo(dxdt[k])

    def rhs_diffusion(self, x, t, dxdt, coef=1.):
        for trac in self.tracers:
            k = self.ix(trac)
            fo.add_diffusion(self.msk, x[k], self.dx, self.nh, coef*self.Kdiff[trac], dxdt[k])
            self.fill_halo(dxdt[k])

    def rhs_torque(self, x, t, dxdt):
        y = dxdt[self.ix('vorticity')]
        y *= self.msk
        fo.add_torque(self.msk, x[self.ix('buoyancy')], self.dx, self.nh, self.p.gravity, y)
        self.fill_halo(y)

    def rhs_noslip(self, x, source):
        g = self.g
        ip, iw = self.ix('psi'), self.ix(self.whosetspsi)
        fo.cornertocell(x[ip], self.work)
        fo.computenoslipsourceterm(g.msknoslip, x[ip], self.work, self.dx, self.dy, self.nh)
        source[:, :] = self.work
        mean = g.domain_integration(source) / self.bcarea
        source -= mean*self.mskbc
        if self.p.enforce_momentum:
            px = fd.computedotprod(self.msk, source, g.xr0, self.nh)
            py = fd.computedotprod(self.msk, source, g.yr0, self.nh)
            px, py = px/self.x2bc, py/self.y2bc
            source -= (px*g.xr0+py*g.yr0)*self.mskbc
        self.fill_halo(source)
        x[iw] -= source

    def invert_vorticity(self, x, flag='full', island=False):
        iu, iv, ip = self.ix('u'), self.ix('v'), self.ix('psi')
        iw = self.ix(self.whosetspsi)
        psi = x[ip]
        fo.celltocorner(x[iw], self.work)
        if island:
            self.work[:, :] -= self.rhsp
        if flag == 'fast':
            self.last = self.gmg.two_vcycle(psi, self.work)
        else:
            self.last = self.gmg.solve(psi, self.work, maxite=4, tol=1e-11)
            if self.g.geometry == 'perio':
                psim = self.g.domain_integration(psi) / self.g.area
                psi -= psim
        psi = psi*self.mskp
        if island:
            psi += self.psi
        self.first_time = False
        fo.computeorthogradient(self.msk, psi, self.dx, self.dy, self.nh, x[iu], x[iv])
        x[ip] = psi


# ---------------------------------------------------------------------------
# time schemes (whole-state combinations; same numpy expressions as timescheme.py)
# ---------------------------------------------------------------------------
class Stepper(object):
    KFORCING = {'RK4_LS': 3, 'RK3_SSP': 2, 'Heun': 1, 'LFAM3': 1}
    DTCOEF = {'Heun': 2., 'RK3_SSP': 1.5}

    def __init__(self, name, state, rhs, fast_axpy=False):
        self.name = name
        self.rhs = rhs
        self.kforcing = self.KFORCING.get(name, 0)
        self.dtcoef = self.DTCOEF.get(name, 1.)
        self.kstage = 0
        self.x = np.zeros_like(state)
        self.dx0 = np.zeros_like(state)
        self.dx1 = np.zeros_like(state)
        self.dx2 = np.zeros_like(state)
        self.xb = np.zeros_like(state)
        self.first = True
        self.second = True
        self.asselin_cst = 0.1
        self.ab2_epsilon = 0.1
        self.fast_axpy = fast_axpy    # OpenMP axpys (same bits as numpy), CPU baseline
        self.forward = getattr(self, 'step_'+name)

    def step_EF(self, x, t, dt):
        self.rhs(x, t, self.dx0)
        x += dt * self.dx0

    def step_AB2(self, x, t, dt):
        self.rhs(x, t, self.dx0)
        if self.first:
            x += dt * self.dx0
            self.first = False
        else:
            x += ((1.5+self.ab2_epsilon)*dt) * self.dx0 - ((0.5+self.ab2_epsilon)*dt)*self.dx1
        self.dx1[:] = self.dx0

    def step_AB3(self, x, t, dt):
        self.rhs(x, t, self.dx0)
        if self.first:
            x += dt * self.dx0
            self.first = False
        elif self.second:
            x += (1.5*dt) * self.dx0 - (0.5*dt)*self.dx1
            self.second = False
        else:
            x += (23*dt/12.) * self.dx0 - (16*dt/12.)*self.dx1+(5*dt/12.)*self.dx2
        self.dx2[:] = self.dx1
        self.dx1[:] = self.dx0

    def step_LF(self, x, t, dt):
        self.x[:] = x
        self.rhs(x, t, self.dx0)
        if self.first:
            x += dt * self.dx0
            self.first = False
        else:
            x[:] = self.xb + (2*dt) * self.dx0
            self.x += self.asselin_cst*(x+self.xb-2*self.x)
        self.xb[:] = self.x

    def step_RK3(self, x, t, dt):
        self.kstage = 0
        self.rhs(x, t, self.dx0)
        self.x = x + (dt/3.) * self.dx0
        self.kstage = 1
        self.rhs(self.x, t+dt/3., self.dx1)
        self.x = x + (0.5*dt)*self.dx1
        self.kstage = 2
        self.rhs(self.x, t+0.5*dt, self.dx2)
        x += dt*self.dx2

    def step_RK4_LS(self, x, t, dt):
        self.kstage = 0
        self.rhs(x, t, self.dx0)
        self.x = x + (0.25*dt) * self.dx0
        self.kstage = 1
        self.rhs(self.x, t+dt*0.25, self.dx0)
        self.x = x + (dt/3.)*self.dx0
        self.kstage = 2
        self.rhs(self.x, t+dt/3., self.dx0)
        self.x = x + (dt/2.)*self.dx0
        self.kstage = 3
        self.rhs(self.x, t+0.5*dt, self.dx0)
        x += dt*self.dx0

    def step_Heun(self, x, t, dt):
        self.kstage = 0
        self.rhs(x, t, self.dx0)
        self.x = x + dt * self.dx0
        self.kstage = 1
        self.rhs(self.x, t+dt, self.dx1)
        x += (0.5*dt)*(self.dx0+self.dx1)

    def step_RK3_SSP(self, x, t, dt):
        self.kstage = 0
        self.rhs(x, t, self.dx0)
        if self.fast_axpy:
            K.axpy1(self.x, x, dt, self.dx0)
        else:
            self.x = x + dt * self.dx0
        self.kstage = 1
        self.rhs(self.x, t+dt, self.dx1)
        if self.fast_axpy:
            K.axpy2(self.x, x, 0.25*dt, self.dx0, self.dx1)
        else:
            self.x = x + (0.25*dt)*(self.dx0+self.dx1)
        self.kstage = 2
        self.rhs(self.x, t+0.5*dt, self.dx2)
        if self.fast_axpy:
            K.axpy3(x, dt/6., self.dx0, self.dx1, self.dx2)
        else:
            x += (dt/6.)*(self.dx0+self.dx1+4*self.dx2)

    def step_LFAM3(self, x, t, dt):
        self.x[:] = x
        self.kstage = 0
        self.rhs(x, t, self.dx0)
        if self.first:
            x += dt * self.dx0
            self.first = False
        else:
            x[:] = self.xb + (2*dt) * self.dx0
            x[:] = (1./12.)*(5.*x + 8.*self.x-self.xb)
            self.kstage = 1
            self.rhs(x, t+dt*.5, self.dx0)
            x[:] = self.x + dt*self.dx0
        self.xb[:] = self.x


# ---------------------------------------------------------------------------
# models
# ---------------------------------------------------------------------------
class Var(object):
    def __init__(self, names, shape):
        self.varname_list = names
        self.state = np.zeros([len(names)]+list(shape))

    def get(self, name):
        return self.state[self.varname_list.index(name)]


class EulerModel(object):
    def __init__(self, param, grid, fast_axpy=False):
        self.p, self.g = param, grid
        names = ['vorticity', 'psi', 'u', 'v', 'source']
        tracers = ['vorticity']
        for extra in getattr(param, 'additional_tracer', []):
            names.append(extra)
            tracers.append(extra)
        param.varname_list, param.tracer_list = names, tracers
        self.var = Var(names, (grid.nyl, grid.nxl))
        self.msk, self.nh = grid.msk, grid.nh
        self.ope = Ops(param, grid, names, tracers, 'vorticity')
        self.tscheme = Stepper(param.timestepping, self.var.state, self.dynamics, fast_axpy)
        if param.spongelayer:
            self.spongemsk = (1-(1+np.tanh((grid.xr - grid.Lx)/0.1))*0.5)
        self.diags = {}
        self.forc = None

    def dynamics(self, x, t, dxdt):
        p, ope = self.p, self.ope
        ope.rhs_adv(x, t, dxdt)
        if self.tscheme.kstage == self.tscheme.kforcing:
            if p.forcing:
                self.forc.add_forcing(x, t, dxdt)
            if p.diffusion:
                ope.rhs_diffusion(x, t, dxdt)
            if p.diffusion or p.forcing:
                ope.invert_vorticity(dxdt, flag='fast')
        else:
            ope.invert_vorticity(dxdt, flag='fast')

    def step(self, t, dt):
        p = self.p
        state = self.var.state
        self.tscheme.forward(state, t, dt)
        if p.noslip:
            source = self.var.get('source')
            self.ope.rhs_noslip(state, source)
            self.ope.invert_vorticity(state, flag='fast', island=p.isisland)
            source /= dt
        if p.spongelayer:
            w = self.var.get('vorticity')
            w *= self.spongemsk
        self.set_psi_from_vorticity()

    def set_psi_from_vorticity(self):
        self.ope.invert_vorticity(self.var.state, island=self.p.isisland)

    def diagnostics(self, var, t):
        g, nh, msk = self.g, self.nh, self.msk
        u, v = var.get('u'), var.get('v')
        w, psi, src = var.get('vorticity'), var.get('psi'), var.get('source')
        ke, maxu = fd.computekemaxu(msk, u, v, nh)
        z, z2 = fd.computesumandnorm(msk, w, nh)
        px = fd.computedotprod(msk, w, g.xr, nh)
        py = fd.computedotprod(msk, w, g.yr, nh)
        angmom = fd.computesum(msk, psi, nh)
        sce = fd.computedotprod(msk, w, src, nh)
        area = g.area
        d = self.diags
        d['maxspeed'] = np.float64(maxu)
        d['ke'] = ke / area
        d['vorticity'] = z / area
        d['enstrophy'] = 0.5*z2 / area
        d['px'] = px / area
        d['py'] = py / area
        d['angmom'] = angmom / area
        d['source'] = sce / area


class BoussinesqModel(object):
    def __init__(self, param, grid, fast_axpy=False):
        self.p, self.g = param, grid
        names = ['vorticity', 'psi', 'u', 'v', 'buoyancy', 'banom']
        tracers = ['vorticity', 'buoyancy']
        for extra in getattr(param, 'additional_tracer', []):
            names.append(extra)
            tracers.append(extra)
        param.varname_list, param.tracer_list = names, tracers
        self.var = Var(names, (grid.nyl, grid.nxl))
        self.bref = self.var.get('buoyancy').copy()
        self.source = np.zeros((grid.nyl, grid.nxl))
        self.msk, self.nh = grid.msk, grid.nh
        self.ope = Ops(param, grid, names, tracers, 'vorticity')
        self.tscheme = Stepper(param.timestepping, self.var.state, self.dynamics, fast_axpy)
        self.diags = {}
        self.forc = None

    def dynamics(self, x, t, dxdt):
        p, ope = self.p, self.ope
        ope.rhs_adv(x, t, dxdt)
        ope.rhs_torque(x, t, dxdt)
        if self.tscheme.kstage == self.tscheme.kforcing:
            coef = self.tscheme.dtcoef
            if p.forcing:
                self.forc.add_forcing(x, t, dxdt, coef=coef)
            if p.diffusion:
                ope.rhs_diffusion(x, t, dxdt, coef=coef)
        ope.invert_vorticity(dxdt, flag='fast')

    def step(self, t, dt):
        p = self.p
        state = self.var.state
        self.tscheme.forward(state, t, dt)
        if p.noslip:
            self.ope.rhs_noslip(state, self.source)
            self.ope.invert_vorticity(state, flag='fast', island=p.isisland)
        banom = self.var.get('banom')
        banom[:, :] = self.var.get('buoyancy')-self.bref

    def set_psi_from_vorticity(self):
        self.ope.invert_vorticity(self.var.state, island=self.p.isisland)

    def diagnostics(self, var, t):
        g, nh, msk = self.g, self.nh, self.msk
        u, v = var.get('u'), var.get('v')
        w, buoy = var.get('vorticity'), var.get('buoyancy')
        ke, maxu = fd.computekemaxu(msk, u, v, nh)
        z, z2 = fd.computesumandnorm(msk, w, nh)
        b, b2 = fd.computesumandnorm(msk, buoy, nh)
        pe = - self.p.gravity * fd.computesum(msk, buoy*g.yr, nh)
        area = g.area
        d = self.diags
        d['maxspeed'] = np.float64(maxu)
        d['ke'] = ke / area
        d['pe'] = pe / area
        d['energy'] = (ke+pe) / area
        d['vorticity'] = z / area
        d['enstrophy'] = 0.5*z2 / area
        d['buoyancy'] = b / area
        d['brms'] = np.sqrt(b2 / area-(b/area)**2)


class Fluxes(object):
    """core/fluxes.py: one step forward, one step with every velocity-like field reversed;
    half sum / half difference of the time-integrated face fluxes = reversible /
    irreversible parts"""

    def __init__(self, param, grid, ope):
        self.ope = ope
        self.modelname = param.modelname
        self.order = param.order
        self.tracers = list(param.tracer_list) + ['uc', 'vc']
        names = list(param.varname_list) + ['uc', 'vc']
        self.nvarstate = len(names)
        self.flx_list = ['flx_%s_%s' % (v, d) for v in self.tracers for d in 'xy']
        self.varnames = names + self.flx_list
        self.fullflx_list = ['%s_%s_%s' % (r, d, v) for r in ('rev', 'irr') for v in self.tracers for d in 'xy']
        shape = [grid.nyl, grid.nxl]
        self.x = np.zeros([len(self.varnames)]+shape)
        self.xe = np.zeros([self.nvarstate]+shape)
        self.xwork = np.zeros([len(self.varnames)]+shape)
        self.flx = np.zeros([len(self.fullflx_list)]+shape)
        self.adv = ff.adv_centered if self.order % 2 == 0 else ff.adv_upwind
        self.cst = np.zeros(5)
        self.cst[0], self.cst[1], self.cst[2], self.cst[4] = grid.dx, grid.dy, 0.05, param.aparab
        self.msk, self.nh = grid.msk, grid.nh
        self.tscheme = Stepper(param.timestepping, self.x, self.advection)

    def ix(self, name):
        return self.varnames.index(name)

    def diag_fluxes(self, x, t, dt):
        nvs = self.nvarstate
        iu, iv, ip, iw = self.ix('u'), self.ix('v'), self.ix('psi'), self.ix('vorticity')
        self.xe[:nvs-2] = x
        y = 0.5*(x[iu]+np.roll(x[iu], 1, axis=1))
        self.ope.fill_halo(y)
        self.xe[nvs-2] = y
        y = 0.5*(x[iv]+np.roll(x[iv], 1, axis=0))
        self.ope.fill_halo(y)
        self.xe[nvs-1] = y
        self.x[:nvs] = self.xe
        self.x[nvs:] = 0.
        self.tscheme.forward(self.x, t, dt)
        self.xwork[:] = self.x
        self.x[:nvs] = self.xe
        for k in (iu, iv, ip, iw, nvs-2, nvs-1):
            self.x[k] *= -1
        self.x[nvs:] = 0.
        self.tscheme.forward(self.x, t+dt, -dt)
        cff = 0.5/dt
        nflx = len(self.flx_list)
        for k in range(nflx):
            ell = nvs+k
            sign = -1 if (k < 2) or (k >= nflx-4) else 1
            self.flx[k] = cff*(self.xwork[ell]+sign*self.x[ell])
            self.flx[nflx+k] = cff*(self.xwork[ell]-sign*self.x[ell])

    def advection(self, x, t, dxdt):
        self.rhs_adv(x, t, dxdt)
        if self.modelname == 'boussinesq':
            self.ope.rhs_torque(x, t, dxdt)
        self.ope.invert_vorticity(dxdt, flag='fast')

    def rhs_adv(self, x, t, dxdt):
        u, v = x[self.ix('u')], x[self.ix('v')]
        self.cst[3] = self.ope.cst[3]
        nvs = self.nvarstate
        fs = self.ope.fs_method
        for it, trac in enumerate(self.tracers):
            k = self.ix(trac)
            self.adv(self.msk, x[k], dxdt[k], u, v, dxdt[nvs+2*it], dxdt[nvs+2*it+1], self.cst,
                     self.nh, fs, self.order)
            for f in (k, nvs+2*it, nvs+2*it+1):
                self.ope.fill_halo(dxdt[f])


class Fluid2d(object):
    def __init__(self, param, grid, fast_axpy=False):
        self.p, self.g = param, grid
        self.dx, self.dy = grid.dx, grid.dy
        self.dt = self.dt0 = param.dt
        grid.finalize_msk()
        self.enforce_momentum = param.enforce_momentum
        if param.modelname == 'euler':
            if param.geometry not in ('closed', 'disc'):
                self.enforce_momentum = False
            self.model = EulerModel(param, grid, fast_axpy)
        elif param.modelname == 'boussinesq':
            self.enforce_momentum = False
            self.model = BoussinesqModel(param, grid, fast_axpy)
        else:
            raise NotImplementedError('oracle model: %s' % param.modelname)
        if param.isisland:
            grid.island.finalize()
            self.model.ope.rhsp = grid.island.rhsp
            self.model.ope.psi = grid.island.psi
        self.diag_fluxes = bool(getattr(param, 'diag_fluxes', False))
        if self.diag_fluxes:
            self.flx = Fluxes(param, grid, self.model.ope)
        self.t = 0.
        self.kt = 0

    def set_dt(self, kt):
        p = self.p
        maxspeed = self.model.diags['maxspeed']
        if p.adaptable_dt and maxspeed != 0:
            self.dt = p.cfl * min(self.dx, self.dy) / maxspeed
            if self.dt > p.dtmax:
                self.dt = p.dtmax
        else:
            self.dt = self.dt0
        self.model.ope.cst[3] = maxspeed

    def enforce_zero_momentum(self):
        if self.enforce_momentum:
            g = self.g
            vor = self.model.var.get('vorticity')
            px = self.model.diags['px']/g.x2
            py = self.model.diags['py']/g.y2
            vor[:] -= (g.xr0 * px + g.yr0 * py)
            self.model.ope.invert_vorticity(self.model.var.state, flag='fast')
