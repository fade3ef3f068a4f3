"""QG: quasi-geostrophic model, prognostic full PV (reference: core/quasigeostrophic.py).
The elliptic operator is the Helmholtz operator (Laplacian - 1/Rd^2), built by the
device multigrid when param.qgoperator is set (hierarchy.py:80-84)."""
import numpy as np

from modelbase import adopt, declare_state, user_object
from operators import Operators
from variables import Var
from timescheme import Timescheme
from runtime import rt

FROM_PARAM = ('timestepping', 'forcing', 'forcing_module', 'diffusion', 'Kdiff', 'noslip', 'beta', 'Rd',
              'ageostrophic', 'bottom_torque')
FROM_GRID = ('nh', 'msk', 'area', 'yr', 'isisland', 'mpitools')


class QG(object):
    def __init__(self, param, grid):
        adopt(self, param, FROM_PARAM)
        adopt(self, grid, FROM_GRID)
        # the full PV is advected; the inversion reads its anomaly w.r.t. the background beta*y.
        # Optional diagnosed fields ride along as advected tracers (quasigeostrophic.py:34-58):
        # 'btorque' (bottom torque) and 'ua', 'va' (ageostrophic velocity)
        extra = (['btorque'] if param.bottom_torque else []) + (['ua', 'va'] if param.ageostrophic else [])
        declare_state(param, grid, ['pv', 'psi', 'u', 'v', 'pvanom', 'vorticity'] + extra, ['pv'] + extra, 'pvanom')
        self.var = Var(param)
        r = rt()
        self.rt = r
        self.ncell = grid.nyl*grid.nxl
        self.source = r.alloc((grid.nyl, grid.nxl))    # (symmetric heap on y-slabs: its halo rows are exchanged)
        ix = self.var.index
        self.ipv, self.ipsi, self.ipva, self.ivor = ix('pv'), ix('psi'), ix('pvanom'), ix('vorticity')
        self.pvback = self.beta*(grid.yr-grid.Ly*.5)*grid.msk
        self.d_pvback = r.to_device(self.pvback, dtype=np.float64)
        param.qgoperator = True      # Helmholtz operator in the multigrid (hierarchy.py:80-84)
        self.ope = Operators(param, grid)
        self.tscheme = Timescheme(param, self.var.dstate)
        self.dx0 = self.tscheme.dx0
        self.kt = 0
        self._ageo = None
        if self.forcing and self.forcing_module != 'embedded':
            self.forc = user_object(self.forcing_module, 'Forcing', param, grid, 'forcing')
        self.diags = {}
        self.tscheme.set(self.dynamics, self.timestepping)

    def step(self, t, dt):
        r, lib = self.rt, self.rt.lib
        s = self.var.dstate
        n, nbytes = self.ncell, self.ncell*8
        ix = self.var.index
        if self.bottom_torque:
            # the background PV is transported for one step; what it has changed by is the
            # torque -J(psi, htopo) (quasigeostrophic.py:96-98, 124-131)
            ibt = ix('btorque')
            lib.copy(s.wptr(ibt), r.ptr(self.d_pvback), nbytes, r.stream)
        if self.ageostrophic:
            iu, iv, iua, iva = ix('u'), ix('v'), ix('ua'), ix('va')
            lib.copy(s.wptr(iua), s.rptr(iu), nbytes, r.stream)
            lib.copy(s.wptr(iva), s.rptr(iv), nbytes, r.stream)
        self.dt = dt
        self.tscheme.forward(s, t, dt)
        if self.noslip:
            self.add_noslip(s)
        self.set_psi_from_pv()
        lib.set_sum(s.wptr(self.ipva), s.rptr(self.ipv), -1., r.ptr(self.d_pvback), n, r.stream)
        lib.set_sum(s.wptr(self.ivor), s.rptr(self.ipva), -(self.Rd**-2), s.rptr(self.ipsi), n, r.stream)
        if self.bottom_torque:
            lib.set_sum(s.wptr(ibt), s.rptr(ibt), -1., r.ptr(self.d_pvback), n, r.stream)
            lib.div_scalar(s.wptr(ibt), dt, n, r.stream)
        if self.ageostrophic:
            # ua = -(vg - va')/dt - pvback*ug ;  va = +(ug - ua')/dt - pvback*vg   with ua', va' the
            # transported copies of (u, v)  (quasigeostrophic.py:133-153)
            if self._ageo is None:
                self._ageo = r.alloc((s.ny, s.nx))
            wa, wb, wc = r.ptr(self.ope.work), r.ptr(self.ope.work2), r.ptr(self._ageo)
            lib.set_sum(wa, s.rptr(iv), -1., s.rptr(iva), n, r.stream)
            lib.div_scalar(wa, dt, n, r.stream)
            lib.scale(wa, -1., n, r.stream)
            lib.copy(wc, s.rptr(iu), nbytes, r.stream)
            lib.mul_field(wc, r.ptr(self.d_pvback), n, r.stream)
            lib.add_scaled(wa, -1., wc, n, r.stream)
            lib.set_sum(wb, s.rptr(iu), -1., s.rptr(iua), n, r.stream)
            lib.div_scalar(wb, dt, n, r.stream)
            lib.copy(wc, s.rptr(iv), nbytes, r.stream)
            lib.mul_field(wc, r.ptr(self.d_pvback), n, r.stream)
            lib.add_scaled(wb, -1., wc, n, r.stream)
            lib.copy(s.wptr(iva), wb, nbytes, r.stream)
            lib.copy(s.wptr(iua), wa, nbytes, r.stream)

    def dynamics(self, x, t, dxdt):
        r, lib = self.rt, self.rt.lib
        lib.zero(dxdt.all_ptr(True), dxdt.size*8, r.stream)
        self.ope._barrier()
        self.ope.rhs_adv(x, t, dxdt)
        if self.tscheme.kstage == self.tscheme.kforcing:
            if self.forcing:
                self.forc.add_forcing(x, t, dxdt)
            if self.diffusion:
                self.ope.rhs_diffusion(x, t, dxdt)
        else:
            lib.copy(dxdt.wptr(self.ipva), dxdt.rptr(self.ipv), self.ncell*8, r.stream)
            self.ope.invert_vorticity(dxdt, flag='fast', island=self.isisland)

    def add_noslip(self, x):
        self.ope.rhs_noslip(x, self.source)
        self.ope.invert_vorticity(x, flag='fast', island=self.isisland)

    def add_backgroundpv(self):
        r, lib = self.rt, self.rt.lib
        s = self.var.dstate
        lib.add_scaled(s.wptr(self.ipv), 1., r.ptr(self.d_pvback), self.ncell, r.stream)

    def set_psi_from_pv(self):
        r, lib = self.rt, self.rt.lib
        s = self.var.dstate
        lib.set_sum(s.wptr(self.ipva), s.rptr(self.ipv), -1., r.ptr(self.d_pvback), self.ncell, r.stream)
        self.ope.invert_vorticity(s, flag='full', island=self.isisland)

    def diagnostics(self, var, t):
        import ctypes
        r, lib = self.rt, self.rt.lib
        s = var.dstate
        nh, ny, nx = self.nh, s.ny, s.nx
        msk, sc = r.ptr(self.ope.d_msk), r.ptr(r.scratch)

        def slot(k):
            return ctypes.c_void_p(r.out.data_ptr()+8*k)

        lib.cornertocell(s.rptr(self.ipsi), r.ptr(self.ope.work), ny, nx, r.stream)
        lib.computesumandnorm(msk, r.ptr(self.ope.work), nh, ny, nx, slot(0), sc, r.stream)
        lib.computekemaxu(msk, s.rptr(var.index('u')), s.rptr(var.index('v')), nh, ny, nx, slot(2), sc, r.stream)
        lib.computesumandnorm(msk, s.rptr(self.ipv), nh, ny, nx, slot(4), sc, r.stream)
        psim, psi2, ke, maxu, z, z2 = r.read_out(6)
        ape = 0.5 * psi2 / self.Rd**2
        cst = self.mpitools.local_to_global([(maxu, 'max'), (ke, 'sum'), (z, 'sum'), (z2, 'sum'), (ape, 'sum')])
        self.diags['maxspeed'] = cst[0]
        self.diags['ke'] = cst[1] / self.area
        self.diags['pv'] = cst[2] / self.area
        self.diags['pv2'] = 0.5*cst[3] / self.area
        self.diags['ape'] = cst[4] / self.area
        self.diags['energy'] = (cst[1]+cst[4]) / self.area
