"""QG: quasi-geostrophic model, prognostic full PV (reference: core/quasigeostrophic.py).
The elliptic operator is the Helmholtz operator (Laplacian - 1/Rd^2), built by the
device multigrid when param.qgoperator is set (hierarchy.py:80-84)."""
import numpy as np
import torch

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
        if param.bottom_torque or param.ageostrophic:
            raise NotImplementedError('QG: bottom_torque / ageostrophic diagnostics are not built yet')
        # the full PV is advected; the inversion reads its anomaly w.r.t. the background beta*y
        declare_state(param, grid, ['pv', 'psi', 'u', 'v', 'pvanom', 'vorticity'], ['pv'], 'pvanom')
        self.var = Var(param)
        r = rt()
        self.rt = r
        self.ncell = grid.nyl*grid.nxl
        self.source = torch.zeros((grid.nyl, grid.nxl), dtype=torch.float64, device=r.device)
        ix = self.var.index
        self.ipv, self.ipsi, self.ipva, self.ivor = ix('pv'), ix('psi'), ix('pvanom'), ix('vorticity')
        self.pvback = self.beta*(grid.yr-grid.Ly*.5)*grid.msk
        self.d_pvback = r.to_device(self.pvback, dtype=np.float64)
        param.qgoperator = True      # Helmholtz operator in the multigrid (hierarchy.py:80-84)
        self.ope = Operators(param, grid)
        self.tscheme = Timescheme(param, self.var.dstate)
        self.dx0 = self.tscheme.dx0
        self.kt = 0
        if self.forcing and self.forcing_module != 'embedded':
            self.forc = user_object(self.forcing_module, 'Forcing', param, grid, 'forcing')
        self.diags = {}
        self.tscheme.set(self.dynamics, self.timestepping)

    def step(self, t, dt):
        r, lib = self.rt, self.rt.lib
        s = self.var.dstate
        self.dt = dt
        self.tscheme.forward(s, t, dt)
        if self.noslip:
            self.add_noslip(s)
        self.set_psi_from_pv()
        lib.set_sum(s.wptr(self.ipva), s.rptr(self.ipv), -1., r.ptr(self.d_pvback), self.ncell, r.stream)
        lib.set_sum(s.wptr(self.ivor), s.rptr(self.ipva), -(self.Rd**-2), s.rptr(self.ipsi), self.ncell, r.stream)

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
