"""Advection: passive tracer in a prescribed steady flow (reference: core/advection.py)."""
import numpy as np

from operators import Operators
from variables import Var
from timescheme import Timescheme
from runtime import rt


class Advection(object):
    def __init__(self, param, grid):
        self.list_param = ['timestepping', 'diffusion', 'Kdiff']
        param.copy(self, self.list_param)
        self.list_grid = ['msk', 'nh', 'area', 'mpitools']
        grid.copy(self, self.list_grid)
        param.varname_list = ['tracer', 'psi', 'u', 'v', 'vorticity']
        param.sizevar = [grid.nyl, grid.nxl]
        self.var = Var(param)
        param.tracer_list = ['tracer']
        param.whosetspsi = ('tracer')
        self.ope = Operators(param, grid)
        self.tscheme = Timescheme(param, self.var.dstate)
        self.tscheme.set(self.advection, self.timestepping)
        self.rt = rt()
        self.diags = {}

    def step(self, t, dt):
        self.tscheme.forward(self.var.dstate, t, dt)
        self.diagnostics(self.var, t)

    def advection(self, x, t, dxdt):
        self.ope.rhs_adv(x, t, dxdt)
        if self.tscheme.kstage == self.tscheme.kforcing:
            if self.diffusion:
                self.ope.rhs_diffusion(x, t, dxdt)

    def set_psi_from_tracer(self):
        self.ope.invert_vorticity(self.var.dstate)

    def diagnostics(self, var, t):
        import ctypes
        r, lib = self.rt, self.rt.lib
        s = var.dstate
        msk, sc = r.ptr(self.ope.d_msk), r.ptr(r.scratch)
        if t == 0.:
            lib.computekemaxu(msk, s.rptr(var.index('u')), s.rptr(var.index('v')), self.nh, s.ny, s.nx,
                              r.ptr(r.out), sc, r.stream)
            ke, maxu = r.read_out(2)
            cst = self.mpitools.local_to_global([(maxu, 'max'), (ke, 'sum')])
            self.diags['maxspeed'] = cst[0]
            self.diags['ke'] = cst[1] / self.area
            self.diags['enstrophy'] = 0.
        lib.computesumandnorm(msk, s.rptr(var.index('tracer')), self.nh, s.ny, s.nx, r.ptr(r.out), sc, r.stream)
        z, z2 = r.read_out(2)
        cst = self.mpitools.local_to_global([(z, 'sum'), (z2, 'sum')])
        self.diags['mean'] = cst[0] / self.area
        self.diags['rms'] = np.sqrt(cst[1] / self.area)
