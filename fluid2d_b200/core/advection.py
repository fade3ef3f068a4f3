"""Advection: passive tracer carried by a prescribed, steady flow (reference interface:
core/advection.py -- Advection(param, grid), .var, .ope, .tscheme, .diags, step(t, dt),
advection(x, t, dxdt), set_psi_from_tracer(), diagnostics(var, t)).

The flow is the orthogradient of a streamfunction that the user stores in 'tracer' once
(whosetspsi = 'tracer') and inverts with set_psi_from_tracer(); afterwards only the tracer
moves.  State, tendencies and reductions live on the device; a step costs the advection
kernel per stage, the RK combinations and two masked reductions (4 scalars to the host).
"""
import numpy as np

from operators import Operators
from runtime import rt
from timescheme import Timescheme
from variables import Var

STATE = ('tracer', 'psi', 'u', 'v', 'vorticity')


class Advection(object):
    def __init__(self, param, grid):
        # what the other classes read from param
        param.varname_list = list(STATE)
        param.tracer_list = [STATE[0]]
        param.whosetspsi = STATE[0]
        param.sizevar = [grid.nyl, grid.nxl]
        for name in ('timestepping', 'diffusion', 'Kdiff'):
            setattr(self, name, getattr(param, name))
        for name in ('msk', 'nh', 'area', 'mpitools'):
            setattr(self, name, getattr(grid, name))
        self.rt = rt()
        self.diags = {}
        self.var = Var(param)
        self.ope = Operators(param, grid)
        self.tscheme = Timescheme(param, self.var.dstate)
        self.tscheme.set(self.advection, self.timestepping)

    # -- time stepping ---------------------------------------------------------
    def advection(self, x, t, dxdt):
        """right-hand side: -div(u tracer) [+ diffusion at the forcing stage]"""
        self.ope.rhs_adv(x, t, dxdt)
        ts = self.tscheme
        if self.diffusion and ts.kstage == ts.kforcing:
            self.ope.rhs_diffusion(x, t, dxdt)

    def step(self, t, dt):
        self.tscheme.forward(self.var.dstate, t, dt)
        self.diagnostics(self.var, t)

    def set_psi_from_tracer(self):
        self.ope.invert_vorticity(self.var.dstate)

    # -- diagnostics -------------------------------------------------------------
    def _reduce2(self, entry, *fields):
        """run a two-output masked reduction of libf2d_b200 on state fields, read both scalars"""
        r, s = self.rt, self.var.dstate
        ptrs = [s.rptr(self.var.index(f)) for f in fields]
        entry(r.ptr(self.ope.d_msk), *ptrs, self.nh, s.ny, s.nx, r.ptr(r.out), r.ptr(r.scratch), r.stream)
        return r.read_out(2)

    def diagnostics(self, var, t):
        """'maxspeed' (sets the CFL time step) and 'ke' once -- the flow never changes --
        then mean and rms of the tracer"""
        lib, area, togl = self.rt.lib, self.area, self.mpitools.local_to_global
        if t == 0.:
            ke, maxu = self._reduce2(lib.computekemaxu, 'u', 'v')
            glo = togl([(maxu, 'max'), (ke, 'sum')])
            self.diags.update(maxspeed=glo[0], ke=glo[1]/area, enstrophy=0.)
        total, squares = self._reduce2(lib.computesumandnorm, 'tracer')
        glo = togl([(total, 'sum'), (squares, 'sum')])
        self.diags['mean'] = glo[0]/area
        self.diags['rms'] = np.sqrt(glo[1]/area)
