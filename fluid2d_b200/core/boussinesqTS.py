"""BoussinesqTS: Boussinesq model with temperature and salinity (reference:
core/boussinesqTS.py, experiments/doublediffusion).  Vorticity, T and S are advected;
density = -alphaT*T + betaS*S is diagnosed (also on the tendencies, since the torque reads
it there); 'danom' is the departure from the reference density.  Every operation is a call
of libf2d_b200.so on device-resident fields."""
import ctypes

import numpy as np

from modelbase import adopt, declare_state, user_object, EMBEDDED_FORCING_NOTE
from operators import Operators
from variables import Var
from timescheme import Timescheme
from runtime import rt
from devarray import HostField

FROM_PARAM = ('forcing', 'noslip', 'timestepping', 'alphaT', 'betaS', 'diffusion', 'Kdiff', 'myrank',
              'forcing_module', 'gravity', 'isisland', 'customized', 'custom_module', 'additional_tracer')
FROM_GRID = ('xr', 'yr', 'nh', 'Lx', 'msk', 'area', 'mpitools')


class BoussinesqTS(object):
    def __init__(self, param, grid):
        adopt(self, param, FROM_PARAM)
        adopt(self, grid, FROM_GRID)
        declare_state(param, grid, ['vorticity', 'psi', 'u', 'v', 'density', 'danom', 'T', 'S'],
                      ['vorticity', 'T', 'S'], 'vorticity',
                      more_tracers=getattr(self, 'additional_tracer', ()))
        self.varname_list = param.varname_list
        self.var = Var(param)
        r = rt()
        self.rt = r
        self.ncell = grid.nyl*grid.nxl
        # reference density: the density field at construction time (zeros); a script may
        # set it afterwards in place, like Boussinesq.bref
        self.dref = HostField(self.var.get('density'))
        self.source = r.alloc((grid.nyl, grid.nxl))    # (symmetric heap on y-slabs: its halo rows are exchanged)
        self.d_yr = r.to_device(self.yr, dtype=np.float64)
        self.ope = Operators(param, grid)
        self.tscheme = Timescheme(param, self.var.dstate)
        self.tscheme.set(self.dynamics, self.timestepping)
        if self.forcing:
            if self.forcing_module == 'embedded':
                self.msg_forcing = EMBEDDED_FORCING_NOTE
            else:
                self.forc = user_object(self.forcing_module, 'Forcing', param, grid, 'forcing')
        self.diags = {}
        if self.customized:
            self.extrastep = user_object(self.custom_module, 'Step', param, grid, 'customized step')

    def step(self, t, dt):
        r, lib = self.rt, self.rt.lib
        state = self.var.dstate
        self.tscheme.forward(state, t, dt)
        self.set_density()
        if self.noslip:
            self.add_noslip(state)
        if self.customized:
            self.extrastep.do(self.var, t, dt)
        idn, ia = self.var.index('density'), self.var.index('danom')
        if not isinstance(self.dref, HostField):
            self.dref = HostField(self.dref)
        lib.set_sum(state.wptr(ia), state.rptr(idn), -1., self.dref.device_ptr(), self.ncell, r.stream)

    def dynamics(self, x, t, dxdt):
        self.ope.rhs_adv(x, t, dxdt)
        self.eos(dxdt)
        # d(density)/dx is a source of vorticity (minus the buoyancy torque)
        self.ope.rhs_torque_density(x, t, dxdt)
        if self.tscheme.kstage == self.tscheme.kforcing:
            coef = self.tscheme.dtcoef
            if self.forcing:
                assert hasattr(self, 'forc'), self.msg_forcing
                self.forc.add_forcing(x, t, dxdt, coef=coef)
            if self.diffusion:
                self.ope.rhs_diffusion(x, t, dxdt, coef=coef)
        self.eos(dxdt)
        self.ope.invert_vorticity(dxdt, flag='fast')

    def add_noslip(self, x):
        self.ope.rhs_noslip(x, self.source)
        self.ope.invert_vorticity(x, flag='fast', island=self.isisland)

    def eos(self, x):
        """density = (-alphaT)*T + betaS*S on a state-shaped buffer (boussinesqTS.py:120-127),
        with the rounding sequence of that numpy expression"""
        r, lib = self.rt, self.rt.lib
        ix = self.varname_list.index
        idn, it, isalt = ix('density'), ix('T'), ix('S')
        n = self.ncell
        lib.copy(x.wptr(idn), x.rptr(it), n*8, r.stream)
        lib.scale(x.wptr(idn), -self.alphaT, n, r.stream)
        lib.add_scaled(x.wptr(idn), self.betaS, x.rptr(isalt), n, r.stream)

    def set_density(self):
        self.eos(self.var.dstate)

    def set_psi_from_vorticity(self):
        self.ope.invert_vorticity(self.var.dstate, island=self.isisland)

    def diagnostics(self, var, t):
        r, lib = self.rt, self.rt.lib
        s = var.dstate
        ix = var.index
        nh, ny, nx = self.nh, s.ny, s.nx
        msk = r.ptr(self.ope.d_msk)
        sc = r.ptr(r.scratch)

        def slot(k):
            return ctypes.c_void_p(r.out.data_ptr()+8*k)

        lib.computekemaxu(msk, s.rptr(ix('u')), s.rptr(ix('v')), nh, ny, nx, slot(0), sc, r.stream)
        lib.computesumandnorm(msk, s.rptr(ix('vorticity')), nh, ny, nx, slot(2), sc, r.stream)
        lib.computesumandnorm(msk, s.rptr(ix('density')), nh, ny, nx, slot(4), sc, r.stream)
        # potential energy: + g * sum(density * y)
        lib.computedotprod(msk, s.rptr(ix('density')), r.ptr(self.d_yr), nh, ny, nx, slot(6), sc, r.stream)
        ke, maxu, z, z2, d, d2, dy = r.read_out(7)
        pe = + self.gravity * dy
        glo = self.mpitools.local_to_global([(maxu, 'max'), (ke, 'sum'), (z, 'sum'), (z2, 'sum'),
                                             (pe, 'sum'), (d, 'sum'), (d2, 'sum')])
        area = self.area
        self.diags['maxspeed'] = glo[0]
        self.diags['ke'] = glo[1] / area
        self.diags['pe'] = glo[4] / area
        self.diags['energy'] = (glo[1]+glo[4]) / area
        self.diags['vorticity'] = glo[2] / area
        self.diags['enstrophy'] = 0.5*glo[3] / area
        self.diags['density'] = glo[5] / area
        self.diags['drms'] = np.sqrt(glo[6] / area-(glo[5]/area)**2)
