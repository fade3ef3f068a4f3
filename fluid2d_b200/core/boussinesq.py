"""Boussinesq: vorticity + buoyancy, on the device (reference: core/boussinesq.py).
Same interface: var, ope, tscheme, diags, step, dynamics, add_noslip,
set_psi_from_vorticity, diagnostics, forc / extrastep hooks."""
import numpy as np

from modelbase import adopt, declare_state, user_object, EMBEDDED_FORCING_NOTE
from operators import Operators
from variables import Var
from timescheme import Timescheme
from runtime import rt
from devarray import HostField

FROM_PARAM = ('timestepping', 'forcing', 'forcing_module', 'diffusion', 'Kdiff', 'noslip', 'gravity',
              'customized', 'custom_module', 'additional_tracer', 'isisland', 'myrank')
FROM_GRID = ('nh', 'Lx', 'msk', 'area', 'xr', 'yr', 'mpitools')


class Boussinesq(object):
    def __init__(self, param, grid):
        adopt(self, param, FROM_PARAM)
        adopt(self, grid, FROM_GRID)
        # vorticity and buoyancy are advected; 'banom' is the diagnosed anomaly b - bref
        declare_state(param, grid, ['vorticity', 'psi', 'u', 'v', 'buoyancy', 'banom'],
                      ['vorticity', 'buoyancy'], 'vorticity',
                      more_tracers=getattr(self, 'additional_tracer', ()))
        self.var = Var(param)
        r = rt()
        self.rt = r
        self.ncell = grid.nyl*grid.nxl
        # reference buoyancy: the buoyancy field at construction time (zeros); scripts set it
        # afterwards in place (model.bref[:, :] = buoy), hence a HostField
        self.bref = HostField(self.var.get('buoyancy'))
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
        if self.noslip:
            self.add_noslip(state)
        if self.customized:
            self.extrastep.do(self.var, t, dt)
        ib, ia = self.var.index('buoyancy'), self.var.index('banom')
        if not isinstance(self.bref, HostField):
            self.bref = HostField(self.bref)        # the script replaced the attribute
        lib.set_sum(state.wptr(ia), state.rptr(ib), -1., self.bref.device_ptr(), self.ncell, r.stream)

    def dynamics(self, x, t, dxdt):
        self.ope.rhs_adv(x, t, dxdt)
        # db/dx is a source of vorticity
        self.ope.rhs_torque(x, t, dxdt)
        if self.tscheme.kstage == self.tscheme.kforcing:
            coef = self.tscheme.dtcoef
            if self.forcing:
                assert hasattr(self, 'forc'), self.msg_forcing
                self.forc.add_forcing(x, t, dxdt, coef=coef)
            if self.diffusion:
                self.ope.rhs_diffusion(x, t, dxdt, coef=coef)
        self.ope.invert_vorticity(dxdt, flag='fast')

    def add_noslip(self, x):
        self.ope.rhs_noslip(x, self.source)
        self.ope.invert_vorticity(x, flag='fast', island=self.isisland)

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
            import ctypes
            return ctypes.c_void_p(r.out.data_ptr()+8*k)

        lib.computekemaxu(msk, s.rptr(ix('u')), s.rptr(ix('v')), nh, ny, nx, slot(0), sc, r.stream)
        lib.computesumandnorm(msk, s.rptr(ix('vorticity')), nh, ny, nx, slot(2), sc, r.stream)
        lib.computesumandnorm(msk, s.rptr(ix('buoyancy')), nh, ny, nx, slot(4), sc, r.stream)
        # potential energy: - g * sum(b * y)   (buoyancy is minus density)
        lib.computedotprod(msk, s.rptr(ix('buoyancy')), r.ptr(self.d_yr), nh, ny, nx, slot(6), sc, r.stream)
        ke, maxu, z, z2, b, b2, by = self.mpitools.reduce_device(r, 7, 0x2)   # slot 1 (max speed) is a maximum
        pe = - self.gravity * by
        glo = [maxu, ke, z, z2, pe, b, b2]
        # domain means (maxspeed is a maximum, not a mean)
        maxu, ke, z, z2, pe, b, b2 = [glo[0]]+[v/self.area for v in glo[1:]]
        self.diags.update(maxspeed=maxu, ke=ke, pe=pe, energy=(glo[1]+glo[4])/self.area, vorticity=z, enstrophy=0.5*z2,
                          buoyancy=b, brms=np.sqrt(b2-b**2))
