"""Euler: 2-D incompressible Euler / Navier-Stokes in vorticity form, on the device.

Same interface as the reference's core/euler.py -- var, ope, tscheme, timers, diags,
step(t, dt), dynamics(x, t, dxdt), add_noslip(x), set_psi_from_vorticity(),
diagnostics(var, t), forc / extrastep hooks, sponge layer, dye / age tracers.  The
state and the tendency buffers are DeviceState objects; user hooks see them through
host views (devarray.py).
"""
import numpy as np

from modelbase import adopt, declare_state, user_object, EMBEDDED_FORCING_NOTE
from operators import Operators
from variables import Var
from timescheme import Timescheme
from timers import Timers
from runtime import rt

FROM_PARAM = ('timestepping', 'forcing', 'forcing_module', 'diffusion', 'Kdiff', 'noslip', 'spongelayer',
              'customized', 'custom_module', 'additional_tracer', 'var_to_save', 'enforce_momentum')
FROM_GRID = ('nh', 'ny', 'dx', 'Lx', 'msk', 'area', 'xr', 'yr', 'r2', 'x0', 'y0', 'x2', 'y2', 'isisland',
             'mpitools')


class Euler(object):
    def __init__(self, param, grid):
        adopt(self, param, FROM_PARAM)
        adopt(self, grid, FROM_GRID)
        # state: the four dynamical fields, the no-slip source, optional wall diagnostics,
        # then the passive tracers the user asked for
        fields = ['vorticity', 'psi', 'u', 'v', 'source']
        fields += [name for name in ('tauw', 'wshear') if name in self.var_to_save]
        declare_state(param, grid, fields, ['vorticity'], 'vorticity',
                      more_tracers=getattr(self, 'additional_tracer', ()))
        self.var = Var(param)
        self.timers = Timers(param)
        self.ope = Operators(param, grid)
        self.tscheme = Timescheme(param, self.var.dstate)
        self.tscheme.set(self.dynamics, self.timestepping)
        if not (self.forcing or self.customized):
            # Fields whose Runge-Kutta combination can be skipped without changing anything:
            #  - 'source' (and tauw/wshear) have identically zero tendencies;
            #  - the intermediate psi is never read (dynamics reads the tracers, u and v);
            #  - the final u, v are overwritten by the inversions that end the step
            #    (their outermost ring, which computeorthogradient leaves alone, has zero
            #    tendencies); the final psi is kept: it is the first guess of the full solve.
            ix = self.var.index
            tr = [ix(trac) for trac in param.tracer_list]
            self.tscheme.fields_stage = tr+[ix('u'), ix('v')]
            self.tscheme.fields_final = tr+[ix('psi')]
            # at the first stage of RK3_SSP the tendency of a tracer is what rhs_adv leaves (the
            # inversion that follows writes psi, u, v only): the advection kernel writes x + dt*dx0
            if self.timestepping == 'RK3_SSP':
                self.tscheme.adv_hook = self.ope
                self.tscheme.fused_fields = tr
                # ... and at the first two stages the tendencies of u, v are what the inversion
                # that ends dynamics() derives: its orthogradient kernel writes the stage velocities
                self.tscheme.uv_hook = self.ope
                self.tscheme.uv_fields = (ix('u'), ix('v'))
        r = rt()
        self.rt = r
        self.d_xr = r.to_device(self.xr, dtype=np.float64)
        self.d_yr = r.to_device(self.yr, dtype=np.float64)
        self.ncell = grid.nyl*grid.nxl
        if self.forcing:
            if self.forcing_module == 'embedded':
                print(EMBEDDED_FORCING_NOTE)
            else:
                self.forc = user_object(self.forcing_module, 'Forcing', param, grid, 'forcing')
        if self.spongelayer:
            # damping factor towards the east end of the domain: 1 = untouched, 0 = fully damped
            self.spongemsk = (1-(1+np.tanh((self.xr - self.Lx)/0.1))*0.5)
            self.d_spongemsk = r.to_device(self.spongemsk, dtype=np.float64)
        self.diags = {}
        if self.customized:
            self.extrastep = user_object(self.custom_module, 'Step', param, grid, 'customized step')

    def step(self, t, dt):
        r, lib = self.rt, self.rt.lib
        state = self.var.dstate
        # 1/ dynamics
        self.tscheme.forward(state, t, dt)
        # 2/ no-slip source
        if self.noslip:
            self.add_noslip(state)
            isrc = self.var.index('source')
            lib.div_scalar(state.wptr(isrc), dt, self.ncell, r.stream)
            if 'tauw' in self.var_to_save:
                itau = self.var.index('tauw')
                lib.copy(state.wptr(itau), state.rptr(isrc), self.ncell*8, r.stream)
                lib.scale(state.wptr(itau), self.dx, self.ncell, r.stream)
                lib.scale(state.wptr(itau), self.dx, self.ncell, r.stream)
        if self.customized:
            self.extrastep.do(self.var, t, dt)
        # 3/ dye and age tracers (host views: a few points per step)
        if 'dye' in self.var.varname_list:
            i, jp, jm = 1, self.ny//2+3, self.ny//2-3
            dye = self.var.get('dye')
            dye[jp+self.nh, i] = 1
            dye[jm+self.nh, i] = -1
        if 'age' in self.var.varname_list:
            age = self.var.get('age')
            age += dt*self.msk
            age[:, self.nh] = 0.
        # 4/ sponge layer
        if self.spongelayer:
            lib.mul_field(state.wptr(self.var.index('vorticity')), r.ptr(self.d_spongemsk), self.ncell, r.stream)
            for name in ('dye', 'age'):
                if name in self.var.varname_list:
                    lib.mul_field(state.wptr(self.var.index(name)), r.ptr(self.d_spongemsk), self.ncell, r.stream)
        self.set_psi_from_vorticity()

    def dynamics(self, x, t, dxdt):
        """tendencies of the tracers + the streamfunction / velocity they imply"""
        self.timers.tic('rhs_adv')
        self.ope.rhs_adv(x, t, dxdt)
        self.timers.toc('rhs_adv')
        if self.tscheme.kstage == self.tscheme.kforcing:
            if self.forcing:
                self.forc.add_forcing(x, t, dxdt)
            if self.diffusion:
                self.ope.rhs_diffusion(x, t, dxdt)
            if self.diffusion or self.forcing:
                self.ope.invert_vorticity(dxdt, flag='fast')
        else:
            self.timers.tic('invert')
            self.ope.invert_vorticity(dxdt, flag='fast')
            self.timers.toc('invert')

    def add_noslip(self, x):
        self.timers.tic('noslip')
        self.ope.rhs_noslip(x, (self.var.dstate, self.var.index('source')))
        self.timers.toc('noslip')
        self.timers.tic('invert')
        self.ope.invert_vorticity(x, flag='fast', island=self.isisland)
        self.timers.toc('invert')

    def set_psi_from_vorticity(self):
        self.ope.invert_vorticity(self.var.dstate, island=self.isisland)

    def diagnostics(self, var, t):
        """integral diagnostics; 'maxspeed' feeds the cfl criterion (one fused pass)"""
        self.timers.tic('diag')
        r, lib = self.rt, self.rt.lib
        s = var.dstate
        ix = var.index
        lib.diag_euler(r.ptr(self.ope.d_msk), s.rptr(ix('u')), s.rptr(ix('v')), s.rptr(ix('vorticity')),
                       s.rptr(ix('psi')), s.rptr(ix('source')), r.ptr(self.d_xr), r.ptr(self.d_yr),
                       self.nh, s.ny, s.nx, r.ptr(r.out), r.ptr(r.scratch), r.stream)
        names = ('maxspeed', 'ke', 'vorticity', 'enstrophy', 'px', 'py', 'angmom', 'source')
        glo = self.mpitools.reduce_device(r, 8, 0x1)      # slot 0 (max speed) is a maximum
        # domain means, except the maximum speed; enstrophy = half the mean square vorticity
        for k, name in enumerate(names):
            self.diags[name] = glo[k] if k == 0 else glo[k]/self.area
        self.diags['enstrophy'] = 0.5*self.diags['enstrophy']
        self.timers.toc('diag')
