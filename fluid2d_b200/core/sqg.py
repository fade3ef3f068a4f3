"""SQG: surface quasi-geostrophic model (reference: core/sqg.py).  The surface PV is advected
by the flux-form kernels; the streamfunction comes from a spectral inversion (fourier.py,
cuFFT) instead of the multigrid.  Doubly periodic, one GPU."""
import ctypes

from modelbase import adopt, declare_state, user_object
from operators import Operators
from variables import Var
from timescheme import Timescheme
from runtime import rt

FROM_PARAM = ('forcing', 'diffusion', 'Kdiff', 'timestepping', 'ageostrophic', 'forcing_module', 'geometry')
FROM_GRID = ('yr', 'nh', 'msk', 'area', 'mpitools')


class SQG(object):
    def __init__(self, param, grid):
        adopt(self, param, FROM_PARAM)
        adopt(self, grid, FROM_GRID)
        assert grid.geometry == "perio", "SQG imposes a biperiodic domain"
        assert param.ageostrophic == False, "Ageostrophic velocity not yet tested"  # noqa: E712
        if param.npx*param.npy != 1:
            raise ValueError('SQG does not support several subdomains (because of the fft)')
        declare_state(param, grid, ['pv', 'psi', 'u', 'v', 'vorticity'], ['pv'], 'pv')
        self.var = Var(param)
        self.rt = rt()
        ix = self.var.index
        self.ipv, self.ivor, self.ipsi = ix('pv'), ix('vorticity'), ix('psi')
        param.sqgoperator = True     # Operators builds the Fourier inversion
        self.ope = Operators(param, grid)
        self.tscheme = Timescheme(param, self.var.dstate)
        self.dx0 = self.tscheme.dx0
        self.kt = 0
        if self.forcing and self.forcing_module != 'embedded':
            self.forc = user_object(self.forcing_module, 'Forcing', param, grid, 'forcing')
        self.diags = {}
        self.tscheme.set(self.dynamics, self.timestepping)

    def step(self, t, dt):
        self.dt = dt
        self.tscheme.forward(self.var.dstate, t, dt)
        self.set_psi_from_pv()

    def dynamics(self, x, t, dxdt):
        r, lib = self.rt, self.rt.lib
        lib.zero(dxdt.all_ptr(True), dxdt.size*8, r.stream)
        self.ope.rhs_adv(x, t, dxdt)
        if self.tscheme.kstage == self.tscheme.kforcing:
            if self.forcing:
                self.forc.add_forcing(x, t, dxdt)
            if self.diffusion:
                self.ope.rhs_diffusion(x, t, dxdt)
        else:
            self.ope.fourier_invert_vorticity(dxdt, flag='fast')

    def set_psi_from_pv(self):
        self.ope.fourier_invert_vorticity(self.var.dstate, flag='full')

    def diagnostics(self, var, t):
        r, lib = self.rt, self.rt.lib
        s = var.dstate
        nh, ny, nx = self.nh, s.ny, s.nx
        msk, sc = r.ptr(self.ope.d_msk), r.ptr(r.scratch)

        def slot(k):
            return ctypes.c_void_p(r.out.data_ptr()+8*k)

        lib.computekemaxu(msk, s.rptr(var.index('u')), s.rptr(var.index('v')), nh, ny, nx, slot(0), sc, r.stream)
        lib.computesumandnorm(msk, s.rptr(self.ipv), nh, ny, nx, slot(2), sc, r.stream)
        ke, maxu, z, z2 = r.read_out(4)
        cst = self.mpitools.local_to_global([(maxu, 'max'), (ke, 'sum'), (z, 'sum'), (z2, 'sum')])
        self.diags['maxspeed'] = cst[0]
        self.diags['ke'] = cst[1] / self.area
        self.diags['pv'] = cst[2] / self.area
        self.diags['pv2'] = 0.5*cst[3] / self.area
