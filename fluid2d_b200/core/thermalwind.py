"""Thermalwind: vorticity, buoyancy and the along-front velocity V in thermal-wind balance
(reference: core/thermalwind.py, experiments/SymmetricInstability).  'qE' is the diagnosed
Ertel PV, J(V + f0*x, b).  The centred differences carry the reference's in-place linear
extrapolation of the first halo line of b and V (operators.py:330-352)."""
import ctypes

import numpy as np

from modelbase import adopt, declare_state, user_object, EMBEDDED_FORCING_NOTE
from operators import Operators
from variables import Var
from timescheme import Timescheme
from runtime import rt

FROM_PARAM = ('forcing', 'noslip', 'timestepping', 'forcing_module', 'additional_tracer', 'myrank',
              'gravity', 'diffusion', 'Kdiff', 'f0')
FROM_GRID = ('xr', 'yr', 'nh', 'Lx', 'msk', 'area', 'mpitools', 'dx')


class Thermalwind(object):
    def __init__(self, param, grid):
        adopt(self, param, FROM_PARAM)
        adopt(self, grid, FROM_GRID)
        if self.noslip:
            raise NotImplementedError('thermalwind: the reference defines no add_noslip for this model')
        if param.npx*param.npy != 1:
            # rhs_thermalwind / compute_pv fill halos by a local periodic wrap and extrapolate the y
            # boundary rows on every rank (the reference's diffz does so on the first / last rank
            # only, operators.py:330-394): not decomposed over slabs
            raise NotImplementedError('thermalwind: this model runs on one GPU (npx = npy = 1)')
        declare_state(param, grid, ['vorticity', 'psi', 'u', 'v', 'buoyancy', 'V', 'qE'],
                      ['vorticity', 'buoyancy', 'V'], 'vorticity',
                      more_tracers=getattr(self, 'additional_tracer', ()))
        self.var = Var(param)
        r = rt()
        self.rt = r
        self.ny, self.nx = grid.nyl, grid.nxl
        self.ncell = self.ny*self.nx
        self.dy = grid.dy
        self.d_xr = r.to_device(self.xr, dtype=np.float64)
        self.d_yr = r.to_device(self.yr, dtype=np.float64)
        self.ope = Operators(param, grid)
        self.tscheme = Timescheme(param, self.var.dstate)
        if self.forcing:
            if self.forcing_module == 'embedded':
                self.msg_forcing = EMBEDDED_FORCING_NOTE
            else:
                self.forc = user_object(self.forcing_module, 'Forcing', param, grid, 'forcing')
        self.diags = {}

    def step(self, t, dt):
        self.tscheme.set(self.dynamics, self.timestepping)
        self.tscheme.forward(self.var.dstate, t, dt)
        self.set_psi_from_vorticity()
        self.compute_pv()

    def compute_pv(self):
        """qE = J(V + f0*x, b)*msk, halo filled (thermalwind.py:86-98); diffx / diffz leave
        their extrapolated halo lines in b, as in the reference"""
        r, lib = self.rt, self.rt.lib
        s = self.var.dstate
        ix = self.var.index
        nh, ny, nx, n = self.nh, self.ny, self.nx, self.ncell
        iq, iV, ib = ix('qE'), ix('V'), ix('buoyancy')
        X, out = r.ptr(self.ope.work2), r.ptr(self.ope.work)
        lib.set_sum(X, s.rptr(iV), self.f0, r.ptr(self.d_xr), n, r.stream)
        lib.extrapolate_bry(X, nh, ny, nx, 0, r.stream)          # diffx(X)
        lib.extrapolate_bry(s.wptr(ib), nh, ny, nx, 1, r.stream)  # diffz(b)
        lib.extrapolate_bry(X, nh, ny, nx, 1, r.stream)          # diffz(X)
        lib.extrapolate_bry(s.wptr(ib), nh, ny, nx, 0, r.stream)  # diffx(b)
        lib.jacobian(r.ptr(self.ope.d_msk), X, s.rptr(ib), self.dx, self.dy, out, ny, nx, r.stream)
        lib.fill_halo(out, nh, ny, nx, r.stream)
        lib.copy(s.wptr(iq), out, n*8, r.stream)

    def dynamics(self, x, t, dxdt):
        self.ope.rhs_adv(x, t, dxdt)
        self.ope.rhs_thermalwind(x, t, dxdt)
        if self.tscheme.kstage == self.tscheme.kforcing:
            if self.forcing:
                assert hasattr(self, 'forc'), self.msg_forcing
                self.forc.add_forcing(x, t, dxdt)
            if self.diffusion:
                self.ope.rhs_diffusion(x, t, dxdt)
        else:
            self.ope.invert_vorticity(dxdt, flag='fast')

    def set_psi_from_vorticity(self):
        self.ope.invert_vorticity(self.var.dstate)

    def diagnostics(self, var, t):
        r, lib = self.rt, self.rt.lib
        s = var.dstate
        ix = var.index
        nh, ny, nx = self.nh, s.ny, s.nx
        msk, sc = r.ptr(self.ope.d_msk), r.ptr(r.scratch)

        def slot(k):
            return ctypes.c_void_p(r.out.data_ptr()+8*k)

        qneg = r.ptr(self.ope.work)
        lib.negative_part(qneg, s.rptr(ix('qE')), self.ncell, r.stream)
        lib.computekemaxu(msk, s.rptr(ix('u')), s.rptr(ix('v')), nh, ny, nx, slot(0), sc, r.stream)
        lib.computesumandnorm(msk, s.rptr(ix('vorticity')), nh, ny, nx, slot(2), sc, r.stream)
        lib.computesumandnorm(msk, s.rptr(ix('buoyancy')), nh, ny, nx, slot(4), sc, r.stream)
        lib.computesumandnorm(msk, s.rptr(ix('V')), nh, ny, nx, slot(6), sc, r.stream)
        lib.computesumandnorm(msk, s.rptr(ix('qE')), nh, ny, nx, slot(8), sc, r.stream)
        lib.computesumandnorm(msk, qneg, nh, ny, nx, slot(10), sc, r.stream)
        lib.computedotprod(msk, s.rptr(ix('buoyancy')), r.ptr(self.d_yr), nh, ny, nx, slot(12), sc, r.stream)
        ke, maxu, z, z2, b, b2, vm, v2, q, q2, qn, qn2, by = r.read_out(13)
        pe = - self.gravity*by          # buoyancy is minus density
        cst = self.mpitools.local_to_global([
            (maxu, 'max'), (ke, 'sum'), (z, 'sum'), (z2, 'sum'), (pe, 'sum'), (b, 'sum'), (b2, 'sum'),
            (q, 'sum'), (q2, 'sum'), (qn, 'sum'), (qn2, 'sum'), (v2, 'sum')])
        a = self.area
        d = self.diags
        d['maxspeed'] = cst[0]
        d['ke'] = (cst[1]) / a
        d['keV'] = (0.5*cst[11])/a
        d['pe'] = cst[4] / a
        d['energy'] = d['ke'] + d['pe'] + d['keV']
        d['vorticity'] = cst[2] / a
        d['enstrophy'] = 0.5*cst[3] / a
        bm = cst[5] / a
        d['buoyancy'] = bm
        d['brms'] = np.sqrt(cst[6] / a - bm**2)
        pvm = cst[7] / a
        pvneg_mean = cst[9] / a
        d['pv_mean'] = pvm
        d['pv_std'] = np.sqrt(cst[8] / a - pvm**2)
        d['pvneg_mean'] = pvneg_mean
        d['pvneg_std'] = np.sqrt(cst[10] / a - pvneg_mean**2)
