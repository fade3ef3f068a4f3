"""Fluxes: diagnostic of the reversible / irreversible parts of the advective fluxes.

Same interface as the reference's core/fluxes.py -- Fluxes(param, grid, ope),
.diag_fluxes(x, t, dt), .flx (the stack [rev_x_*, rev_y_*, ..., irr_x_*, irr_y_*] named by
.fullflx_list), .flx_list, .nvarstate, .advection / .rhs_adv as the right-hand side of its
own Timescheme -- with the extended state (model state + cell-centred velocities uc, vc +
two face-flux slots per advected quantity), the work copies and the flux stack resident in
HBM.  The advection kernel is the model's (f2d_adv_upwind / f2d_adv_centered) with its
face-flux outputs switched on (core/fortran_fluxes.f90).

One diagnostic = one step forward from x, one step backward in time from the same x with
every velocity-like field reversed; the fluxes accumulated by the time scheme in the two
integrations are then half-summed / half-differenced (fluxes.py:100-177).
"""
import copy
import ctypes

from devarray import DeviceState, TrackedArray
from runtime import rt
from timescheme import Timescheme


class Fluxes(object):
    def __init__(self, param, grid, ope):
        self.list_param = ['timestepping', 'varname_list', 'tracer_list', 'order', 'aparab',
                           'sizevar', 'flux_splitting_method', 'modelname']
        p = copy.deepcopy(param)
        # two more advected variables: the velocities at cell centres, whose fluxes
        # measure the dissipation of kinetic energy
        newvariables = ['uc', 'vc']
        p.tracer_list = list(p.tracer_list)+newvariables
        p.varname_list = list(p.varname_list)+newvariables
        self.flx_list = ['flx_%s_%s' % (v, d) for v in p.tracer_list for d in ['x', 'y']]
        self.nvarstate = len(p.varname_list)
        p.varname_list = p.varname_list+self.flx_list
        self.fullflx_list = ['%s_%s_%s' % (r, d, v) for r in ['rev', 'irr']
                             for v in p.tracer_list for d in ['x', 'y']]
        p.copy(self, self.list_param)
        if self.modelname not in ('euler', 'boussinesq', 'thermalwind'):
            raise NotImplementedError('diag_fluxes: model %s (the reference does euler, boussinesq and '
                                      'thermalwind, fluxes.py:9-13)' % self.modelname)
        self.list_grid = ['nh', 'dx', 'dy', 'msk']
        grid.copy(self, self.list_grid)
        self.ope = ope
        self.rt = rt()
        ny, nx = self.sizevar
        self.ny, self.nx, self.fieldsize = ny, nx, ny*nx
        nall = len(self.varname_list)
        self.x = DeviceState(nall, ny, nx)
        self.xe = DeviceState(self.nvarstate, ny, nx)     # model state + (uc, vc)
        self.xwork = DeviceState(nall, ny, nx)
        self._flx = DeviceState(len(self.fullflx_list), ny, nx)
        self.upwind = self.order % 2 == 1
        self.tscheme = Timescheme(p, self.x)
        self.tscheme.set(self.advection, p.timestepping)
        self.fs_method = ope.fs_method

    @property
    def flx(self):
        """host view [len(fullflx_list), ny, nx] of the flux stack (refreshed from the device)"""
        return self._flx.host_view(None)

    def ix(self, name):
        return self.varname_list.index(name)

    def _copy_fields(self, dst, k_dst, src, k_src, count):
        r = self.rt
        for k in range(1, count):
            src.rptr(k_src+k)
            dst.wptr(k_dst+k)
        r.lib.copy(dst.wptr(k_dst), src.rptr(k_src), count*self.fieldsize*8, r.stream)

    def _zero_fluxes(self, s):
        r = self.rt
        nvs, nall = self.nvarstate, len(self.varname_list)
        for k in range(nvs+1, nall):
            s.wptr(k)
        r.lib.zero(s.wptr(nvs), (nall-nvs)*self.fieldsize*8, r.stream)

    def diag_fluxes(self, x, t, dt):
        """x: the model state -- its DeviceState, or the host view Var.state hands out"""
        r, lib = self.rt, self.rt.lib
        if isinstance(x, TrackedArray) and x._own is not None:
            x = x._own[0]
        if not isinstance(x, DeviceState):
            tmp = DeviceState(self.nvarstate-2, self.ny, self.nx)
            tmp.upload_all_from(x)
            x = tmp
        nvs, fs = self.nvarstate, self.fieldsize
        iu, iv, ip, iw = self.ix('u'), self.ix('v'), self.ix('psi'), self.ix('vorticity')
        self._copy_fields(self.xe, 0, x, 0, nvs-2)
        lib.flx_cellvel(x.rptr(iu), x.rptr(iv), self.xe.wptr(nvs-2), self.xe.wptr(nvs-1),
                        self.nh, self.ny, self.nx, self.ope.fillmode, r.stream)
        self.ope._xch(self.xe.wptr(nvs-2))      # (y-slabs: halo rows from the neighbours)
        self.ope._xch(self.xe.wptr(nvs-1))
        # forward step
        self._copy_fields(self.x, 0, self.xe, 0, nvs)
        self._zero_fluxes(self.x)
        self.tscheme.forward(self.x, t, dt)
        self._copy_fields(self.xwork, 0, self.x, 0, len(self.varname_list))
        # backward step from the same state, velocities reversed
        self._copy_fields(self.x, 0, self.xe, 0, nvs)
        for k in (iu, iv, ip, iw, nvs-2, nvs-1):
            lib.scale(self.x.wptr(k), -1., fs, r.stream)
        self._zero_fluxes(self.x)
        self.tscheme.forward(self.x, t+dt, -dt)
        # tendencies: divide by dt; vorticity, uc, vc are odd under time reversal
        cff = 0.5/dt
        nflx = len(self.flx_list)
        for k in range(nflx):
            ell = nvs+k
            sign = -1. if (k < 2) or (k >= nflx-4) else 1.
            lib.flx_split(self._flx.wptr(k), self._flx.wptr(nflx+k), self.xwork.rptr(ell), self.x.rptr(ell),
                          cff, sign, fs, r.stream)

    def advection(self, x, t, dxdt):
        self.rhs_adv(x, t, dxdt)
        if self.modelname == 'boussinesq':
            self.ope.rhs_torque(x, t, dxdt)
        elif self.modelname == 'thermalwind':
            self.ope.rhs_thermalwind(x, t, dxdt)
        self.ope.invert_vorticity(dxdt, flag='fast')

    def rhs_adv(self, x, t, dxdt):
        """-div(u tracer) and the two face fluxes of every advected quantity, halos filled
        (fluxes.py:179-213)"""
        r, lib, ope = self.rt, self.rt.lib, self.ope
        iu, iv = self.ix('u'), self.ix('v')
        cst = (ctypes.c_double*5)(self.dx, self.dy, 0.05, ope.cst[3], self.aparab)
        adv = lib.adv_upwind if self.upwind else lib.adv_centered
        msk = None if ope.all_fluid else r.ptr(ope.d_msk)
        for itrac, trac in enumerate(self.tracer_list):
            ik = self.ix(trac)
            ifx = self.nvarstate+itrac*2
            xf, yf = dxdt.wptr(ifx), dxdt.wptr(ifx+1)
            adv(msk, x.rptr(ik), dxdt.wptr(ik), x.rptr(iu), x.rptr(iv), xf, yf, cst,
                self.nh, self.fs_method, self.order, self.ny, self.nx, ope.fillmode, r.stream)
            ope._xch(dxdt.wptr(ik))
            ope._fill(xf, self.ny, self.nx)
            ope._fill(yf, self.ny, self.nx)
