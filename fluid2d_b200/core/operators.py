"""Operators: the right-hand-side terms and the vorticity inversion, on the device.

Same interface as the reference's core/operators.py -- rhs_adv(x, t, dxdt),
rhs_diffusion(x, t, dxdt, coef), rhs_torque(x, t, dxdt), rhs_noslip(x, source),
invert_vorticity(x, flag, island), fill_halo(a), cst, gmg, mskp, mskbc, bcarea -- where
x and dxdt are DeviceState objects (the model state or a tendency buffer) instead of
numpy arrays.  Every method enqueues kernels of libf2d_b200.so on the current CUDA
stream and returns; only the 'full' inversion synchronises (its iteration count depends
on residual norms, hierarchy.py:169).
"""
import ctypes

import numpy as np
import torch

from param import Param
from runtime import rt
from gmg.hierarchy import Gmg
from devarray import TrackedArray


class Operators(Param):
    def __init__(self, param, grid):
        self.list_param = ['varname_list', 'tracer_list', 'whosetspsi', 'mpi', 'npx', 'npy', 'nh',
                           'gravity', 'f0', 'beta', 'Rd', 'qgoperator', 'order', 'Kdiff', 'diffusion',
                           'enforce_momentum', 'isisland', 'aparab', 'flux_splitting_method',
                           'hydroepsilon', 'myrank', 'geometry', 'sqgoperator']
        param.copy(self, self.list_param)
        self.list_grid = ['msk', 'nxl', 'nyl', 'dx', 'dy', 'bcarea', 'mpitools', 'msknoslip', 'mskbc',
                          'domain_integration', 'nh', 'xr0', 'yr0', 'i0', 'j0', 'area']
        grid.copy(self, self.list_grid)
        self.first_time = True
        self.grid = grid
        r = rt()
        self.rt = r
        self.lib = r.lib
        ny, nx = self.nyl, self.nxl
        self.shape = (ny, nx)
        self.ncell = ny*nx

        # device copies of the masks (set-up: integer logic on the host, one upload)
        msk = np.ascontiguousarray(self.msk, dtype=np.int8)
        self.d_msk = r.to_device(msk)
        # NULL mask selects the all-fluid advection kernels (no mask traffic)
        self.all_fluid = bool(msk.all())
        # corner mask: a corner is fluid iff its 4 cells are (operators.py:59-67)
        mskp = np.zeros((ny, nx), dtype=np.int8)
        mskp[:-1, :-1] = msk[:-1, :-1] & msk[:-1, 1:] & msk[1:, :-1] & msk[1:, 1:]
        self.mskp = mskp
        self.d_mskp = r.to_device(mskp)
        self.d_msknoslip = r.to_device(np.ascontiguousarray(self.msknoslip, dtype=np.int8))

        # work arrays of the inversion
        # y-slab decomposition: kernels store their x images only (fill mode 2) and the
        # y halo rows are pushed to the neighbours by f2d_comm_exchange_y
        self.comm = r.comm
        self.fillmode = 2 if r.comm is not None else 1
        self.work = r.alloc((ny, nx))
        self.work2 = r.alloc((ny, nx))

        pp = {'np': 1, 'mp': 1, 'nh': param.nh, 'n': nx-2*self.nh, 'm': ny-2*self.nh,
              'omega': 8./9., 'dx': grid.dx, 'dy': grid.dy, 'hydroepsilon': param.hydroepsilon,
              'relaxation': param.relaxation}
        if hasattr(self, 'qgoperator'):
            pp['qgoperator'] = True
            pp['Rd'] = self.Rd
        if self.myrank == 0:
            print('-'*50)
            print(' Multigrid hierarchy (device)')
            print('-'*50)
        self.gmg = Gmg(pp, mskp.astype(np.float64), comm=r.comm)
        if self.myrank == 0:
            for g in self.gmg.grid:
                print('Level %2i: %5ix%5i' % (g.lev, g.n, g.m))

        if hasattr(self, 'sqgoperator'):
            from fourier import Fourier
            self.fourier = Fourier(param, grid, r.device)

        grid.fill_halo = self.fill_halo
        self.set_boundary_msk()

        self.cst = np.zeros(5,)
        self.cst[0] = grid.dx
        self.cst[1] = grid.dy
        self.cst[2] = 0.05
        self.cst[3] = 0   # umax: updated at each time step by Fluid2d.set_dt
        self.upwind = self.order % 2 == 1
        self.cst[4] = self.aparab if self.upwind else 0
        list_fs_method = ['minmax', 'parabolic']
        if self.flux_splitting_method in list_fs_method:
            self.fs_method = list_fs_method.index(self.flux_splitting_method)
        else:
            print('Warning: %s does not exist' % self.flux_splitting_method)
            print('replaced with the default: parabolic')
            self.fs_method = list_fs_method.index('parabolic')

        if type(self.Kdiff) != dict:
            K = self.Kdiff
            self.Kdiff = {}
            for trac in self.tracer_list:
                self.Kdiff[trac] = K
        if self.diffusion:
            print('diffusion coefficients')
            print('  => ', self.Kdiff)
        self.rhsp = None   # island fields, set by Fluid2d (host arrays)
        self.psi = None
        self._d_rhsp = None
        self._d_psi_island = None
        self.last_solve = (0, 0.)
        self.fuse = None     # request of the time scheme: stage update fused into the advection kernel
        self.fused = []
        self.fuse_uv = None  # request of the time scheme: stage velocities written by the next inversion
        self.fused_uv = False

    # ------------------------------------------------------------------
    def set_boundary_msk(self):
        """mask of the fluid cells that touch a no-slip wall (operators.py:157-186)"""
        msk = self.msknoslip
        z = (np.roll(msk, -1, axis=1)+np.roll(msk, -1, axis=0)
             + np.roll(msk, +1, axis=1)+np.roll(msk, +1, axis=0)-4*msk)
        z = z*msk
        mskbc = self.msk*0
        mskbc[z < 0] = 1
        mskbc *= self.msknoslip
        self.fill_halo(mskbc)
        self.mskbc = mskbc
        self.d_mskbc = self.rt.to_device(np.ascontiguousarray(mskbc, dtype=np.int8))
        self.bcarea = self.domain_integration(self.mskbc)
        self.x2bc = self.domain_integration((self.xr0)**2 * self.mskbc*self.msknoslip)
        self.y2bc = self.domain_integration((self.yr0)**2 * self.mskbc*self.msknoslip)
        self.d_xr0 = self.rt.to_device(self.xr0, dtype=np.float64)
        self.d_yr0 = self.rt.to_device(self.yr0, dtype=np.float64)

    def ix(self, name):
        return self.varname_list.index(name)

    # ------------------------------------------------------------------
    def fill_halo(self, x):
        """periodic halo fill (fortran_multigrid.f90 fillhalo through the device).
        Accepts a host array (round trip through HBM; set-up convenience for user
        scripts, e.g. grid.fill_halo(noise)), a state view, or a device tensor."""
        r = self.rt
        if isinstance(x, torch.Tensor):
            self._fill(r.ptr(x), x.shape[0], x.shape[1])
            return
        a = np.asarray(x)
        if self.comm is not None:
            # slabs: through a scratch field of the symmetric heap (exchange with the neighbours)
            self.work2.copy_(torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)))
            self._fill(r.ptr(self.work2), a.shape[0], a.shape[1])
            x[...] = self.work2.cpu().numpy().astype(a.dtype)
            return
        if a.dtype == np.int8:
            d = r.to_device(a)
            self.lib.fill_halo_i8(r.ptr(d), self.nh, a.shape[0], a.shape[1], r.stream)
        else:
            d = r.to_device(a, dtype=np.float64)
            self.lib.fill_halo(r.ptr(d), self.nh, a.shape[0], a.shape[1], r.stream)
        x[...] = d.cpu().numpy()

    def _fill(self, ptr, ny, nx):
        """full halo fill of a device field: local wrap, or x images + neighbour exchange"""
        r = self.rt
        if self.comm is None:
            self.lib.fill_halo(ptr, self.nh, ny, nx, r.stream)
        else:
            # the field was written whole (halo rows included) by this rank; a faster
            # neighbour must not push its rows before that write has completed here
            self.lib.comm_barrier(self.comm, r.stream)
            self.lib.fill_halo_x(ptr, self.nh, ny, nx, r.stream)
            self.lib.comm_exchange_y(self.comm, ptr, self.nh, ny, nx, r.stream)

    def _barrier(self):
        """device-side barrier of all ranks (no-op on one GPU): needed before an exchange of
        a field whose halo rows this rank has just overwritten (memset, upload)"""
        if self.comm is not None:
            self.lib.comm_barrier(self.comm, self.rt.stream)

    def _xch(self, ptr):
        """y halo rows of a field whose producer used fill mode 2 (no-op on one GPU)"""
        if self.comm is not None:
            self.lib.comm_exchange_y(self.comm, ptr, self.nh, self.nyl, self.nxl, self.rt.stream)

    # ------------------------------------------------------------------
    def rhs_adv(self, x, t, dxdt):
        """dxdt[tracer] = -div(u tracer) for every tracer, halo filled (operators.py:214-236): one
        launch for all the tracers of the model (they share the velocity tiles).  When the time
        scheme has asked for it (self.fuse = (xout, xbase, coef, fields), Timescheme.RK3_SSP), the
        same kernel also writes the stage state xout[k] = xbase[k] + coef*dxdt[k] of the tracers
        in `fields`; self.fused says which ones it did."""
        r, lib = self.rt, self.lib
        iu, iv = self.ix('u'), self.ix('v')
        cst = (ctypes.c_double*5)(*self.cst)
        msk = None if self.all_fluid else r.ptr(self.d_msk)
        tracers = [self.ix(trac) for trac in self.tracer_list]
        self.fused = []
        groups = [(tracers, None)]
        if self.fuse is not None and self.comm is None:
            xout, xbase, coef, fields = self.fuse
            yes = [k for k in tracers if k in fields]
            no = [k for k in tracers if k not in fields]
            groups = [(g, f) for g, f in ((yes, (xout, xbase, coef)), (no, None)) if g]
            self.fused = yes
        pu, pv = x.rptr(iu), x.rptr(iv)
        for group, fuse in groups:
            n = len(group)
            arr = ctypes.c_void_p*n
            q = arr(*[x.rptr(k).value for k in group])
            dq = arr(*[dxdt.wptr(k).value for k in group])
            xb = xo = None
            coef = 0.
            if fuse is not None:
                xb = arr(*[fuse[1].rptr(k).value for k in group])
                xo = arr(*[fuse[0].wptr(k).value for k in group])
                coef = fuse[2]
            lib.adv_multi(msk, q, dq, n, pu, pv, cst, self.nh, 1 if self.upwind else 0, self.fs_method,
                          self.order, xb, xo, coef, self.nyl, self.nxl, self.fillmode, r.stream)
        for ik in tracers:
            self._xch(dxdt.wptr(ik))

    def rhs_diffusion(self, x, t, dxdt, coef=1.):
        r, lib = self.rt, self.lib
        for trac in self.tracer_list:
            ik = self.ix(trac)
            lib.add_diffusion(r.ptr(self.d_msk), x.rptr(ik), self.dx, self.nh, coef*self.Kdiff[trac],
                              dxdt.wptr(ik), self.nyl, self.nxl, self.fillmode, r.stream)
            self._xch(dxdt.wptr(ik))

    def rhs_torque(self, x, t, dxdt):
        r, lib = self.rt, self.lib
        ib, iw = self.ix('buoyancy'), self.ix('vorticity')
        lib.add_torque(r.ptr(self.d_msk), x.rptr(ib), self.dx, self.nh, self.gravity, dxdt.wptr(iw),
                       self.nyl, self.nxl, 1, self.fillmode, r.stream)
        self._xch(dxdt.wptr(iw))

    def rhs_torque_density(self, x, t, dxdt):
        """the torque of the BoussinesqTS model reads the density: same kernel with -g and
        without the mask product on the tendency (operators.py:316-328)"""
        r, lib = self.rt, self.lib
        ib, iw = self.ix('density'), self.ix('vorticity')
        lib.add_torque(r.ptr(self.d_msk), x.rptr(ib), self.dx, self.nh, -self.gravity, dxdt.wptr(iw),
                       self.nyl, self.nxl, 0, self.fillmode, r.stream)
        self._xch(dxdt.wptr(iw))

    def rhs_thermalwind(self, x, t, dxdt):
        """thermal-wind model: g*db/dx - f0*dV/dz on the vorticity, -f0*u on V
        (operators.py:357-394).  diffx / diffz first extrapolate the first halo line of b
        and V linearly IN PLACE (operators.py:330-352): the stage state keeps those values."""
        r, lib = self.rt, self.lib
        nh, ny, nx, n = self.nh, self.nyl, self.nxl, self.ncell
        iu, ib, iw, iV = self.ix('u'), self.ix('buoyancy'), self.ix('vorticity'), self.ix('V')
        msk, y = r.ptr(self.d_msk), r.ptr(self.work)
        lib.extrapolate_bry(x.wptr(ib), nh, ny, nx, 0, r.stream)
        lib.extrapolate_bry(x.wptr(iV), nh, ny, nx, 1, r.stream)
        lib.tw_torque(msk, x.rptr(ib), x.rptr(iV), self.dx, self.dy, self.gravity, self.f0, y, ny, nx, r.stream)
        lib.fill_halo(y, nh, ny, nx, r.stream)
        lib.add_scaled(dxdt.wptr(iw), 1., y, n, r.stream)
        lib.tw_coriolis(msk, x.rptr(iu), self.f0, y, ny, nx, r.stream)
        lib.fill_halo(y, nh, ny, nx, r.stream)
        lib.add_scaled(dxdt.wptr(iV), 1., y, n, r.stream)

    def rhs_noslip(self, x, source):
        """vorticity source along the walls that cancels the tangential velocity
        (operators.py:245-290); `source` is (DeviceState, field index) or a device tensor"""
        r, lib = self.rt, self.lib
        ip, iw = self.ix('psi'), self.ix(self.whosetspsi)
        ny, nx, n = self.nyl, self.nxl, self.ncell
        src = source[0].wptr(source[1]) if isinstance(source, tuple) else r.ptr(source)
        work = r.ptr(self.work)
        lib.cornertocell(x.rptr(ip), work, ny, nx, r.stream)
        lib.noslip_source(r.ptr(self.d_msknoslip), x.rptr(ip), work, self.dx, self.dy, self.nh, ny, nx, r.stream)
        lib.copy(src, work, n*8, r.stream)
        # zero net source: subtract its mean over the boundary cells
        lib.domain_sum(src, self.nh, ny, nx, r.ptr(r.out), r.ptr(r.scratch), r.stream)
        if self.comm is not None:
            # y-slabs: the integral is over the whole domain (bcarea is the global one)
            lib.comm_allreduce(self.comm, r.ptr(r.out), 1, 0, r.stream)
        lib.sub_devscalar_mask(src, r.ptr(r.out), float(self.bcarea[0]), r.ptr(self.d_mskbc), n, r.stream)
        if self.enforce_momentum:
            lib.computedotprod(r.ptr(self.d_msk), src, r.ptr(self.d_xr0), self.nh, ny, nx,
                               r.ptr(r.out), r.ptr(r.scratch), r.stream)
            lib.computedotprod(r.ptr(self.d_msk), src, r.ptr(self.d_yr0), self.nh, ny, nx,
                               ctypes.c_void_p(r.out.data_ptr()+8), r.ptr(r.scratch), r.stream)
            px, py = r.read_out(2)
            cst = self.mpitools.local_to_global([(px, 'sum'), (py, 'sum')])
            px, py = cst[0]/self.x2bc, cst[1]/self.y2bc
            lib.sub_lin2_mask(src, float(px), r.ptr(self.d_xr0), float(py), r.ptr(self.d_yr0),
                              r.ptr(self.d_mskbc), n, r.stream)
        self._fill(src, ny, nx)
        lib.add_scaled(x.wptr(iw), -1., src, n, r.stream)

    # ------------------------------------------------------------------
    def _island_fields(self):
        if self._d_rhsp is None:
            self._d_rhsp = self.rt.to_device(self.rhsp, dtype=np.float64)
            self._d_psi_island = self.rt.to_device(self.psi, dtype=np.float64)
        return self.rt.ptr(self._d_rhsp), self.rt.ptr(self._d_psi_island)

    def fourier_invert_vorticity(self, x, flag='full'):
        """SQG: psi (and the diagnosed vorticity) from the surface pv by a spectral inversion
        (cuFFT, fourier.py), then (u, v) from psi (operators.py:396-419)"""
        r, lib = self.rt, self.lib
        iu, iv, ip, ivor, ipv = self.ix('u'), self.ix('v'), self.ix('psi'), self.ix('vorticity'), self.ix('pv')
        ny, nx = self.nyl, self.nxl
        x.rptr(ipv)                       # host edits of pv reach the device first
        ppsi, pvor = x.wptr(ip), x.wptr(ivor)
        self.fourier.invert(x.dev[ipv], x.dev[ip], x.dev[ivor])
        lib.fill_halo(ppsi, self.nh, ny, nx, r.stream)
        lib.fill_halo(pvor, self.nh, ny, nx, r.stream)
        self.first_time = False
        lib.orthogradient(r.ptr(self.d_msk), ppsi, self.dx, self.dy, self.nh, x.wptr(iu), x.wptr(iv),
                          ny, nx, r.stream)

    def invert_vorticity(self, x, flag='full', island=False):
        """psi from x[whosetspsi] by multigrid, then (u, v) from psi (operators.py:421-498).
        flag 'fast' = two V-cycles from the current psi; 'full' = F-cycles to 1e-11."""
        r, lib = self.rt, self.lib
        iu, iv, ip, iw = self.ix('u'), self.ix('v'), self.ix('psi'), self.ix(self.whosetspsi)
        rhsp, psi_island = self._island_fields() if island else (None, None)
        full = 0 if flag == 'fast' else 1
        nite, res = ctypes.c_int(), ctypes.c_double()
        if full and self.first_time and self.myrank == 0:
            print('-'*50)
            print(' Convergence of the vorticity inversion')
            print('-'*50)
        # all-fluid domain without island: the mask-free orthogradient (no mask traffic)
        nomask = self.all_fluid and not island and self.nxl % 2 == 0
        fu = getattr(self, 'fuse_uv', None)
        if fu is not None and x is fu[0]:
            # Timescheme.RK3_SSP asked for the stage velocities: this inversion's orthogradient
            # kernel writes out[u] = base[u] + coef*([extra[u] +] x[u]) (v alike) as well.  (On
            # slabs too: no exchange separates the orthogradient from the combination it replaces,
            # and the kernel writes every cell of the local array exactly as f2d_ts_xpay would.)
            _, base, extra, out, coef = fu
            lib.mg_set_uv_stage(self.gmg.h, base.rptr(iu), base.rptr(iv),
                                extra.rptr(iu) if extra is not None else None,
                                extra.rptr(iv) if extra is not None else None,
                                out.wptr(iu), out.wptr(iv), coef)
            self.fused_uv = True
            self.fuse_uv = None
        lib.invert_vorticity(self.gmg.h, None if nomask else r.ptr(self.d_msk),
                             None if nomask else r.ptr(self.d_mskp), x.rptr(iw), x.wptr(ip),
                             x.wptr(iu), x.wptr(iv), r.ptr(self.work), rhsp, psi_island, full,
                             1 if self.geometry == 'perio' else 0, float(np.asarray(self.area).ravel()[0]),
                             self.dx, self.dy, self.nh, ctypes.byref(nite), ctypes.byref(res),
                             r.ptr(r.scratch), r.stream)
        if full:
            if self.first_time and self.myrank == 0:
                print(' ite = %i / res = %.2e' % (nite.value, res.value))
            self.last_solve = (nite.value, res.value)
        self.first_time = False
