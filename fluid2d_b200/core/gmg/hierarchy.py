"""Gmg: Python face of the device multigrid (reference: core/gmg/hierarchy.py:21-218
and core/gmg/level.py).  The hierarchy -- masks, Galerkin matrices, work arrays, the
V/F cycles as CUDA graphs -- lives inside libf2d_b200.so (f2d_mg_*); this class keeps
the reference's entry points: Gmg(param, mskp), nlevs, twoVcycle(x, b, param),
solve(x, b, param) -> (nite, res), Vcycle(lev), Fcycle(lev).
"""
import ctypes

import numpy as np
import torch

from runtime import rt


class LevelInfo(object):
    """read-only window on one level (level.Grid: n, m, nv, mv, msk, A)"""

    def __init__(self, gmg, lev):
        self._g, self.lev = gmg, lev
        ny, nx = ctypes.c_int(), ctypes.c_int()
        gmg.lib.mg_level_shape(gmg.h, lev, ctypes.byref(ny), ctypes.byref(nx))
        self.mv, self.nv = ny.value, nx.value
        self.m, self.n = self.mv-2*gmg.nh, self.nv-2*gmg.nh
        self.matrix_mode = gmg.lib.mg_level_matrix_mode(gmg.h, lev)

    def _fetch(self, which, count, dtype):
        r = rt()
        buf = torch.empty(count, dtype=dtype, device=r.device)
        nbytes = count*buf.element_size()
        self._g.lib.copy(r.ptr(buf), self._g.lib.mg_level_ptr(self._g.h, self.lev, which), nbytes, r.stream)
        torch.cuda.current_stream().synchronize()
        return buf.cpu().numpy()

    @property
    def msk(self):
        return self._fetch(0, self.mv*self.nv, torch.int8).reshape(self.mv, self.nv)

    @property
    def A(self):
        """[mv, nv, 5] like the reference's A[:, :, :5] (stored as 5 planes on the device)"""
        a = self._fetch(1, 5*self.mv*self.nv, torch.float64).reshape(5, self.mv, self.nv)
        return np.ascontiguousarray(np.moveaxis(a, 0, 2))


class Gmg(object):
    def __init__(self, param, mskp, comm=None):
        """param: dict with n, m (local interior sizes), nh, dx, dy, omega, hydroepsilon
        [, qgoperator, Rd]; mskp: corner mask, 0/1 (numpy or device float64 tensor)"""
        r = rt()
        self.lib = r.lib
        self.nh = param['nh']
        if param.get('np', 1)*param.get('mp', 1) != 1:
            raise NotImplementedError('Gmg: one subdomain per handle (multi-GPU goes through slabs)')
        ny, nx = param['m']+2*self.nh, param['n']+2*self.nh
        if isinstance(mskp, np.ndarray):
            mskp = r.to_device(mskp, dtype=np.float64)
        assert tuple(mskp.shape) == (ny, nx)
        Rd = float(param['Rd']) if param.get('qgoperator', False) else 0.
        self.h = ctypes.c_void_p()
        args = (r.ptr(mskp), ny, nx, float(param['dx']), float(param['dy']),
                float(param.get('omega', 8./9.)), float(param.get('hydroepsilon', 1.)), Rd, r.stream)
        if comm is None:
            self.lib.mg_create(ctypes.byref(self.h), *args)
        else:
            # y-slabs: (ny, nx) is the local slab; coarse levels are gathered and replicated
            self.lib.mg_create_slab(ctypes.byref(self.h), comm, *args)
        # line relaxation (level.py:153-163): asked for, or implied by flat cells
        self.relaxation = param.get('relaxation', 'default')
        if float(param.get('hydroepsilon', 1.))*float(param['dy'])/float(param['dx']) <= .2:
            self.relaxation = 'tridiagonal'
        if self.relaxation == 'tridiagonal':
            if comm is not None:
                raise ValueError('Small aspect ratio experiment requires param.npy = 1')
            self.lib.mg_set_relaxation(self.h, 1)
        self.slab_levels = self.lib.mg_slab_levels(self.h)
        self.nlevs = self.lib.mg_nlevels(self.h)
        self.nglo, self.mglo = param['n'], param['m']
        self.grid = [LevelInfo(self, lev) for lev in range(self.nlevs)]
        self.npre, self.npost, self.nvcyc, self.ndeepest = 1, 1, 1, 16

    def __del__(self):
        try:
            if self.h:
                self.lib.mg_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def twoVcycle(self, x, b, param=None):
        """x, b: device addresses (ctypes.c_void_p) of psi (first guess / result) and rhs"""
        self.lib.mg_two_vcycle(self.h, x, b, rt().stream)
        return 1, 0.

    def solve(self, x, b, param=None):
        param = param or {}
        nite, res = ctypes.c_int(), ctypes.c_double()
        self.lib.mg_solve(self.h, x, b, float(param.get('tol', 1e-11)), int(param.get('maxite', 4)),
                          ctypes.byref(nite), ctypes.byref(res), rt().stream)
        return nite.value, res.value

    def Vcycle(self, lev1):
        self.lib.mg_vcycle(self.h, lev1, rt().stream)

    def Fcycle(self, lev1):
        self.lib.mg_fcycle(self.h, lev1, rt().stream)
