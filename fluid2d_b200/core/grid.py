"""Grid: sizes, coordinates and masks of (the local part of) the domain.

Same attributes and behaviour as the reference's core/grid.py:9-140 -- nxl/nyl with the
halo, dx/dy, xr/yr (meshgrid of cell centres), the default mask per geometry
(set_msk), msknoslip, finalize_msk (area, barycentre, xr0/yr0, x2/y2, r2),
domain_integration, and fill_halo (installed by Operators).  Masks and coordinates are
host numpy arrays because user scripts edit them before Fluid2d() is built; they are
uploaded once when the operators are created.
"""
import numpy as np

from param import Param
from mpitools import Mpitools


class Grid(Param):
    def __init__(self, param):
        import os
        rank = int(os.environ.get('RANK', '0')) if param.npx*param.npy > 1 else 0
        if param.npx*param.npy > 1:
            if param.npx != 1:
                raise NotImplementedError('the domain is decomposed in y-slabs: use npx = 1, npy = number of GPUs')
            from runtime import ensure_dist
            ensure_dist()
        param.myrank = rank
        param.nbproc = param.npx*param.npy
        param.nx = int(param.nx)
        param.ny = int(param.ny)
        self.list_param = ['nx', 'ny', 'npx', 'npy', 'Lx', 'Ly', 'myrank', 'nh', 'geometry', 'mpi',
                           'enforce_momentum', 'isisland', 'hydroepsilon']
        param.copy(self, self.list_param)
        self.debug = False
        self.mpitools = Mpitools(param)
        nh = self.nh
        self.nxl = self.nx//self.npx+2*nh
        self.nyl = self.ny//self.npy+2*nh
        self.i0 = self.myrank % self.npx
        self.j0 = (self.myrank // self.npx) % self.npy
        self.dx = self.Lx/self.nx
        self.dy = self.Ly/self.ny
        if self.dx != self.dy and self.myrank == 0:
            print('dx and dy are different')
            print('the model does not allow it')
            print('model is not yet fully validated')
        ishift = self.i0*self.nx//self.npx
        jshift = self.j0*self.ny//self.npy
        self.x1d = (np.arange(self.nxl)+0.5-nh+ishift)*self.dx
        self.y1d = (np.arange(self.nyl)+0.5-nh+jshift)*self.dy
        self.xr, self.yr = np.meshgrid(self.x1d, self.y1d)
        self.set_msk()
        if self.isisland:
            from island import Island
            self.island = Island(param, self)

    def set_msk(self):
        """default mask of param.geometry (0 = solid); users may edit grid.msk afterwards"""
        nh = self.nh
        msk = np.ones((self.nyl, self.nxl), dtype=np.int8)
        walls_y = self.geometry in ['xperio', 'xchannel', 'closed', 'disc']
        walls_x = self.geometry in ['yperio', 'ychannel', 'closed', 'disc']
        if walls_y and self.j0 == self.npy-1:
            msk[-nh:, :] = 0
        if walls_y and self.j0 == 0:
            msk[:nh, :] = 0
        if walls_x and self.i0 == self.npx-1:
            msk[:, -nh:] = 0
        if walls_x and self.i0 == 0:
            msk[:, :nh] = 0
        if self.geometry == 'disc':
            r = np.sqrt((self.xr/self.Lx-0.5)**2 + (self.yr/self.Ly-0.5)**2)
            msk[r >= 0.5] = 0
        self.msk = msk
        self.msknoslip = self.msk.copy()
        self.finalize_msk()

    def finalize_msk(self):
        """quantities that depend on the mask; call again after editing grid.msk"""
        msk = self.msk
        self.area = self.domain_integration(msk)
        x0 = self.domain_integration(self.xr*msk) / self.area
        y0 = self.domain_integration(self.yr*msk) / self.area
        self.xr0 = (self.xr - x0)*msk
        self.yr0 = (self.yr - y0)*msk
        self.x2 = self.domain_integration((self.xr0)**2*msk) / self.area
        self.y2 = self.domain_integration((self.yr0)**2*msk) / self.area
        self.x0 = x0
        self.y0 = y0
        if (self.myrank == 0) and (self.debug):
            print('domain barycenter is at (x0,y0)=(%g,%g)' % (x0, y0))
            print('domain has %i interior points' % self.area)
        self.r2 = self.xr0**2 + self.yr0**2

    def domain_integration(self, z2d):
        """sum over the interior cells of all subdomains (host arrays; set-up time)"""
        nh = self.nh
        integral = np.sum(np.asarray(z2d)[nh:-nh, nh:-nh])*1.
        return self.mpitools.local_to_global([(integral, 'sum')])

    def fill_halo(self, x):
        raise RuntimeError('grid.fill_halo becomes available once Fluid2d (the operators) is built')
