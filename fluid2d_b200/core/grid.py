"""Grid: sizes, coordinates and masks of (the local part of) the domain.

Same attributes and behaviour as the reference's core/grid.py:9-140 -- nxl/nyl with the
halo, dx/dy, xr/yr (meshgrid of cell centres), the default mask per geometry
(set_msk), msknoslip, finalize_msk (area, barycentre, xr0/yr0, x2/y2, r2),
domain_integration, and fill_halo (installed by Operators).  Masks and coordinates are
host numpy arrays because user scripts edit them before Fluid2d() is built; they are
uploaded once when the operators are created.
"""
import os

import numpy as np

from param import Param
from mpitools import Mpitools

FROM_PARAM = ('nx', 'ny', 'npx', 'npy', 'Lx', 'Ly', 'nh', 'geometry', 'myrank', 'mpi', 'isisland',
              'enforce_momentum', 'hydroepsilon')
# geometries with solid walls on the y sides (south/north) / on the x sides (west/east)
WALLS_SOUTH_NORTH = ('xperio', 'xchannel', 'closed', 'disc')
WALLS_WEST_EAST = ('yperio', 'ychannel', 'closed', 'disc')


def _this_rank(param):
    """rank of this process in the y-slab decomposition (one process per GPU)"""
    nranks = param.npx*param.npy
    if nranks == 1:
        return 0
    if param.npx != 1:
        raise NotImplementedError('the domain is decomposed in y-slabs: use npx = 1, npy = number of GPUs')
    from runtime import ensure_dist
    ensure_dist()
    return int(os.environ.get('RANK', '0'))


def default_mask(geometry, xr, yr, Lx, Ly, nh, i0, j0, npx, npy):
    """int8 mask of a subdomain, 1 = fluid: walls are nh cells thick and sit in the halo of
    the subdomains that touch the corresponding side; 'disc' also masks outside the circle"""
    msk = np.ones(xr.shape, dtype=np.int8)
    if geometry in WALLS_SOUTH_NORTH:
        if j0 == npy-1:
            msk[-nh:, :] = 0
        if j0 == 0:
            msk[:nh, :] = 0
    if geometry in WALLS_WEST_EAST:
        if i0 == npx-1:
            msk[:, -nh:] = 0
        if i0 == 0:
            msk[:, :nh] = 0
    if geometry == 'disc':
        radius = np.sqrt((xr/Lx-0.5)**2 + (yr/Ly-0.5)**2)
        msk[radius >= 0.5] = 0
    return msk


class Grid(Param):
    def __init__(self, param):
        param.myrank = _this_rank(param)
        param.nbproc = param.npx*param.npy
        param.nx, param.ny = int(param.nx), int(param.ny)
        param.copy(self, FROM_PARAM)
        self.debug = False
        self.mpitools = Mpitools(param)
        nh = self.nh
        # local sizes (halo included) and position of this subdomain
        self.nxl, self.nyl = self.nx//self.npx+2*nh, self.ny//self.npy+2*nh
        self.i0, self.j0 = self.myrank % self.npx, (self.myrank // self.npx) % self.npy
        self.dx, self.dy = self.Lx/self.nx, self.Ly/self.ny
        if self.dx != self.dy and self.myrank == 0:
            print('dx = %g and dy = %g differ: the model is not fully validated for that' % (self.dx, self.dy))
        # cell-centre coordinates, halo cells included
        first_i, first_j = self.i0*self.nx//self.npx, self.j0*self.ny//self.npy
        self.x1d = (np.arange(self.nxl)+0.5-nh+first_i)*self.dx
        self.y1d = (np.arange(self.nyl)+0.5-nh+first_j)*self.dy
        self.xr, self.yr = np.meshgrid(self.x1d, self.y1d)
        self.set_msk()
        if self.isisland:
            from island import Island
            self.island = Island(param, self)

    def set_msk(self):
        """default mask of param.geometry (0 = solid); users may edit grid.msk afterwards"""
        self.msk = default_mask(self.geometry, self.xr, self.yr, self.Lx, self.Ly, self.nh,
                                self.i0, self.j0, self.npx, self.npy)
        self.msknoslip = self.msk.copy()
        self.finalize_msk()

    def finalize_msk(self):
        """quantities that depend on the mask; call again after editing grid.msk"""
        msk, integrate = self.msk, self.domain_integration
        self.area = integrate(msk)
        # barycentre of the fluid, coordinates relative to it, second moments
        self.x0 = integrate(self.xr*msk) / self.area
        self.y0 = integrate(self.yr*msk) / self.area
        self.xr0 = (self.xr - self.x0)*msk
        self.yr0 = (self.yr - self.y0)*msk
        self.x2 = integrate((self.xr0)**2*msk) / self.area
        self.y2 = integrate((self.yr0)**2*msk) / self.area
        self.r2 = self.xr0**2 + self.yr0**2
        if self.debug and self.myrank == 0:
            print('fluid barycentre (%g, %g), %i fluid cells' % (self.x0, self.y0, self.area))

    def domain_integration(self, z2d):
        """sum over the interior cells of all subdomains (host arrays; set-up time)"""
        nh = self.nh
        integral = np.sum(np.asarray(z2d)[nh:-nh, nh:-nh])*1.
        return self.mpitools.local_to_global([(integral, 'sum')])

    def fill_halo(self, x):
        raise RuntimeError('grid.fill_halo becomes available once Fluid2d (the operators) is built')
