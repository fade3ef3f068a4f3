"""Islands: non-zero streamfunction on interior boundaries (reference: core/island.py).
Host-side set-up; the two fields it produces (rhsp, psi) are uploaded by Fluid2d and
used inside Operators.invert_vorticity."""
import numpy as np

from param import Param


class Island(Param):
    def __init__(self, param, grid):
        self.nxl, self.nyl, self.dx, self.dy = grid.nxl, grid.nyl, grid.dx, grid.dy
        self.rhsp = np.zeros((self.nyl, self.nxl))
        self.psi = np.zeros((self.nyl, self.nxl))
        self.nbisland = 0
        self.data = []

    def add(self, idx, psi0):
        self.data.append({'idx': idx, 'psi0': psi0})
        self.nbisland += 1

    def finalize(self, mskp_model=None):
        print('found %i islands' % self.nbisland)
        shape = (self.nyl, self.nxl)
        for isl in self.data:
            cells = np.ones(shape, dtype=np.int8)
            cells[isl['idx']] = 0
            # corner is fluid iff its four cells are (celltocorner(mask) == 1)
            fluid = np.zeros(shape, dtype=np.int8)
            fluid[:-1, :-1] = cells[:-1, :-1] & cells[:-1, 1:] & cells[1:, :-1] & cells[1:, 1:]
            inside = (1-fluid).astype(np.int8)
            nb = (np.roll(inside, -1, axis=1)+np.roll(inside, -1, axis=0)
                  + np.roll(inside, +1, axis=1)+np.roll(inside, +1, axis=0))
            z = nb*isl['psi0']/(self.dx*self.dy)
            self.rhsp[nb > 0] = z[nb > 0]
            self.psi[inside == 1] = isl['psi0']
        print('island are ok')
