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
        """per island: psi = psi0 on every corner that touches it, and on the fluid corners next
        to it the right-hand-side correction (number of island neighbours) * psi0 / (dx dy)"""
        print('found %i islands' % self.nbisland)
        shape = (self.nyl, self.nxl)
        cell_area = self.dx*self.dy
        for island in self.data:
            psi0 = island['psi0']
            solid = np.zeros(shape, dtype=bool)
            solid[island['idx']] = True
            # a corner belongs to the island unless its four cells are all outside it
            touches = np.ones(shape, dtype=np.int8)
            touches[:-1, :-1] = (solid[:-1, :-1] | solid[:-1, 1:] | solid[1:, :-1] | solid[1:, 1:])
            neighbours = sum(np.roll(touches, shift, axis=axis) for axis in (0, 1) for shift in (-1, 1))
            near = neighbours > 0
            self.rhsp[near] = (neighbours*psi0/cell_area)[near]
            self.psi[touches == 1] = psi0
        print('island are ok')
