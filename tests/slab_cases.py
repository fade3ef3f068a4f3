"""Set-ups of the y-slab (multi-rank) tests, shared by the CPU worker (tests/slab_emu_worker.py,
gloo + emulated communicator) and the GPU worker (tests/slab_worker.py, one process per GPU):
builders(api, datadir, nx, ny, npy) whose initial state does not depend on the decomposition."""
import numpy as np

import cases


class RankAwareCoolRoof(object):
    """buoyancy flux +Q through the bottom row of the domain, -Q through the top row (the
    forcing of experiments/RayleighBenard/forcing_rayleigh.py, which tests grid.j0 / npy the
    same way); scales the buoyancy tendency by coef"""

    def __init__(self, param, grid):
        Q, nh = 1e-2, param.nh
        self.forc = grid.yr*0.
        if grid.j0 == grid.npy-1:
            self.forc[-nh-1, :] = -Q
        if grid.j0 == 0:
            self.forc[nh, :] = +Q
        self.forc *= grid.msk
        self.forc *= (1./grid.dx)

    def add_forcing(self, x, t, dxdt, coef=1.):
        dxdt[4] += self.forc
        dxdt[4] *= coef


def rayleigh_benard(api, datadir, nx, ny, npy):
    """the rb case of tests/golden/cases.py (Boussinesq, x-channel, forcing + diffusion +
    NO-SLIP walls) with a rank-independent initial state: the noise is drawn for the whole
    domain and every rank takes its rows"""
    param = api.Param('default.xml')
    param.modelname = 'boussinesq'
    cases._common(param, 'rb_slab', datadir)
    param.nx, param.ny, param.npy = nx, ny, npy
    param.Lx, param.Ly = 2., 2.*ny/nx
    param.geometry = 'xchannel'
    param.cfl, param.adaptable_dt, param.dt, param.dtmax = 1., True, .1, .1
    param.order = 5
    param.aparab = 0.02
    param.var_to_save = ['vorticity', 'buoyancy', 'v', 'psi']
    param.gravity = 1.
    param.forcing = True
    param.forcing_module = 'embedded'
    param.diffusion = True
    param.noslip = True
    grid = api.Grid(param)
    visco = .002*grid.dy
    param.Kdiff = {'vorticity': visco, 'buoyancy': visco}
    f2d = api.Fluid2d(param, grid)
    model = f2d.model
    model.forc = RankAwareCoolRoof(param, grid)
    nh = grid.nh
    np.random.seed(1)
    glob = np.random.normal(size=(ny, nx))
    glob -= glob.mean()
    rows = ny//npy
    noise = np.zeros_like(grid.yr)
    noise[nh:-nh, nh:-nh] = glob[grid.j0*rows:(grid.j0+1)*rows]
    noise *= grid.msk
    grid.fill_halo(noise)
    buoy = model.var.get('buoyancy')
    buoy += 1e-1*noise
    model.set_psi_from_vorticity()
    return f2d


BUILDERS = {
    "freedecay": lambda api, d, nx, ny, npy: cases.freedecay(api, d, nx, ny=ny, npy=npy),
    "rb": rayleigh_benard,
    "freedecay_flx": lambda api, d, nx, ny, npy: cases.freedecay(api, d, nx, ny=ny, npy=npy, diag_fluxes=True),
}
