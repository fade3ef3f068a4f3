"""Worker of tests/test_slab_emulated.py: two CPU processes (gloo), the product's host layer on
the emulated C ABI with the y-slab communicator (tests/emu_device.py).  Builds freedecay with
npy = world size, advances it, gathers the global fields and compares them with the same
global problem run on ONE rank in the same process (the CPU twin of tests/slab_worker.py).

    torchrun --nproc-per-node 2 tests/slab_emu_worker.py out.json nx ny nsteps
"""
import contextlib
import io
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))

import emu_device  # noqa: E402
import cases  # noqa: E402


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


def glue(emu, stack):
    """[k][local rows][nx] on every rank -> [k][global rows][nx]"""
    parts = [np.stack(emu.lib._gather(stack[k])) for k in range(stack.shape[0])]
    return np.stack([np.concatenate([p[0][:3]]+[q[3:-3] for q in p]+[p[-1][-3:]], axis=0) for p in parts])


def main():
    out, nx, ny, nsteps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    build = BUILDERS[sys.argv[5] if len(sys.argv) > 5 else "freedecay"]
    world = int(os.environ["WORLD_SIZE"])
    api, emu = emu_device.install()
    with contextlib.redirect_stdout(io.StringIO()):
        f2d = build(api, tempfile.mkdtemp(), nx, ny, world)
        res = cases.run_steps(f2d, (nsteps,))[nsteps]
        flx = cases.run_fluxes(f2d) if getattr(f2d, "diag_fluxes", False) else None
    import torch.distributed as dist
    rank = dist.get_rank()
    glob = glue(emu, np.array(res[0]))
    gflx = glue(emu, np.array(flx)) if flx is not None else None
    report = {"rank": rank, "kt": f2d.kt, "t": res[1], "dt": res[2], "diags": res[3],
              "solve": list(f2d.model.ope.last_solve), "slab_levels": f2d.model.ope.gmg.slab_levels}
    if rank == 0:
        # the same global problem on one rank (a fresh single-rank emulator; no process group use)
        emu_device.uninstall()
        api1, emu1 = emu_device.install()
        with contextlib.redirect_stdout(io.StringIO()):
            one = build(api1, tempfile.mkdtemp(), nx, ny, 1)
            r1 = cases.run_steps(one, (nsteps,))[nsteps]
            flx1 = cases.run_fluxes(one) if flx is not None else None
        ref = np.array(r1[0])
        names = list(one.model.var.varname_list)
        report["fields_equal"] = {nm: bool(np.array_equal(glob[k][3:-3], ref[k][3:-3])) for k, nm in enumerate(names)}
        report["maxdiff"] = {nm: float(np.abs(glob[k][3:-3]-ref[k][3:-3]).max()) for k, nm in enumerate(names)}
        if flx is not None:
            report["fluxes_equal"] = bool(np.array_equal(gflx[:, 3:-3], np.array(flx1)[:, 3:-3]))
            report["fluxes_maxdiff"] = float(np.abs(gflx[:, 3:-3]-np.array(flx1)[:, 3:-3]).max())
        report["one"] = {"t": r1[1], "dt": r1[2], "diags": r1[3], "solve": list(one.model.ope.last_solve)}
    everyone = [None]*world
    dist.all_gather_object(everyone, report)
    if rank == 0:
        json.dump(everyone, open(out, "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
