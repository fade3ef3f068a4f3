"""Worker of tests/test_slab_host.py: two CPU processes (gloo).  Exercises the host side of
the y-slab decomposition -- Grid's local coordinates / masks / global integrals and
Mpitools.local_to_global -- without any device call.

    torchrun --nproc-per-node 2 tests/slab_host_worker.py out.json
"""
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import torch.distributed as dist  # noqa: E402
import fluid2d_b200  # noqa: E402


def main():
    out = sys.argv[1]
    api = fluid2d_b200.api()
    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ["RANK"])
    rep = {}
    for geometry in ("closed", "perio", "xchannel", "disc"):
        param = api.Param()
        param.nx, param.ny, param.npx, param.npy = 32, 48, 1, world
        param.Lx, param.Ly = 2., 3.
        param.geometry = geometry
        grid = api.Grid(param)
        nh = grid.nh
        # the same problem on one rank: plain numpy, no process group involved
        one = api.Param()
        one.nx, one.ny, one.npx, one.npy = 32, 48, 1, 1
        one.Lx, one.Ly = 2., 3.
        one.geometry = geometry
        full = api.Grid(one)
        j0 = rank*(param.ny//world)
        rows = slice(j0, j0+grid.nyl)
        r = {
            "shape": [grid.nyl, grid.nxl],
            "yr": float(np.abs(grid.yr-full.yr[rows]).max()),
            "xr": float(np.abs(grid.xr-full.xr[rows]).max()),
            "msk": int(np.abs(grid.msk[nh:-nh].astype(int)-full.msk[rows][nh:-nh].astype(int)).max()),
            "area": [float(grid.area), float(full.area)],
            "x0": [float(grid.x0), float(full.x0)], "y0": [float(grid.y0), float(full.y0)],
            "x2": [float(grid.x2), float(full.x2)], "y2": [float(grid.y2), float(full.y2)],
            "r2": float(np.abs(grid.r2[nh:-nh]-full.r2[rows][nh:-nh]).max()),
        }
        # the walls of a closed / channel domain exist on the outer ranks only
        r["south_wall"] = int(grid.msk[:nh].max() == 0)
        r["north_wall"] = int(grid.msk[-nh:].max() == 0)
        rep[geometry] = r
    mt = grid.mpitools
    glo = mt.local_to_global([(float(rank+1), "sum"), (float(rank+1), "max"), (-float(rank), "max")])
    rep["reduce"] = [float(v) for v in glo]
    everyone = [None]*world
    dist.all_gather_object(everyone, rep)
    if rank == 0:
        json.dump(everyone, open(out, "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
