"""The multi-GPU (y-slab) HOST layer on two CPU processes (gloo) with arithmetic attached:
the product's Python on the emulated C ABI whose communicator entry points run over the
process group (tests/emu_device.py: halo rows exchanged with the neighbours, scalars
all-reduced in rank order, the slab multigrid as the global problem on the oracle's
hierarchy).  Freedecay split in two slabs must give, BIT FOR BIT, the fields of the same
global problem on one rank -- the CPU twin of tests/test_gpu_slabs.py, which needs two GPUs.
What it checks is the host side of the decomposition (which fields are exchanged and when,
local masks and coordinates, global reductions, dt); the peer-memory kernels are the GPU
test's business."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("case,nx,ny,nsteps,nranks", [("freedecay", 32, 64, 3, 2), ("freedecay", 32, 32, 2, 2),
                                                      ("freedecay", 16, 64, 2, 4),
                                                      # Boussinesq x-channel with NO-SLIP walls, forcing, diffusion
                                                      ("rb", 64, 32, 3, 2),
                                                      # diag_fluxes: the reversible / irreversible stacks too
                                                      ("freedecay_flx", 32, 64, 2, 2)])
def test_slabs_reproduce_single_rank_run(tmp_path, case, nx, ny, nsteps, nranks):
    out = str(tmp_path/"rep.json")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks),
           "--master-addr", "127.0.0.1", "--master-port", "29541",
           os.path.join(HERE, "slab_emu_worker.py"), out, str(nx), str(ny), str(nsteps), case]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:]+p.stderr[-4000:]
    reports = json.load(open(out))
    r0 = reports[0]
    assert all(r0["fields_equal"].values()), r0["maxdiff"]
    assert r0.get("fluxes_equal", True), r0.get("fluxes_maxdiff")
    assert r0["slab_levels"] >= 1
    for rep in reports:
        assert rep["kt"] == nsteps
        assert rep["t"] == r0["one"]["t"] and rep["dt"] == r0["one"]["dt"]
        assert rep["solve"][0] == r0["one"]["solve"][0]
        for k, v in r0["one"]["diags"].items():      # sums folded in rank order (a cancelling sum is rounding noise)
            assert abs(rep["diags"][k]-v) <= 1e-12*abs(v)+1e-13, (k, rep["diags"][k], v)


def test_every_rank_streams_its_own_history_file(tmp_path):
    """two slabs through Fluid2d.loop(): each rank's <expname>_his_<rank> file exists after the
    loop (no end-of-run dump needed), holds one record per iteration + the initial one, of that
    rank's slab"""
    out = str(tmp_path/"rep.json")
    nx, ny, nsteps, nranks = 32, 64, 3, 2
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks),
           "--master-addr", "127.0.0.1", "--master-port", "29543",
           os.path.join(HERE, "slab_emu_worker.py"), out, str(nx), str(ny), str(nsteps), "freedecay_his"]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:]+p.stderr[-4000:]
    reports = json.load(open(out))
    assert len(reports) == nranks
    for rep in reports:
        assert rep["exists"] and "_his_%03i." % rep["rank"] in rep["hisfile"]
        assert rep["nrec"] == nsteps+1 and rep["shape"] == [nsteps+1, ny//nranks, nx]
        assert rep["last_is_my_slab"]
        assert rep["diag_exists"] == (rep["rank"] == 0)
