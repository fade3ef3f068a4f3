"""Decomposition invariance: the same global problem on 2 GPUs
(y-slabs, peer halo exchange, gathered coarse levels) must give the fields of the
single-GPU run -- every cell is computed by the same arithmetic, only the sums behind
the residual norms, the mean of psi and the diagnostics are folded in another order.
Needs 2 GPUs (skipped otherwise)."""
import json
import os
import socket
import subprocess
import sys
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import cases  # noqa: E402

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def oracle_states(build, nsteps):
    """the same global problem on the CPU oracle (oracle/model.py): states after 0, 1 and nsteps steps"""
    import types
    from oracle import model as om
    api = types.SimpleNamespace(Param=om.Param, Grid=om.Grid, Fluid2d=om.Fluid2d)
    f = build(api, tempfile.mkdtemp())
    out = {"oracle_state0": np.array(f.model.var.state)}
    res = cases.run_steps(f, (1, nsteps))
    for k in (1, nsteps):
        out["oracle_state%d" % k] = res[k][0]
    return out


def check_against_oracle(rep, nsteps):
    """contract of BASELINE.json against the reference's algorithm: 1e-12 after one step, 1e-9 after ten"""
    assert rep.get("oracle_errors"), "the worker did not compare with the oracle"
    for key, e in rep["oracle_errors"].items():
        tol = 1e-9 if key.startswith("state%d:" % nsteps) and nsteps > 1 else 1e-12
        assert e <= tol, ("oracle", key, e)


def ngpus():
    import torch
    return torch.cuda.device_count()


# the last case is large enough for the slab levels to run the standard 64 x 32 tiles AND the
# small-level 64 x 8 tiles, each with the one-kernel level descent (k_zsmooth_resid_restrict<PEER>)
@pytest.mark.parametrize("nx,ny,min_cells", [(64, 128, "1500"), (64, 128, "100000000"), (128, 256, "3000"),
                                             (1024, 2048, "100000")])
def test_two_slabs_match_single_gpu(nx, ny, min_cells):
    slabs_match_single_gpu(2, nx, ny, min_cells)


@pytest.mark.parametrize("nranks,nx,ny,min_cells", [(4, 128, 512, "3000"), (4, 512, 2048, "100000"),
                                                    (8, 128, 1024, "3000")])
def test_more_slabs_match_single_gpu(nranks, nx, ny, min_cells):
    """the same invariance on 4 and 8 ranks (every rank has two distinct neighbours; the gather
    assembles 4 / 8 slabs); skipped on boxes with fewer GPUs"""
    slabs_match_single_gpu(nranks, nx, ny, min_cells)


def slabs_match_single_gpu(nranks, nx, ny, min_cells):
    if ngpus() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    import fluid2d_b200
    api = fluid2d_b200.api()
    nsteps = 3
    d = tempfile.mkdtemp()
    f2d = cases.freedecay(api, d, nx, ny=ny)
    ref = {"state0": np.array(f2d.model.var.state)}
    print("single-GPU solve:", f2d.model.ope.last_solve)
    res = cases.run_steps(f2d, (1, nsteps))
    for k in (1, nsteps):
        ref["state%d" % k] = res[k][0]
        ref["dt%d" % k] = np.array(res[k][2])
    ref.update(oracle_states(lambda api_, d_: cases.freedecay(api_, d_, nx, ny=ny), nsteps))
    refpath = os.path.join(d, "ref.npz")
    np.savez(refpath, **ref)
    out = os.path.join(d, "out.json")
    env = dict(os.environ, F2D_SLAB_MIN_CELLS=min_cells)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks),
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(REPO, "tests", "slab_worker.py"), refpath, out, str(nx), str(ny), str(nsteps)]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-3000:]
    rep = json.load(open(out))
    print(rep)
    if min_cells != "100000000":
        assert rep["slab_levels"] >= 1
    for key, e in rep["errors"].items():
        assert e <= 1e-12, (key, e)
    check_against_oracle(rep, nsteps)
    for k in (1, nsteps):
        assert abs(rep["dt%d" % k]-rep["dt%d_ref" % k]) <= 1e-12*rep["dt%d_ref" % k]


def test_two_slabs_match_single_gpu_with_noslip_walls():
    """Boussinesq x-channel with no-slip walls, forcing and diffusion (tests/slab_cases.py: rb):
    masked slab multigrid, boundary integral of the no-slip source through the all-reduce.
    The host side of this is pinned on two CPU ranks by tests/test_slab_emulated.py."""
    if ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    import fluid2d_b200
    from slab_cases import BUILDERS
    api = fluid2d_b200.api()
    nx, ny, nsteps = 128, 64, 3
    d = tempfile.mkdtemp()
    f2d = BUILDERS["rb"](api, d, nx, ny, 1)
    ref = {"state0": np.array(f2d.model.var.state)}
    res = cases.run_steps(f2d, (1, nsteps))
    for k in (1, nsteps):
        ref["state%d" % k] = res[k][0]
        ref["dt%d" % k] = np.array(res[k][2])
    ref.update(oracle_states(lambda api_, d_: BUILDERS["rb"](api_, d_, nx, ny, 1), nsteps))
    refpath = os.path.join(d, "ref.npz")
    np.savez(refpath, **ref)
    out = os.path.join(d, "out.json")
    env = dict(os.environ, F2D_SLAB_MIN_CELLS="1500")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           os.path.join(REPO, "tests", "slab_worker.py"), refpath, out, str(nx), str(ny), str(nsteps), "rb"]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-3000:]
    rep = json.load(open(out))
    for key, e in rep["errors"].items():
        assert e <= 1e-12, (key, e)
    check_against_oracle(rep, nsteps)
