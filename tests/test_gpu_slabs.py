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


def ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("nx,ny,min_cells", [(64, 128, "1500"), (64, 128, "100000000"), (128, 256, "3000")])
def test_two_slabs_match_single_gpu(nx, ny, min_cells):
    if ngpus() < 2:
        pytest.skip("needs 2 GPUs")
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
    refpath = os.path.join(d, "ref.npz")
    np.savez(refpath, **ref)
    out = os.path.join(d, "out.json")
    env = dict(os.environ, F2D_SLAB_MIN_CELLS=min_cells)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
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
    for k in (1, nsteps):
        assert abs(rep["dt%d" % k]-rep["dt%d_ref" % k]) <= 1e-12*rep["dt%d_ref" % k]


def test_two_slabs_match_single_gpu_with_noslip_walls():
    """Boussinesq x-channel with no-slip walls, forcing and diffusion (tests/slab_cases.py: rb):
    masked slab multigrid, boundary integral of the no-slip source through the all-reduce.
    The host side of this is pinned on two CPU ranks by tests/test_slab_emulated.py."""
    if ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    if os.environ.get("F2D_TEST_UNVERIFIED_SLABS") != "1":
        # the masked slab multigrid has not run on real GPUs yet (written after the round's GPU
        # budget was spent): opt-in until tools/r2_first_gpu_session.sh slabs has been green once
        pytest.skip("set F2D_TEST_UNVERIFIED_SLABS=1 (first run: tools/r2_first_gpu_session.sh slabs)")
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
