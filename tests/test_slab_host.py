"""Host side of the multi-GPU (y-slab) path on two CPU processes (gloo): local grids are
the right windows of the global grid, global integrals/barycentres agree with the
single-rank values, Mpitools reduces over the ranks.  No device call is made."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def reports(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("slabhost")/"rep.json")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(HERE, "slab_host_worker.py"), out]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:]+p.stderr[-4000:]
    return json.load(open(out))


@pytest.mark.parametrize("geometry", ["closed", "perio", "xchannel", "disc"])
def test_local_grid_is_window_of_global(reports, geometry):
    for rank, rep in enumerate(reports):
        r = rep[geometry]
        assert r["shape"] == [48//2+6, 32+6]
        assert r["yr"] < 1e-14 and r["xr"] < 1e-14
        assert r["msk"] == 0
        assert r["r2"] < 1e-12
        for k in ("area", "x0", "y0", "x2", "y2"):
            a, b = r[k]
            assert abs(a-b) <= 1e-12*max(1., abs(b)), (k, a, b)
        walls_y = geometry in ("closed", "xchannel", "disc")
        assert r["south_wall"] == int(walls_y and rank == 0)
        assert r["north_wall"] == int(walls_y and rank == len(reports)-1)


def test_mpitools_reduces_over_ranks(reports):
    for rep in reports:
        assert rep["reduce"] == [3., 2., 0.]
