"""End-to-end parity of the product (fluid2d_b200: Python host + CUDA kernels through
the C ABI) against the golden fixtures produced by the REFERENCE's own Python running on
the oracle kernels (tests/golden/make_golden.py).

Every case is built by the same user-script lines (tests/golden/cases.py) that built the
fixture, this time against fluid2d_b200's Param / Grid / Fluid2d.  Tolerances are the
contract of BASELINE.json: relative L2 <= 1e-12 after one step, <= 1e-9 after ten,
per prognostic / diagnosed field; mask and index handling exact (checked per kernel in
test_gpu_kernels.py with the -fmad=false build).
"""
import os
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import cases  # noqa: E402

GOLDEN = os.path.dirname(os.path.abspath(cases.__file__))
TOL = {0: 1e-12, 1: 1e-12, 10: 1e-9}


def rel(a, b):
    n = np.linalg.norm(b.ravel())
    d = np.linalg.norm((a-b).ravel())
    return d/n if n > 0 else d


def check_fluxes(f2d, gold, key, tol, report):
    """core/fluxes.py through the device Fluxes driver: reversible / irreversible stack.
    Each irreversible flux is a half difference of two integrations whose half sum is the
    reversible one, so both are measured against the larger of the two norms."""
    assert f2d.flx.fullflx_list == [str(s) for s in gold["flxnames"]]
    flx = cases.run_fluxes(f2d)
    g = gold[key]
    nflx = len(f2d.flx.flx_list)
    for k, nm in enumerate(f2d.flx.fullflx_list):
        scale = max(np.linalg.norm(g[k % nflx]), np.linalg.norm(g[nflx + k % nflx]))
        e = np.linalg.norm(flx[k]-g[k])/(scale if scale > 0 else 1.)
        report.append((key, nm, e))
        assert e <= tol, "%s %s: rel L2 %.3e > %.0e" % (key, nm, e, tol)


@pytest.mark.parametrize("name", sorted(set(cases.CASES)-cases.LATE))
def test_case_matches_reference_run(name):
    import fluid2d_b200
    api = fluid2d_b200.api()
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    f2d = cases.CASES[name](api, tempfile.mkdtemp())
    model = f2d.model
    names = list(model.var.varname_list)
    assert names == [str(s) for s in gold["varnames"]]
    np.testing.assert_array_equal(np.asarray(model.ope.msk), gold["grid_msk"])
    # multigrid hierarchy: masks exact, matrices to rounding
    gmg = model.ope.gmg
    assert gmg.nlevs == int(gold["mg_nlevs"])
    for lev in range(0 if name in cases.LIGHT else gmg.nlevs):
        np.testing.assert_array_equal(gmg.grid[lev].msk, gold["mg_msk%i" % lev])
        Aref = gold["mg_A%i" % lev]
        np.testing.assert_allclose(gmg.grid[lev].A, Aref, rtol=1e-13, atol=1e-13*np.abs(Aref).max())
    state0 = np.array(model.var.state, copy=True)
    report = []
    for k, nm in enumerate(names):
        e = rel(state0[k], gold["state0"][k])
        report.append((0, nm, e))
        assert e <= TOL[0], "initial %s: rel L2 %.3e" % (nm, e)
    if "flx0" in gold:
        check_fluxes(f2d, gold, "flx0", TOL[1], report)
    res = cases.run_steps(f2d)
    for nstep, (state, t, dt, diags) in sorted(res.items()):
        g = gold["state%i" % nstep]
        tol = TOL[nstep]
        assert abs(dt-float(gold["dt%i" % nstep])) <= tol*abs(dt)
        assert abs(t-float(gold["t%i" % nstep])) <= tol*abs(t)
        for k, nm in enumerate(names):
            e = rel(state[k], g[k])
            report.append((nstep, nm, e))
            assert e <= tol, "%s after %d steps: rel L2 %.3e > %.0e" % (nm, nstep, e, tol)
        for dn, dv in diags.items():
            gv = float(gold["diag%i_%s" % (nstep, dn)])
            assert abs(dv-gv) <= max(tol*100*abs(gv), 1e-13), (dn, dv, gv)
    if "flx10" in gold:
        check_fluxes(f2d, gold, "flx10", TOL[10], report)
    print(name, " ".join("%s:%s=%.1e" % r for r in report))
