"""The oracle's restated orchestration (oracle/model.py) must reproduce, bit for bit,
the fixtures produced by the REFERENCE's own Python running on the same oracle
kernels (tests/golden/make_golden.py).  This pins oracle/model.py, which is what the
GPU box uses where /root/reference is absent."""
import io
import os
import sys
import tempfile
import types

import numpy as np
import pytest

import cases
from oracle import model as om

GOLDEN = os.path.dirname(os.path.abspath(cases.__file__))
ORACLE_API = types.SimpleNamespace(Param=om.Param, Grid=om.Grid, Fluid2d=om.Fluid2d)

SUPPORTED = ['freedecay_64', 'freedecay_32_o3_notracer', 'vortex_64',
             'vortex_32_twall_o5', 'rb_64', 'karman_32', 'freedecay_32_flx', 'rb_32_flx'] + sorted(cases.LIGHT)


@pytest.mark.parametrize("name", SUPPORTED)
def test_oracle_model_matches_reference_run(name):
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    f2d = cases.CASES[name](ORACLE_API, tempfile.mkdtemp())
    model = f2d.model
    assert list(gold["varnames"]) == list(model.var.varname_list)
    np.testing.assert_array_equal(model.var.state, gold["state0"])
    mg = model.ope.gmg
    assert mg.nlevs == int(gold["mg_nlevs"])
    for lev in range(0 if name in cases.LIGHT else mg.nlevs):
        np.testing.assert_array_equal(mg.msk[lev], gold["mg_msk%i" % lev])
        np.testing.assert_array_equal(mg.A[lev], gold["mg_A%i" % lev])
    if "flx0" in gold:
        assert list(gold["flxnames"]) == f2d.flx.fullflx_list
        np.testing.assert_array_equal(cases.run_fluxes(f2d), gold["flx0"])
    res = cases.run_steps(f2d)
    for k, (state, t, dt, diags) in res.items():
        assert dt == float(gold["dt%i" % k])
        assert t == float(gold["t%i" % k])
        np.testing.assert_array_equal(state, gold["state%i" % k])
        for dn, dv in diags.items():
            assert dv == float(gold["diag%i_%s" % (k, dn)]), dn
    if "flx10" in gold:
        np.testing.assert_array_equal(cases.run_fluxes(f2d), gold["flx10"])
