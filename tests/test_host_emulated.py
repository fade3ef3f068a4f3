"""The product's HOST layer (fluid2d_b200/core: models, Operators, Timescheme, Fluxes, dt
control, diagnostics) run end to end on a CPU emulation of the C ABI (tests/emu_device.py:
every entry point executed by the oracle's kernels) and compared with the fixtures frozen
from the reference's own Python (tests/golden/*.npz).

Bar: BIT-IDENTICAL states, time steps, diagnostics and flux stacks after 0, 1 and 10 steps.
That is possible because every floating-point operation of a step is either inside an
entry point (executed here by the same oracle kernels the reference ran on) or a scalar
expression of the host layer that must be written exactly as the reference writes it.
A refactoring of the host layer that reorders a sum, drops a halo fill or passes the wrong
field fails here without a GPU.  The CUDA kernels themselves are NOT exercised (see the
-m gpu tests); the emulator is test infrastructure and unreachable from the product.
"""
import contextlib
import io
import os
import tempfile

import numpy as np
import pytest

import cases
import emu_device

GOLDEN = os.path.dirname(os.path.abspath(cases.__file__))


@pytest.fixture()
def emu():
    api, rt_ = emu_device.install()
    yield api, rt_
    emu_device.uninstall()


@pytest.mark.parametrize("name", sorted(cases.CASES))
def test_host_layer_on_emulated_abi_reproduces_reference_run(name, emu):
    api, rt_ = emu
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    with contextlib.redirect_stdout(io.StringIO()):
        f2d = cases.CASES[name](api, tempfile.mkdtemp())
    model = f2d.model
    names = list(model.var.varname_list)
    assert names == [str(s) for s in gold["varnames"]]
    np.testing.assert_array_equal(np.asarray(model.ope.msk), gold["grid_msk"])
    gmg = model.ope.gmg
    assert gmg.nlevs == int(gold["mg_nlevs"])
    for lev in range(0 if name in cases.LIGHT else gmg.nlevs):
        np.testing.assert_array_equal(gmg.grid[lev].msk, gold["mg_msk%i" % lev])
        np.testing.assert_array_equal(gmg.grid[lev].A, gold["mg_A%i" % lev])
    if name in cases.SPECTRAL:
        np.testing.assert_allclose(np.array(model.var.state, copy=True), gold["state0"], rtol=0,
                                   atol=1e-13*np.abs(gold["state0"]).max())
    else:
        np.testing.assert_array_equal(np.array(model.var.state, copy=True), gold["state0"])
    with contextlib.redirect_stdout(io.StringIO()):
        if "flx0" in gold:
            assert f2d.flx.fullflx_list == [str(s) for s in gold["flxnames"]]
            np.testing.assert_array_equal(cases.run_fluxes(f2d), gold["flx0"])
        res = cases.run_steps(f2d)
    for nstep, (state, t, dt, diags) in sorted(res.items()):
        if name in cases.SPECTRAL:
            # torch's CPU FFT against numpy's: equal to rounding (1e-12 is the GPU contract)
            for k, nm in enumerate(names):
                g = gold["state%i" % nstep][k]
                e = np.linalg.norm(state[k]-g)/max(np.linalg.norm(g), 1e-300)
                assert e <= 1e-12, "%s after %d steps: rel L2 %.2e" % (nm, nstep, e)
            assert abs(dt-float(gold["dt%i" % nstep])) <= 1e-12*abs(dt)
            continue
        for k, nm in enumerate(names):
            np.testing.assert_array_equal(state[k], gold["state%i" % nstep][k],
                                          err_msg="%s after %d steps" % (nm, nstep))
        assert dt == float(gold["dt%i" % nstep])
        assert t == float(gold["t%i" % nstep])
        for dn, dv in diags.items():
            assert dv == float(gold["diag%i_%s" % (nstep, dn)]), (nstep, dn)
    if "flx10" in gold:
        with contextlib.redirect_stdout(io.StringIO()):
            np.testing.assert_array_equal(cases.run_fluxes(f2d), gold["flx10"])


def test_emulator_covers_the_entry_points_the_host_layer_binds():
    """every lib.<entry point> the host layer mentions exists in the emulator, and every
    emulated entry point is declared in include/f2d_b200.h (the emulator cannot drift into
    an interface the real library does not have)"""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "f2d_b200.h")).read()
    declared = set(re.findall(r"\bf2d_([a-z0-9_]+)\s*\(", header))
    emulated = {n for n in dir(emu_device.EmuLib) if not n.startswith("_") and callable(getattr(emu_device.EmuLib, n))}
    assert emulated <= declared, sorted(emulated-declared)
    used = set()
    core = os.path.join(root, "fluid2d_b200", "core")
    for dirpath, _d, files in os.walk(core):
        for f in files:
            if f.endswith(".py"):
                used |= set(re.findall(r"\blib\.([a-z0-9_]+)\(", open(os.path.join(dirpath, f)).read()))
    used &= declared
    set_up_only = {"comm_create", "comm_connect", "comm_alloc"}     # CUDA IPC arena: EmuRuntime.ensure_comm / alloc
    assert used-set_up_only <= emulated, sorted(used-set_up_only-emulated)


def _load(path):
    import output
    return output.load_records(path)


def test_loop_writes_history_and_diagnostics_on_emulated_abi(emu):
    """Fluid2d.loop() with a snapshot at every iteration: history records are the float32
    interiors of the state, the diag file follows the loop (the CPU twin of
    tests/test_gpu_loop_output.py)"""
    api, rt_ = emu
    with contextlib.redirect_stdout(io.StringIO()):
        f2d = cases.freedecay(api, tempfile.mkdtemp(), 32, diag_fluxes=True)
        out = f2d.output
        out.freq_his = out.freq_diag = 0.
        out.tnexthis = out.tnextdiag = 0.
        f2d.exacthistime = False
        f2d.loop(nsteps=3)
    assert f2d.kt == 3
    his = _load(out.hisfile)
    state = np.array(f2d.model.var.state, copy=True)
    for name in out.var_to_save:
        k = f2d.model.var.index(name)
        assert his[name].shape == (4, 32, 32) and his[name].dtype == np.float32
        np.testing.assert_array_equal(his[name][-1], state[k][3:-3, 3:-3].astype(np.float32))
    diag = _load(out.diagfile)
    assert len(diag["t"]) == 4 and diag["ke"][-1] <= diag["ke"][0]
    flx = _load(out.flxfile)
    stack = np.array(f2d.flx.flx, copy=True)
    for k, name in enumerate(f2d.flx.fullflx_list):
        np.testing.assert_array_equal(flx[name][-1], stack[k][3:-3, 3:-3].astype(np.float32))


def test_restart_round_trip_on_emulated_abi(emu):
    """two jobs through Restart (restart.py: tend is the LENGTH of a job; the second job starts
    from the file the first one wrote; output files carry the job index).  What a restart
    carries over is the model state in double with its halos and the clocks -- bit for bit; the
    time scheme's tendency buffers (whose psi is the first guess of the truncated inversions)
    are not in the file, in the reference either, so the continuation agrees with an
    uninterrupted run to the truncation of those inversions, not to rounding."""
    import glob
    import types
    api, rt_ = emu
    made = {}

    class F(api.Fluid2d):
        def __init__(self, param, grid):
            made['param'], made['grid'] = param, grid
            api.Fluid2d.__init__(self, param, grid)
    api2 = types.SimpleNamespace(Param=api.Param, Grid=api.Grid, Fluid2d=F)
    datadir = tempfile.mkdtemp()
    from restart import Restart
    with contextlib.redirect_stdout(io.StringIO()):
        f1 = cases.freedecay(api2, datadir, 32)
        f1.exacthistime = False
        f1.tend = 1.0
        made['param'].ninterrestart = 2
        r1 = Restart(made['param'], made['grid'], f1)
        assert r1.nextrestart == 0 and f1.kt > 0
        end1 = np.array(f1.model.var.state, copy=True)
        f2 = cases.freedecay(api2, datadir, 32)
        f2.exacthistime = False
        f2.tend = 1.0
        made['param'].ninterrestart = 2
        r2 = Restart(made['param'], made['grid'], f2, launch=False)
    assert r2.lastrestart == 0 and r2.nextrestart == 1
    np.testing.assert_array_equal(np.array(f2.model.var.state, copy=True), end1)
    assert (f2.t, f2.dt, f2.kt) == (f1.t, f1.dt, f1.kt)
    assert (f2.output.tnexthis, f2.output.tnextdiag) == (f1.output.tnexthis, f1.output.tnextdiag)
    assert os.path.basename(f1.output.diagfile).startswith('freedecay_32_00_diag')
    assert os.path.basename(f2.output.diagfile).startswith('freedecay_32_01_diag')
    with contextlib.redirect_stdout(io.StringIO()):
        r2.launchf2d(f2)
        ref = cases.freedecay(api, tempfile.mkdtemp(), 32)
        ref.exacthistime = False
        ref.loop(nsteps=f2.kt)
    assert f2.kt > f1.kt and f2.t >= f1.t+1.0
    assert len(glob.glob(datadir+'/freedecay_32/freedecay_32_*_restart_000.*')) == 2
    a, b = np.array(f2.model.var.state, copy=True), np.array(ref.model.var.state, copy=True)
    assert np.linalg.norm(a-b) <= 1e-2*np.linalg.norm(b)
