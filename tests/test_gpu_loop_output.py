"""Fluid2d.loop() end to end on the device: the time loop with history / diagnostics /
flux output (core/fluid2d.py:188-349, core/output.py), snapshots cast to float32 on the
device (f2d_pack_interior_f32), and the stop-and-continue path through Restart."""
import os
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import cases  # noqa: E402


def _load(path):
    import output
    return output.load_records(path)


@pytest.mark.parametrize("diag_fluxes", [False, True])
def test_loop_writes_history_diagnostics_and_fluxes(diag_fluxes):
    import fluid2d_b200
    api = fluid2d_b200.api()
    f2d = cases.freedecay(api, tempfile.mkdtemp(), 32, diag_fluxes=diag_fluxes)
    out = f2d.output
    out.freq_his = 0.       # a snapshot at every iteration
    out.freq_diag = 0.
    out.tnexthis = out.tnextdiag = 0.
    f2d.exacthistime = False
    nsteps = 3
    f2d.loop(nsteps=nsteps)
    assert f2d.kt == nsteps
    his = _load(out.hisfile)
    nh = 3
    state = np.array(f2d.model.var.state, copy=True)
    for name in out.var_to_save:
        k = f2d.model.var.index(name)
        assert his[name].shape == (nsteps+1, 32, 32) and his[name].dtype == np.float32
        np.testing.assert_array_equal(his[name][-1], state[k][nh:-nh, nh:-nh].astype(np.float32))
    assert len(his["t"]) == nsteps+1 and abs(float(his["t"][-1])-f2d.t) <= 1e-6*f2d.t
    diag = _load(out.diagfile)
    assert len(diag["t"]) == nsteps+1
    np.testing.assert_allclose(diag["ke"][-1], float(np.ravel(f2d.model.diags["ke"])[0]), rtol=1e-6)
    assert diag["ke"][-1] <= diag["ke"][0]*(1+1e-6)       # free decay: energy does not grow
    if diag_fluxes:
        flx = _load(out.flxfile)
        names = f2d.flx.fullflx_list
        stack = np.array(f2d.flx.flx, copy=True)
        for k, name in enumerate(names):
            assert flx[name].shape == (nsteps+1, 32, 32)
            np.testing.assert_array_equal(flx[name][-1], stack[k][nh:-nh, nh:-nh].astype(np.float32))
        # the reversible vorticity flux dominates the irreversible one
        assert np.abs(flx["rev_x_vorticity"][-1]).max() > np.abs(flx["irr_x_vorticity"][-1]).max()


def test_pack_interior_f32_matches_numpy():
    import ctypes
    import gpu_util as g
    from fluid2d_b200 import _lib
    lib = _lib.lib()
    import torch
    rng = np.random.default_rng(0)
    for ny, nx in ((10, 12), (70, 38), (135, 262)):
        x = rng.standard_normal((ny, nx))*1e3
        out = torch.zeros((ny-6, nx-6), dtype=torch.float32, device="cuda")
        lib.pack_interior_f32(g.ptr(g.keep(x)), ctypes.c_void_p(out.data_ptr()), 3, ny, nx, g.stream())
        np.testing.assert_array_equal(g.host(out), x[3:-3, 3:-3].astype(np.float32))


def test_inplace_ops_on_field_views_run_on_the_device_and_match_numpy():
    """the forcing-hook idiom `dxdt[k] += f; dxdt[k] *= c` (boussinesq.py:110, user scripts)"""
    import torch
    from fluid2d_b200 import activate
    activate()
    from devarray import DeviceState
    rng = np.random.default_rng(1)
    ny, nx = 38, 70
    st = DeviceState(3, ny, nx)
    a0 = rng.standard_normal((3, ny, nx))
    st.upload_all_from(a0)
    f = rng.standard_normal((ny, nx))
    ref = a0.copy()
    v = st[1]                 # host view; the device copy is current
    d2h = st.d2h_bytes
    v += f
    v *= 0.37
    v -= 2.*f
    v *= f
    ref[1] += f
    ref[1] *= 0.37
    ref[1] -= 2.*f
    ref[1] *= f
    assert st.d2h_bytes == d2h, "the field must not have crossed PCIe"
    assert not st.host_fresh[1]
    np.testing.assert_array_equal(st.numpy(), ref)
    # patterns that do not qualify take the host path and still give numpy's answer
    w = st[2]
    w[3:-3, :] += 1.5
    ref[2][3:-3, :] += 1.5
    r = st[0][::-1]
    r += f
    ref[0][::-1] += f
    np.testing.assert_array_equal(st.numpy(), ref)
    torch.cuda.synchronize()
