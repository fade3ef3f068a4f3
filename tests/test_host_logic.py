"""CPU-side tests (no GPU): the C ABI library loads and exports every symbol the header
declares; the host-side Param / Grid mirror the reference's defaults and masks (checked
against the oracle's pinned restatement); the host<->device coherence protocol of
DeviceState / TrackedArray (exercised on CPU tensors)."""
import ctypes
import os
import sys

import numpy as np
import pytest

import fluid2d_b200
from fluid2d_b200 import _lib

fluid2d_b200.activate()


def test_library_exports_every_declared_symbol():
    protos = _lib.parse_header()
    assert len(protos) >= 60
    for strict in (False, True):
        L = _lib.lib(strict=strict)
        for name in protos:
            assert hasattr(L.cdll, name), name
        assert L.abi_version() == 1
        assert L.reduce_scratch_len() > 0


def test_header_declares_the_reference_kernels():
    protos = _lib.parse_header()
    for fn in ["adv_upwind", "adv_centered", "celltocorner", "cornertocell", "orthogradient",
               "add_diffusion", "add_torque", "noslip_source", "fill_halo", "computedotprod", "computemax",
               "computesum", "computesumandnorm", "computenormmaxu", "computekemaxu", "computekemaxuv",
               "computekewithpsi", "computenorm", "mg_create", "mg_smooth", "mg_residual", "mg_restrict",
               "mg_interpolate", "mg_two_vcycle", "mg_solve", "invert_vorticity"]:
        assert "f2d_" + fn in protos, fn


def test_argument_errors_do_not_need_a_gpu():
    """shape / order / nh validation happens before any CUDA call and never exits"""
    L = _lib.lib()
    cst = (ctypes.c_double*5)(0.1, 0.1, 0.05, 0., 0.05)
    one = ctypes.c_void_p(8)
    with pytest.raises(_lib.F2DError) as e:
        L.adv_upwind(None, one, one, one, one, None, None, cst, 2, 1, 5, 32, 32, 0, None)
    assert e.value.code == L.ERR_NH and "NHALO" in str(e.value)
    with pytest.raises(_lib.F2DError) as e:
        L.adv_upwind(None, one, one, one, one, None, None, cst, 3, 1, 4, 32, 32, 0, None)
    assert e.value.code == L.ERR_ARG
    with pytest.raises(_lib.F2DError):
        L.fill_halo(None, 3, 10, 10, None)
    h = ctypes.c_void_p()
    with pytest.raises(_lib.F2DError):
        L.mg_create(ctypes.byref(h), one, 6+24, 6+32, 0.1, 0.1, 8./9., 1., 0., None)  # 24 not a power of 2


def test_param_defaults_match_reference_names():
    from param import Param
    from oracle import model as om
    p = Param('default.xml')
    for k, v in om.DEFAULTS.items():
        if k in ('beta', 'Rd', 'gravity'):    # not defaults: scripts add them (param.py free attributes)
            assert not hasattr(p, k)
            continue
        assert getattr(p, k) == v, k
    p.timestepping = 'nope'
    with pytest.raises(ValueError):
        p.checkall()
    class O(object):
        pass
    o = O()
    assert Param().copy(o, ['nx', 'notthere']) == ['notthere'] and o.nx == 128


@pytest.mark.parametrize("geometry", ['perio', 'closed', 'disc', 'xchannel', 'ychannel'])
def test_grid_matches_pinned_oracle(geometry):
    from param import Param
    from grid import Grid
    from oracle import model as om
    a, b = Param(), om.Param()
    for p in (a, b):
        p.nx, p.ny, p.Lx, p.Ly, p.geometry = 64, 32, 2., 1., geometry
    ga, gb = Grid(a), om.Grid(b)
    np.testing.assert_array_equal(ga.msk, gb.msk)
    np.testing.assert_array_equal(ga.xr, gb.xr)
    np.testing.assert_array_equal(ga.yr0, gb.yr0)
    assert ga.area == gb.area and ga.x2 == gb.x2 and ga.dx == gb.dx
    assert (ga.nxl, ga.nyl) == (70, 38)


def test_island_matches_pinned_oracle():
    from param import Param
    from grid import Grid
    from oracle import model as om
    out = []
    for P, G in ((Param, Grid), (om.Param, om.Grid)):
        p = P()
        p.nx, p.ny, p.geometry, p.isisland = 32, 32, 'closed', True
        g = G(p)
        idx = np.where((g.xr-0.5)**2+(g.yr-0.4)**2 < 0.02)
        g.msk[idx] = 0
        g.island.add(idx, 0.3)
        g.island.finalize()
        out.append((g.island.rhsp.copy(), g.island.psi.copy()))
    np.testing.assert_array_equal(out[0][0], out[1][0])
    np.testing.assert_array_equal(out[0][1], out[1][1])


def test_device_state_coherence_protocol():
    import torch
    from devarray import DeviceState
    s = DeviceState(3, 6, 8, device=torch.device('cpu'))
    v = s.host_view(1)
    v[:] = 2.                                    # host write -> device copy stale
    assert s.dev_fresh == [True, False, True]
    s.to_device(1)
    assert float(s.dev[1].sum()) == 2.*48 and s.dev_fresh[1]
    s.dev[1] += 1.                               # "kernel" writes the field
    s.wptr(1)
    assert not s.host_fresh[1]
    assert float(np.sum(v)) == 3.*48             # old view refreshes itself on read (ufunc)
    v *= 2.                                      # in-place operator marks the device stale again
    assert not s.dev_fresh[1]
    w = s[1]
    w += 1.
    s[1] = w                                     # the dxdt[k] += f idiom of user forcings
    s.to_device(None)
    assert float(s.dev[1][0, 0]) == 7.
    whole = s.host_view(None)
    whole[0][2:4, 2:4] = 5.
    assert not s.dev_fresh[0]
    assert s.numpy()[0, 2, 2] == 5.


def test_gridinfo_levels():
    """level sizes of the BASELINE configs"""
    from oracle.model import level_sizes
    assert len(level_sizes(4096, 4096)) == 11
    assert len(level_sizes(128, 128)) == 6
    assert level_sizes(2048, 1024)[-1] == (4, 8)
    assert level_sizes(4096, 1024)[-1] == (4, 16)


def test_inplace_ops_on_field_views_host_path():
    """TrackedArray in-place arithmetic (the forcing-hook idiom) on a CPU DeviceState takes the
    host path and must equal numpy; the device fast path is covered by the GPU tests"""
    import numpy as np
    import torch
    import fluid2d_b200
    fluid2d_b200.activate()
    from devarray import DeviceState
    rng = np.random.default_rng(4)
    st = DeviceState(2, 12, 14, device=torch.device("cpu"))
    a0 = rng.standard_normal((2, 12, 14))
    st.upload_all_from(a0)
    f = rng.standard_normal((12, 14))
    ref = a0.copy()
    v = st[1]
    v += f
    v *= 0.25
    v -= f
    ref[1] += f
    ref[1] *= 0.25
    ref[1] -= f
    w = st[0]
    w[2:5, :] *= 3.
    ref[0][2:5, :] *= 3.
    np.testing.assert_array_equal(st.numpy(), ref)
