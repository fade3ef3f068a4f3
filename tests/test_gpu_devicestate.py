"""DeviceState on a real device: uploads run on a copy stream of their own, every stale field at
once with one event per field (fluid2d_b200/core/devarray.py).  Whatever the order in which
kernels ask for fields and the host rewrites them, a kernel must see the host's last write and
the host must read the device's last write."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup():
    import torch
    import fluid2d_b200
    fluid2d_b200.activate()
    from devarray import DeviceState
    from runtime import rt
    return torch, DeviceState, rt()


def test_upload_batches_follow_host_writes():
    torch, DeviceState, r = _setup()
    nvar, ny, nx = 5, 70, 134
    ds = DeviceState(nvar, ny, nx)
    rng = np.random.default_rng(3)
    host = ds[:]                      # TrackedArray on the pinned mirror
    ref = rng.standard_normal((nvar, ny, nx))
    host[...] = ref
    n = ny*nx
    for rep in range(6):
        # ask for the fields in a different order each time; a kernel doubles the one asked for first
        order = list(rng.permutation(nvar))
        k = int(order[0])
        p = ds.wptr(k)
        r.lib.ts_axpy(p, 1.0, ctypes.c_void_p(p.value), n, r.stream)   # x += 1.0 * x
        ref[k] *= 2.
        for f in order[1:]:
            ds.rptr(int(f))
        # the host rewrites another field while uploads of the batch may still be in flight
        j = int(order[1])
        new = rng.standard_normal((ny, nx))
        ds[j] = new
        ref[j] = new
        got = np.array(ds[:])
        np.testing.assert_array_equal(got, ref)
        # everything goes back to the device next time
        ds._host_written(None)
    assert ds.h2d_bytes > 0 and ds.d2h_bytes > 0


def test_kernel_sees_last_host_write_of_every_field():
    torch, DeviceState, r = _setup()
    nvar, ny, nx = 4, 38, 70
    ds = DeviceState(nvar, ny, nx)
    rng = np.random.default_rng(4)
    n = ny*nx
    for rep in range(4):
        vals = rng.standard_normal((nvar, ny, nx))
        v = ds[:]
        v[...] = vals
        out = torch.zeros((ny, nx), dtype=torch.float64, device="cuda")
        # out = sum over fields, each requested one by one (the first request starts the batch)
        for f in range(nvar):
            r.lib.ts_axpy(ctypes.c_void_p(out.data_ptr()), 1.0, ds.rptr(f), n, r.stream)
        torch.cuda.synchronize()
        want = np.zeros((ny, nx))
        for f in range(nvar):
            want = want + 1.0*vals[f]
        np.testing.assert_array_equal(out.cpu().numpy(), want)
