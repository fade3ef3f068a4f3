"""CPU stand-in for the device runtime, for tests of the HOST logic only.

The host layer (fluid2d_b200/core) is Python that enqueues calls of libf2d_b200.so.  This
module swaps the runtime singleton for one whose library records every call (entry point +
scalar arguments) and computes nothing, with the state held in CPU tensors.  What a test can
then check without a GPU is the orchestration: which entry points a scenario reaches, in
which order, with which scalar arguments -- `tests/test_host_trace.py` freezes those traces
from the GPU-verified host layer so that refactorings of the Python cannot change what is
sent to the device.  No arithmetic is validated here (that is what the -m gpu tests do).
"""
import ctypes
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


_REGISTRY = []   # (start address, bytes, label, bytes per field) of every known buffer, in creation order


def register(t, kind, fieldbytes=0):
    label = "%s%d" % (kind, sum(1 for r in _REGISTRY if r[2].startswith(kind)))
    _REGISTRY.append((t.data_ptr(), t.numel()*t.element_size(), label, fieldbytes, t))
    return t


def _pointer(addr):
    """which buffer (and which field of a state) an address points into"""
    for start, nbytes, label, fb, _keep in _REGISTRY:
        if start <= addr < start+max(nbytes, 1):
            off = addr-start
            return "%s:%d" % (label, off//fb) if fb else ("%s+%d" % (label, off) if off else label)
    return "P"


def _summary(a):
    if a is None:
        return "N"
    if isinstance(a, bool):
        return int(a)
    if isinstance(a, (int, np.integer)):
        return int(a)
    if isinstance(a, (float, np.floating)):
        return float("%.12g" % float(a))
    if isinstance(a, ctypes.Array):
        if a._type_ is ctypes.c_void_p:        # host array of device pointers (f2d_adv_multi)
            return [_pointer(v) if v else "N" for v in a]
        return [float("%.12g" % float(v)) for v in a]
    if isinstance(a, ctypes.c_void_p):
        return _pointer(a.value) if a.value else "N"
    return type(a).__name__


def _level_sizes(m, n):
    out = []
    while True:
        out.append((m, n))
        if n <= 4 or m <= 4:
            break
        m, n = m//2, n//2
    return out


class FakeLib(object):
    def __init__(self):
        self.calls = []
        self.levels = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)

        def call(*args):
            self.calls.append([name]+[_summary(a) for a in args])
            return self._effect(name, args)
        call.__name__ = name
        return call

    def _effect(self, name, args):
        if name == "reduce_scratch_len":
            return 64
        if name in ("launch_count", "mg_slab_levels", "comm_rank"):
            return 0
        if name == "abi_version" or name == "comm_size":
            return 1
        if name in ("mg_create", "mg_create_slab"):
            k = 1 if name == "mg_create" else 2
            args[0]._obj.value = 1
            ny, nx = args[k+1], args[k+2]
            self.levels = _level_sizes(ny-6, nx-6)
            return 0
        if name == "mg_nlevels":
            return len(self.levels)
        if name == "mg_level_shape":
            m, n = self.levels[args[1]]
            args[2]._obj.value, args[3]._obj.value = m+6, n+6
            return 0
        if name == "mg_level_matrix_mode":
            return 1
        if name == "mg_level_ptr":
            return 0
        if name == "invert_vorticity":
            if args[10]:      # full solve: (nite, res) come back through the two host pointers
                args[16]._obj.value, args[17]._obj.value = 3, 1e-12
            return 0
        if name == "mg_solve":
            args[5]._obj.value, args[6]._obj.value = 3, 1e-12
            return 0
        return 0


class FakeRuntime(object):
    def __init__(self):
        self.lib = FakeLib()
        self.device = torch.device("cpu")
        del _REGISTRY[:]
        self.scratch = register(torch.zeros(64, dtype=torch.float64), "scratch")
        self.out = register(torch.zeros(16, dtype=torch.float64), "out")
        self.out_host = torch.zeros(16, dtype=torch.float64)
        self.comm = None
        self.nranks = 1
        self.rank = 0

    def ensure_comm(self, nranks, fieldbytes):
        if nranks != 1:
            raise NotImplementedError("mock runtime: one rank")

    def alloc(self, shape, dtype=torch.float64):
        return register(torch.zeros(shape, dtype=dtype), "A")

    @property
    def stream(self):
        return None

    def ptr(self, t):
        return ctypes.c_void_p(t.data_ptr()) if t is not None else None

    def to_device(self, a, dtype=None):
        return register(torch.from_numpy(np.ascontiguousarray(a, dtype=dtype).copy()), "U")

    def read_out(self, n):
        """what the reductions 'returned': a deterministic, non-trivial sequence (slot k of the
        m-th read is 0.05 + 0.01 k + 0.001 m), so that the host arithmetic on the diagnostics
        and the time step it sets show up in the traces"""
        self.reads = getattr(self, "reads", 0)+1
        return [0.05+0.01*k+0.001*self.reads for k in range(n)]


def install():
    """activate the host layer on top of the recording runtime; returns (api, runtime)"""
    import fluid2d_b200
    api = fluid2d_b200.api()
    import runtime
    import devarray
    fake = FakeRuntime()
    runtime._rt = fake
    patch_devicestate(devarray)
    return api, fake


class _NoStream(object):
    cuda_stream = 0

    def synchronize(self):
        pass


def patch_devicestate(devarray):
    """DeviceState blocks live in CPU tensors (shared with tests/emu_device.py)"""
    if not getattr(devarray.DeviceState, "_mock_patched", False):
        real_init = devarray.DeviceState.__init__

        def init(self, nvar, ny, nx, device=None):
            real_init(self, nvar, ny, nx, device=torch.device("cpu"))
            register(self.dev, "S", ny*nx*8)
        devarray.DeviceState.__init__ = init
        devarray.DeviceState._mock_patched = real_init
    if not torch.cuda.is_available():
        torch.cuda.synchronize = lambda *a, **k: None
        torch.cuda.current_stream = lambda *a, **k: _NoStream()


def uninstall():
    """back to the real runtime (a GPU test that follows in the same process must get
    device-resident state again)"""
    import runtime
    import devarray
    runtime._rt = None
    real_init = getattr(devarray.DeviceState, "_mock_patched", False)
    if real_init:
        devarray.DeviceState.__init__ = real_init
        devarray.DeviceState._mock_patched = False
