"""helpers for the -m gpu tests: device buffers through torch, calls through the C ABI"""
import ctypes

import numpy as np
import torch

from fluid2d_b200 import _lib


def dev(a):
    """numpy -> contiguous cuda tensor (float64 / int8 preserved)"""
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


_alive = []


def keep(a):
    """device copy that stays allocated until release() (a bare dev(x).data_ptr() would
    hand the caching allocator's block straight to the next temporary)"""
    t = dev(a)
    _alive.append(t)
    if len(_alive) > 4096:
        torch.cuda.synchronize()
        del _alive[:2048]
    return t


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def host(t):
    torch.cuda.synchronize()
    return t.cpu().numpy()


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def scratch(L):
    return torch.zeros(L.reduce_scratch_len(), dtype=torch.float64, device="cuda")


def rel_l2(a, b):
    d = np.linalg.norm((a - b).ravel())
    n = np.linalg.norm(b.ravel())
    return d / n if n > 0 else d


def check(a, b, strict, tol=1e-13, what=""):
    """bit-exact for the -fmad=false build, relative L2 <= tol for the product build"""
    if strict:
        np.testing.assert_array_equal(a, b, err_msg=what)
    else:
        e = rel_l2(a, b)
        assert e <= tol, "%s: rel L2 %.3e > %.1e" % (what, e, tol)


def libs():
    return [("strict", _lib.lib(strict=True), True), ("product", _lib.lib(strict=False), False)]


def check_res(res, res_ref):
    """the residual norm a solve reports against the oracle's: relative 1e-6 while it is a
    meaningful number (> 1e-9); once the solve has converged to rounding level the norm is
    rounding noise (an FMA build differs in the 4th digit at 5e-13), so absolute 1e-12"""
    if abs(res_ref) > 1e-9:
        assert abs(res - res_ref) <= 1e-6 * abs(res_ref), (res, res_ref)
    else:
        assert abs(res - res_ref) <= 1e-12, (res, res_ref)
