"""Process-wide handles of the device runtime: the C-ABI library, the CUDA device and
stream (torch is the allocator / stream / process-group plumbing), reduction scratch.

There is no CPU path: importing this module on a machine without a CUDA device works
(so that the host-side classes can be unit-tested), but the first device call raises.
"""
import ctypes
import os

import numpy as np
import torch

from fluid2d_b200 import _lib


class Runtime(object):
    def __init__(self):
        self.lib = _lib.lib(strict=bool(int(os.environ.get('F2D_STRICT', '0'))))
        if not torch.cuda.is_available():
            raise RuntimeError('fluid2d_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
        local = int(os.environ.get('LOCAL_RANK', '0'))
        if torch.cuda.device_count() > 1:
            torch.cuda.set_device(local % torch.cuda.device_count())
        self.device = torch.device('cuda', torch.cuda.current_device())
        self.scratch = torch.zeros(self.lib.reduce_scratch_len(), dtype=torch.float64, device=self.device)
        self.out = torch.zeros(16, dtype=torch.float64, device=self.device)
        self.out_host = torch.zeros(16, dtype=torch.float64).pin_memory()

    @property
    def stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def ptr(self, t):
        return ctypes.c_void_p(t.data_ptr()) if t is not None else None

    def to_device(self, a, dtype=None):
        a = np.ascontiguousarray(a, dtype=dtype)
        return torch.from_numpy(a).to(self.device)

    def read_out(self, n):
        """device scalars self.out[:n] -> python floats (one synchronising D2H)"""
        self.out_host[:n].copy_(self.out[:n], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.out_host[:n].tolist()


_rt = None


def rt():
    global _rt
    if _rt is None:
        _rt = Runtime()
    return _rt
