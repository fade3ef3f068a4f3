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
        self.comm = None        # f2d_comm_t* when the domain is split in y-slabs
        self.nranks = 1
        self.rank = 0

    # -- multi-GPU: symmetric heap shared through CUDA IPC ------------------------
    def ensure_comm(self, nranks, fieldbytes):
        """create (once) the communicator of the y-slab decomposition: every rank
        allocates an arena of the same size, the IPC handles travel through
        torch.distributed, f2d_comm_connect maps the peers' arenas"""
        if self.comm is not None or nranks == 1:
            return
        import torch.distributed as dist
        ensure_dist()
        if dist.get_world_size() != nranks:
            raise RuntimeError('param.npy = %d but %d processes were launched' % (nranks, dist.get_world_size()))
        self.rank, self.nranks = dist.get_rank(), nranks
        arena = int(os.environ.get('F2D_ARENA_BYTES', 80*fieldbytes + (256 << 20)))
        handle = ctypes.create_string_buffer(64)
        comm = ctypes.c_void_p()
        self.lib.comm_create(ctypes.byref(comm), self.rank, nranks, arena, handle)
        everyone = [None]*nranks
        dist.all_gather_object(everyone, bytes(handle.raw))
        self.lib.comm_connect(comm, b''.join(everyone))
        self.comm = comm
        dist.barrier()

    def alloc(self, shape, dtype=torch.float64):
        """zeroed device tensor; from the symmetric heap when the domain is decomposed"""
        if self.comm is None:
            return torch.zeros(shape, dtype=dtype, device=self.device)
        n = int(np.prod(shape))
        itemsize = torch.empty((), dtype=dtype).element_size()
        p = self.lib.comm_alloc(self.comm, n*itemsize)
        if not p:
            raise MemoryError('symmetric heap exhausted: raise F2D_ARENA_BYTES')

        class _Arena(object):
            pass
        holder = _Arena()
        holder.__cuda_array_interface__ = {
            'shape': (n,), 'typestr': '<f8' if dtype == torch.float64 else '|i1',
            'data': (int(p), False), 'version': 2}
        t = torch.as_tensor(holder, device=self.device)
        return t.view(shape)

    def exchange_y(self, ptr, ny, nx, nh=3):
        self.lib.comm_exchange_y(self.comm, ptr, nh, ny, nx, self.stream)

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


def ensure_dist():
    """process group of the slab decomposition (one process per GPU, launched by torchrun)"""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return
    if 'RANK' not in os.environ:
        raise RuntimeError('npy > 1 needs one process per GPU: launch with torchrun '
                           '(python -m torch.distributed.run --nproc-per-node N ...)')
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')) % torch.cuda.device_count())
    dist.init_process_group('nccl' if torch.cuda.is_available() else 'gloo')


_rt = None


def rt():
    global _rt
    if _rt is None:
        _rt = Runtime()
    return _rt
