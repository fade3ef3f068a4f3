"""Global reductions of a few scalars across the ranks (reference: core/mpitools.py,
allgather + numpy sum/max).  One process per GPU; when torch.distributed is initialised
the scalars go through one all_gather on the process group, otherwise this is the
identity (single rank).

reduce_device() is the path of the per-step diagnostics on y-slabs: the partial sums already
sit on the device, so they are all-reduced there (f2d_comm_allreduce: one kernel over peer
memory, sums folded in rank order, a bit mask selects the maxima) and cross PCIe once --
no host-synchronous collective inside the time loop."""
import numpy as np


class Mpitools(object):
    def __init__(self, param):
        self.npx = param.npx
        self.npy = param.npy
        self.myrank = param.myrank
        self.nbproc = self.npx*self.npy

    def local_to_global(self, list_scalars):
        nb = len(list_scalars)
        cst = np.zeros((nb,))
        for k in range(nb):
            cst[k] = list_scalars[k][0]
        if self.nbproc > 1:
            import torch
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                dev = 'cuda' if dist.get_backend() == 'nccl' else 'cpu'
                mine = torch.from_numpy(cst.copy()).to(dev)
                everyone = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
                dist.all_gather(everyone, mine)
                glo = np.stack([e.cpu().numpy() for e in everyone])
                for k in range(nb):
                    ope = list_scalars[k][1]
                    cst[k] = np.max(glo[:, k]) if ope == 'max' else np.sum(glo[:, k])
        return cst

    def reduce_device(self, r, n, maxmask=0):
        """r.out[:n] (device scalars written by the reduction kernels) -> global values as floats;
        bit k of maxmask marks slot k as a maximum instead of a sum"""
        if r.comm is not None:
            r.lib.comm_allreduce(r.comm, r.ptr(r.out), n, maxmask, r.stream)
        return r.read_out(n)
