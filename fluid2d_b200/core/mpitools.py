"""Global reductions of a few scalars across the ranks (reference: core/mpitools.py,
allgather + numpy sum/max).  One process per GPU; when torch.distributed is initialised
the scalars go through one all_gather on the process group, otherwise this is the
identity (single rank)."""
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
