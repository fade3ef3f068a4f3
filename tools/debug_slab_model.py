"""debug: halo consistency of the slab model state after set-up (torchrun, 2 ranks)"""
import os, sys, tempfile
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
import torch, torch.distributed as dist
import fluid2d_b200, cases
api = fluid2d_b200.api()
world = int(os.environ["WORLD_SIZE"])
so = sys.stdout; sys.stdout = sys.stderr
f2d = cases.freedecay(api, tempfile.mkdtemp(), 64, ny=128, npy=world)
sys.stdout = so
model = f2d.model
rank = dist.get_rank()
ope = model.ope

def allparts(t):
    parts = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(parts, t.contiguous())
    return [p.cpu().numpy() for p in parts]

def check(name, t):
    P = allparts(t)
    if rank == 0:
        for r in range(world):
            n = (r+1) % world
            top = np.abs(P[r][-3:, :] - P[n][3:6, :]).max()      # my top halo vs north's first interior rows
            bot = np.abs(P[n][:3, :] - P[r][-6:-3, :]).max()     # north's bottom halo vs my last interior rows
            xl = np.abs(P[r][:, :3] - P[r][:, -6:-3]).max()
            print("%-10s rank %d: top-halo mismatch %.3e  north-bottom-halo mismatch %.3e  x-halo mismatch %.3e  absmax %.3e"
                  % (name, r, top, bot, xl, np.abs(P[r]).max()), flush=True)

s = model.var.dstate
for k, nm in enumerate(model.var.varname_list):
    check(nm, s.dev[k])
check("work", ope.work)
print(rank, "last solve", ope.last_solve, "slab levels", ope.gmg.slab_levels, flush=True)
# one more full solve, printing residual
model.set_psi_from_vorticity()
print(rank, "again     ", ope.last_solve, flush=True)
dist.barrier()
