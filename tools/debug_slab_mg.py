"""debug: slab multigrid vs single-GPU multigrid on the same global problem (torchrun, >= 2 ranks)"""
import ctypes, os, sys
import numpy as np
import torch
import torch.distributed as dist
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from fluid2d_b200 import _lib

nx, ny = int(sys.argv[1]), int(sys.argv[2])
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
L = _lib.lib()
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
ptr = lambda t: ctypes.c_void_p(t.data_ptr())
nh = 3
ml = ny//world
fieldbytes = (ml+6)*(nx+6)*8
handle = ctypes.create_string_buffer(64)
comm = ctypes.c_void_p()
L.comm_create(ctypes.byref(comm), rank, world, 200*fieldbytes + (64 << 20), handle)
ev = [None]*world
dist.all_gather_object(ev, bytes(handle.raw))
L.comm_connect(comm, b''.join(ev))
dist.barrier()

def arena(shape):
    n = int(np.prod(shape))
    p = L.comm_alloc(comm, n*8)
    class H: pass
    h = H(); h.__cuda_array_interface__ = {'shape': (n,), 'typestr': '<f8', 'data': (int(p), False), 'version': 2}
    return torch.as_tensor(h, device='cuda').view(shape)

rng = np.random.default_rng(0)
G = rng.standard_normal((ny, nx)); G -= G.mean()
P0 = 0.01*rng.standard_normal((ny, nx))
def with_halo(a):
    return np.pad(a, 3, mode='wrap')
Gh, Ph = with_halo(G), with_halo(P0)
geom = sys.argv[4] if len(sys.argv) > 4 else "perio"
msk = np.ones((ny+6, nx+6))
if geom in ("xchannel", "closed"):
    msk[:3, :] = 0; msk[-3:, :] = 0
if geom == "closed":
    msk[:, :3] = 0; msk[:, -3:] = 0
if geom == "obstacle":
    yy, xx = np.mgrid[0:ny+6, 0:nx+6]
    msk[:3, :] = 0; msk[-3:, :] = 0
    msk[(yy-ny*0.5)**2+(xx-nx*0.3)**2 < (0.15*min(nx, ny))**2] = 0
cm = np.zeros((ny+6, nx+6))
cm[:-1, :-1] = ((msk[:-1, :-1]+msk[:-1, 1:]+msk[1:, :-1]+msk[1:, 1:]) == 4)*1.
def slab(a):
    return np.ascontiguousarray(a[rank*ml:rank*ml+ml+6, :])
# what the host layer hands to f2d_mg_create_slab: the corner mask of the LOCAL cell mask
# (operators.py:59-67 per rank: its last row and column are 0 on every rank)
mskl = slab(msk)
cml = np.zeros((ml+6, nx+6))
cml[:-1, :-1] = ((mskl[:-1, :-1]+mskl[:-1, 1:]+mskl[1:, :-1]+mskl[1:, 1:]) == 4)*1.
Gh *= cm; Ph *= cm
h = ctypes.c_void_p()
L.mg_create_slab(ctypes.byref(h), comm, ptr(torch.from_numpy(cml).cuda()), ml+6, nx+6, 1./nx, 1./nx, 8./9., 1., 0., s)
print(rank, "levels", L.mg_nlevels(h), "slab levels", L.mg_slab_levels(h), "modes", [L.mg_level_matrix_mode(h, l) for l in range(L.mg_nlevels(h))], flush=True)
psi = arena((ml+6, nx+6)); rhs = arena((ml+6, nx+6))
psi.copy_(torch.from_numpy(slab(Ph))); rhs.copy_(torch.from_numpy(slab(Gh)))
torch.cuda.synchronize(); dist.barrier()

def gather(t):
    parts = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(parts, t.contiguous())
    parts = [p.cpu().numpy() for p in parts]
    return np.concatenate([parts[0][:3]] + [p[3:-3] for p in parts] + [parts[-1][-3:]], axis=0)

if rank == 0:
    h1 = ctypes.c_void_p()
    L.mg_create(ctypes.byref(h1), ptr(torch.from_numpy(cm).cuda()), ny+6, nx+6, 1./nx, 1./nx, 8./9., 1., 0., s)
    p1 = torch.from_numpy(Ph.copy()).cuda(); r1 = torch.from_numpy(Gh.copy()).cuda()

order = sys.argv[3].split(",") if len(sys.argv) > 3 else ["two_vcycle", "two_vcycle", "solve"]
for step in order:
    nite, res = ctypes.c_int(), ctypes.c_double()
    if step == "two_vcycle":
        L.mg_two_vcycle(h, ptr(psi), ptr(rhs), s)
    else:
        L.mg_solve(h, ptr(psi), ptr(rhs), 1e-11, 4, ctypes.byref(nite), ctypes.byref(res), s)
    torch.cuda.synchronize()
    g = gather(psi)
    if rank == 0:
        n1, r1s = ctypes.c_int(), ctypes.c_double()
        if step == "two_vcycle":
            L.mg_two_vcycle(h1, ptr(p1), ptr(r1), s)
        else:
            L.mg_solve(h1, ptr(p1), ptr(r1), 1e-11, 4, ctypes.byref(n1), ctypes.byref(r1s), s)
        torch.cuda.synchronize()
        ref = p1.cpu().numpy()
        e = np.linalg.norm((g-ref)[3:-3, :])/np.linalg.norm(ref[3:-3, :])
        eh = np.abs(g-ref).max()
        print(step, "rel err interior %.3e  max abs incl halos %.3e" % (e, eh), "nite", nite.value, n1.value, "res", res.value, r1s.value, flush=True)
dist.barrier()
