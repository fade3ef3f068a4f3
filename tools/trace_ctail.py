"""Diagnostics for the cluster tail kernel: per-barrier SM clock stamps of one V-cycle of a
128^2 hierarchy (f2d_mg_set_trace), and the duration of one V-cycle launch for the cluster
tail / the one-CTA tail / the per-level kernels.   python tools/trace_ctail.py [n]"""
import ctypes
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from fluid2d_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
L = _lib.lib()
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
ptr = lambda t: ctypes.c_void_p(t.data_ptr())
cm = torch.ones((n+6, n+6), dtype=torch.float64, device="cuda")
cm[-1, :] = 0
cm[:, -1] = 0
x = torch.zeros((n+6, n+6), dtype=torch.float64, device="cuda")
b = torch.randn((n+6, n+6), dtype=torch.float64, device="cuda")
L.fill_halo(ptr(b), 3, n+6, n+6, s)


def timeit(f, reps=20):
    f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return 1e3*e0.elapsed_time(e1)/reps


for name, env in (("cluster tail", {"F2D_MG_CTAIL": "1"}), ("one-CTA tail", {}),
                  ("per-level kernels", {"F2D_MG_NO_TAIL": "1"})):
    for k in ("F2D_MG_CTAIL", "F2D_MG_NO_TAIL"):
        os.environ[k] = env.get(k, "0")
    h = ctypes.c_void_p()
    L.mg_create(ctypes.byref(h), ptr(cm), n+6, n+6, 1./n, 1./n, 8./9., 1., 0., s)
    L.copy(L.mg_level_ptr(h, 0, 3), ptr(b), (n+6)*(n+6)*8, s) if False else None
    us = timeit(lambda: L.mg_two_vcycle(h, ptr(x), ptr(b), s))
    print("%-18s two V-cycles of %d^2: %.1f us" % (name, n, us))
    if name == "cluster tail":
        cap = 4096
        tr = torch.zeros(cap, dtype=torch.int64, device="cuda")
        L.mg_set_trace(h, ptr(tr), cap)
        L.mg_set_graphs(h, 0)
        L.mg_vcycle(h, 0, s)
        torch.cuda.synchronize()
        t = tr.cpu().numpy()
        k = int(t[0])
        st = t[1:1+k]
        d = st[1:]-st[:-1]
        print("barriers: %d, total %.1f us at 1.965 GHz" % (k, (st[-1]-st[0])/1965.))
        print("cycles between barriers:", " ".join(str(int(v)) for v in d))
        L.mg_set_trace(h, None, 0)
    L.mg_destroy(h)
