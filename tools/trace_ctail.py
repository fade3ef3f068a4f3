"""Diagnostics for the coarse-tail kernels: duration of one V-cycle of an n^2 hierarchy (n <= 256:
the whole hierarchy is the tail) for the cluster tail in several configurations, the one-CTA tail
and the per-level kernels; per-barrier SM clock stamps of the cluster tail (f2d_mg_set_trace);
and the V-cycle / F-cycle of a 4096^2 hierarchy with each configuration.
    python tools/trace_ctail.py [n] [big]"""
import ctypes
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from fluid2d_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
big = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
L = _lib.lib()
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
ptr = lambda t: ctypes.c_void_p(t.data_ptr())


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


CONFIGS = [("periodic tail 16, <=128", {"F2D_CTAIL_NC": "16", "F2D_PTAIL_MAXN": "128"}),
           ("periodic tail 16, <=256", {"F2D_CTAIL_NC": "16", "F2D_PTAIL_MAXN": "256"}),
           ("periodic tail 16, <=128, all dist", {"F2D_CTAIL_NC": "16", "F2D_PTAIL_MAXN": "128", "F2D_CTAIL_MINCELLS": "0"}),
           ("periodic tail 16, <=128, dist>=8k", {"F2D_CTAIL_NC": "16", "F2D_PTAIL_MAXN": "128", "F2D_CTAIL_MINCELLS": "8192"}),
           ("periodic tail 8, <=128", {"F2D_CTAIL_NC": "8", "F2D_PTAIL_MAXN": "128"}),
           ("periodic tail 1 (one CTA, <=64)", {"F2D_CTAIL_NC": "1", "F2D_PTAIL_MAXN": "64"}),
           ("cluster 16, <=256", {"F2D_CTAIL_NC": "16", "F2D_CTAIL_MAXN": "256", "F2D_MG_NO_PTAIL": "1"}),
           ("cluster 16, <=128", {"F2D_CTAIL_NC": "16", "F2D_CTAIL_MAXN": "128", "F2D_MG_NO_PTAIL": "1"}),
           ("one-CTA tail (round 1)", {"F2D_MG_NO_CTAIL": "1"}),
           ("per-level kernels", {"F2D_MG_NO_TAIL": "1"})]
KEYS = ("F2D_CTAIL_NC", "F2D_CTAIL_MAXN", "F2D_CTAIL_MINCELLS", "F2D_MG_NO_CTAIL", "F2D_MG_NO_TAIL", "F2D_MG_NO_PTAIL",
        "F2D_PTAIL_MAXN")


def setenv(env):
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update(env)


def hierarchy(m):
    cm = torch.ones((m+6, m+6), dtype=torch.float64, device="cuda")
    cm[-1, :] = 0
    cm[:, -1] = 0
    h = ctypes.c_void_p()
    L.mg_create(ctypes.byref(h), ptr(cm), m+6, m+6, 1./m, 1./m, 8./9., 1., 0., s)
    return h


x = torch.zeros((n+6, n+6), dtype=torch.float64, device="cuda")
b = torch.randn((n+6, n+6), dtype=torch.float64, device="cuda")
L.fill_halo(ptr(b), 3, n+6, n+6, s)
for name, env in CONFIGS:
    setenv(env)
    h = hierarchy(n)
    us = timeit(lambda: L.mg_two_vcycle(h, ptr(x), ptr(b), s))/2
    usf = timeit(lambda: L.mg_fcycle(h, 0, s))
    print("%-30s %d^2: V-cycle %.1f us, F-cycle %.1f us" % (name, n, us, usf))
    if (name.startswith("cluster") or name.startswith("periodic")) and os.environ.get("F2D_TRACE", "1") == "1":
        cap = 4096
        tr = torch.zeros(cap, dtype=torch.int64, device="cuda")
        L.mg_set_trace(h, ptr(tr), cap)
        L.mg_set_graphs(h, 0)
        L.mg_vcycle(h, 0, s)
        torch.cuda.synchronize()
        t = tr.cpu().numpy()
        k = int(t[0])
        st = t[1:1+k]
        if k > 1:
            d = st[1:]-st[:-1]
            print("   barriers: %d, first to last %.1f us at 1.965 GHz" % (k, (st[-1]-st[0])/1965.))
            print("   cycles between barriers:", " ".join(str(int(v)) for v in d))
        L.mg_set_trace(h, None, 0)
    L.mg_destroy(h)
if big:
    X = torch.zeros((big+6, big+6), dtype=torch.float64, device="cuda")
    Bv = torch.randn((big+6, big+6), dtype=torch.float64, device="cuda")
    Bv -= Bv[3:-3, 3:-3].mean()
    L.fill_halo(ptr(Bv), 3, big+6, big+6, s)
    for name, env in CONFIGS:
        setenv(env)
        h = hierarchy(big)
        ms = timeit(lambda: L.mg_two_vcycle(h, ptr(X), ptr(Bv), s), reps=10)/2e3
        msf = timeit(lambda: L.mg_fcycle(h, 0, s), reps=5)/1e3
        print("%-30s %d^2: V-cycle %.4f ms, F-cycle %.4f ms" % (name, big, ms, msf))
        L.mg_destroy(h)
