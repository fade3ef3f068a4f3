#!/usr/bin/env python
"""Do the reference's experiment scripts run UNCHANGED on fluid2d_b200's Python API?

BUILD-CONTAINER ONLY (reads /root/reference/experiments; no GPU needed).  Each script is
executed twice, as it is (runpy, cwd = the script's directory, the same numpy seed), the time
loop cut after --nsteps iterations:

  reference : the reference's own Python on the oracle kernels (oracle/refshim.py)
  product   : fluid2d_b200's host layer on the CPU emulation of the C ABI
              (tests/emu_device.py: every entry point executed by the same oracle kernels)

and the final model states, clocks and diagnostics are compared.  Agreement says the host
layer accepts what the script does (parameters, masks, islands, forcing modules, custom
steps, state edits through var.get) and asks the device for the same operations in the same
order; it says nothing about the CUDA kernels (the -m gpu tests do).  Only the interactive
figure and the movie are switched off, and files go to a scratch directory.

    python tools/experiment_compat.py                 # every script, table on stdout
    python tools/experiment_compat.py --json out.json
    python tools/experiment_compat.py --one IMPL SCRIPT OUT.npz   (worker)
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXPERIMENTS = "/root/reference/experiments"


def worker(impl, script, out, nsteps):
    import numpy as np
    sys.path.insert(0, REPO)
    if impl == "reference":
        from oracle import refshim
        refshim.install()
    else:
        sys.path.insert(0, os.path.join(REPO, "tests"))
        import emu_device
        emu_device.install()
    import fluid2d as F
    scratch = tempfile.mkdtemp(prefix="f2d_compat_")
    made = []
    real_init = F.Fluid2d.__init__

    def init(self, param, grid, *a, **k):
        param.plot_interactive = False
        param.generate_mp4 = False
        param.datadir = scratch
        real_init(self, param, grid, *a, **k)
        made.append(self)
        model = self.model
        step = model.step
        count = [0]

        def counted(t, dt):
            step(t, dt)
            count[0] += 1
            if count[0] >= nsteps:
                self.stop = True       # the loop finishes this iteration and leaves
        model.step = counted
    F.Fluid2d.__init__ = init
    os.chdir(os.path.dirname(script))
    sys.path.insert(0, os.path.dirname(script))
    sys.argv = [script]
    np.random.seed(1)
    import runpy
    runpy.run_path(script, run_name="__main__")
    if not made:
        raise RuntimeError("the script never built a Fluid2d")
    f2d = made[-1]
    model = f2d.model
    diags = {k: float(np.ravel(v)[0]) for k, v in model.diags.items()}
    np.savez(out, varnames=np.array(model.var.varname_list), state=np.array(model.var.state, copy=True),
             t=float(f2d.t), kt=int(f2d.kt), dt=float(f2d.dt),
             diag_names=np.array(sorted(diags)), diag_values=np.array([diags[k] for k in sorted(diags)]))


def scripts():
    found = []
    for dirpath, _d, files in os.walk(EXPERIMENTS):
        for f in sorted(files):
            p = os.path.join(dirpath, f)
            if f.endswith(".py") and "Fluid2d(" in open(p).read():
                found.append(p)
    return sorted(found)


def run_one(impl, script, out, nsteps, timeout):
    cmd = [sys.executable, os.path.abspath(__file__), "--one", impl, script, out, "--nsteps", str(nsteps)]
    t0 = time.time()
    try:
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout, text=True)
    except subprocess.TimeoutExpired:
        return "timeout", time.time()-t0
    if p.returncode != 0 or not os.path.exists(out):
        lines = [ln for ln in (p.stderr or "").strip().splitlines() if ln.strip()]
        return "failed: " + (lines[-1] if lines else "exit %d" % p.returncode), time.time()-t0
    return "ok", time.time()-t0


def compare(a, b):
    import numpy as np
    A, B = np.load(a), np.load(b)
    if list(A["varnames"]) != list(B["varnames"]):
        return "variables differ", None
    if int(A["kt"]) != int(B["kt"]):
        return "kt %d vs %d" % (int(A["kt"]), int(B["kt"])), None
    worst = 0.
    for k in range(len(A["varnames"])):
        n = np.linalg.norm(A["state"][k])
        worst = max(worst, np.linalg.norm(A["state"][k]-B["state"][k])/(n if n > 0 else 1.))
    same_diags = (list(A["diag_names"]) == list(B["diag_names"])
                  and np.array_equal(A["diag_values"], B["diag_values"], equal_nan=True))
    exact = (np.array_equal(A["state"], B["state"]) and float(A["t"]) == float(B["t"]) and same_diags)
    return ("bit-identical" if exact else "rel L2 %.1e%s" % (worst, "" if same_diags else ", diags differ")), worst


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--one", nargs=3, metavar=("IMPL", "SCRIPT", "OUT"))
    ap.add_argument("--nsteps", type=int, default=3)
    ap.add_argument("--timeout", type=float, default=600.)
    ap.add_argument("--json")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    if args.one:
        worker(args.one[0], args.one[1], args.one[2], args.nsteps)
        return
    from concurrent.futures import ThreadPoolExecutor
    tmp = tempfile.mkdtemp(prefix="f2d_compat_out_")
    todo = [s for s in scripts() if args.only in s]

    def both(s):
        tag = os.path.relpath(s, EXPERIMENTS).replace("/", "_")[:-3]
        r = run_one("reference", s, os.path.join(tmp, tag+"_ref.npz"), args.nsteps, args.timeout)
        p = run_one("product", s, os.path.join(tmp, tag+"_prod.npz"), args.nsteps, args.timeout)
        verdict = None
        if r[0] == "ok" and p[0] == "ok":
            verdict = compare(os.path.join(tmp, tag+"_ref.npz"), os.path.join(tmp, tag+"_prod.npz"))[0]
        return {"script": os.path.relpath(s, EXPERIMENTS), "reference": r[0], "product": p[0],
                "agreement": verdict, "seconds": round(r[1]+p[1], 1)}
    with ThreadPoolExecutor(max_workers=max(1, (os.cpu_count() or 2)//2)) as ex:
        rows = list(ex.map(both, todo))
    for r in rows:
        print("%-52s ref: %-40s product: %-40s %s" % (r["script"], r["reference"][:40], r["product"][:40],
                                                      r["agreement"] or "-"))
    if args.json:
        json.dump({"nsteps": args.nsteps, "rows": rows}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
