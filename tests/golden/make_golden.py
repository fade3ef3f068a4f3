"""Generate the golden fixtures tests/golden/*.npz.

BUILD-CONTAINER ONLY (needs /root/reference).  For every case of cases.CASES it runs
the REFERENCE's own Python -- Param, Grid, Fluid2d, Euler/Boussinesq/QG, Timescheme,
Operators, gmg.Gmg -- with the reference's f2py kernels replaced by the CPU oracle
(oracle/refshim.py), and freezes:

  state0            full model state handed to the time loop
  state{k}, t{k}, dt{k}, diag{k}_<name>   after k = 1 and 10 iterations of the loop body
  mg_nlevs, mg_msk{l}, mg_A{l}            multigrid masks and 5-diagonal matrices
  mg_shape{l}                             (m, n) interior size per level
  flxnames, flx0, flx10                   (diag_fluxes cases) core/fluxes.py's stack of reversible /
                                          irreversible fluxes at the initial state and after 10 steps
  varnames, grid_msk

Usage:  python tests/golden/make_golden.py [case ...]
"""
import io
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle import refshim  # noqa: E402


def reference_api(kernels="oracle"):
    refshim.install(kernels=kernels)
    import types
    from param import Param
    from grid import Grid
    from fluid2d import Fluid2d
    return types.SimpleNamespace(Param=Param, Grid=Grid, Fluid2d=Fluid2d, name="reference")


def generate(names=None, kernels="oracle", check=False, nsteps=None):
    """check=True: nothing is written; the run is compared with the committed fixture bit for
    bit (used with kernels="fortran_source": the reference's Python on the reference's own
    Fortran source must reproduce what it gave on the C restatement); nsteps limits the run"""
    import numpy as np
    sys.path.insert(0, HERE)
    import cases
    api = reference_api(kernels)
    datadir = tempfile.mkdtemp(prefix="f2d_golden_")
    real_stdout = sys.stdout
    for name, builder in cases.CASES.items():
        if names and name not in names:
            continue
        sys.stdout = io.StringIO()
        try:
            f2d = builder(api, datadir)
            log = sys.stdout
        finally:
            pass
        sys.stdout = io.StringIO()   # Fluid2d installs a tee Logger; silence it
        model = f2d.model
        out = {}
        out["varnames"] = np.array(model.var.varname_list)
        out["state0"] = np.array(model.var.state, copy=True)
        out["grid_msk"] = np.array(f2d.msk if hasattr(f2d, "msk") else model.msk, copy=True)
        gmg = model.ope.gmg
        out["mg_nlevs"] = np.array(gmg.nlevs)
        for lev in range(0 if name in cases.LIGHT else gmg.nlevs):
            g = gmg.grid[lev]
            out["mg_msk%i" % lev] = np.array(g.msk, dtype=np.int8, copy=True)
            out["mg_A%i" % lev] = np.array(g.A, copy=True)
            out["mg_shape%i" % lev] = np.array([g.m, g.n])
        if getattr(f2d, "diag_fluxes", False):
            out["flxnames"] = np.array(f2d.flx.fullflx_list)
            out["flx0"] = cases.run_fluxes(f2d)
        res = cases.run_steps(f2d) if nsteps is None else cases.run_steps(f2d, (nsteps,))
        for k, (state, t, dt, diags) in res.items():
            out["state%i" % k] = state
            out["t%i" % k] = np.array(t)
            out["dt%i" % k] = np.array(dt)
            for dn, dv in diags.items():
                out["diag%i_%s" % (k, dn)] = np.array(dv)
        if getattr(f2d, "diag_fluxes", False) and nsteps is None:
            out["flx10"] = cases.run_fluxes(f2d)
        sys.stdout = real_stdout
        path = os.path.join(HERE, name + ".npz")
        if check:
            gold = np.load(path)
            bad = [k for k, v in out.items() if k in gold.files and not np.array_equal(np.asarray(v), gold[k])]
            missing = [k for k in out if k not in gold.files]
            print("%-28s %s kernels: %d records compared with the fixture: %s" % (
                name, kernels, len(out)-len(missing), "IDENTICAL" if not bad else "DIFFERENT: %s" % bad))
            if bad:
                sys.exit(1)
            continue
        np.savez_compressed(path, **out)
        print("%-28s -> %s (%.0f KB)  t10=%.6g  maxspeed=%.6g" % (
            name, os.path.relpath(path, REPO), os.path.getsize(path)/1024.,
            float(out["t10"]), float(out["diag10_maxspeed"])))
    sys.stdout = real_stdout


if __name__ == "__main__":
    if not refshim.available():
        sys.exit("make_golden.py needs the reference at /root/reference (build container only)")
    argv = sys.argv[1:]
    opts = {"kernels": "oracle", "check": False, "nsteps": None}
    names = []
    while argv:
        a = argv.pop(0)
        if a == "--kernels":
            opts["kernels"] = argv.pop(0)
        elif a == "--check":
            opts["check"] = True
        elif a == "--steps":
            opts["nsteps"] = int(argv.pop(0))
        else:
            names.append(a)
    generate(names, **opts)
