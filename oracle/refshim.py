"""Import the reference's own Python (read-only at /root/reference) on top of the CPU
oracle kernels.  BUILD-CONTAINER ONLY: /root/reference does not exist on the GPU box,
so nothing under tests/ -m gpu, smoke() or bench.py may call install().

What is substituted, and why:
  mpi4py               -> oracle/shim/mpi4py      (not installed; single rank)
  fortran_* f2py mods  -> oracle/shim/fortran_*.py (no gfortran; C restatement)
  gmg.fortran_multigrid-> oracle.kernels.fortran_multigrid
  output               -> oracle/shim/output.py   (netCDF4 not installed)
  numpy.NaN            -> numpy.nan               (gmg/level.py imports the alias
                                                   numpy 2 removed)
Everything else (param, grid, variables, timescheme, operators, euler, boussinesq,
quasigeostrophic, island, fluid2d, gmg/level, gmg/hierarchy, gmg/halo ...) is the
reference's unmodified source.
"""
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "core", "gmg"))


def install(reference_root=REFERENCE_ROOT, kernels="oracle"):
    """kernels="oracle": the f2py modules are replaced by the C restatement (default);
    kernels="fortran_source": by the reference's own Fortran source, executed through
    oracle/fortran_source.py (slow: small grids only)"""
    import numpy
    use_source = kernels == "fortran_source"
    if not hasattr(numpy, "NaN"):
        numpy.NaN = numpy.nan
    here = os.path.dirname(os.path.abspath(__file__))
    repo = os.path.dirname(here)
    for p in (repo, os.path.join(here, "shim"), os.path.join(reference_root, "core")):
        if p in sys.path:
            sys.path.remove(p)
    # shim first so that its fortran_*/output/mpi4py win; then the reference core
    sys.path.insert(0, os.path.join(reference_root, "core"))
    sys.path.insert(0, os.path.join(here, "shim"))
    sys.path.insert(0, repo)
    from oracle import kernels
    mod = types.ModuleType("gmg.fortran_multigrid")
    if use_source:
        from oracle import fortran_source
        for key in ("fortran_advection", "fortran_fluxes", "fortran_operators", "fortran_diag"):
            m = types.ModuleType(key)
            m.__dict__.update(fortran_source.f2py_namespace(key))
            sys.modules[key] = m
        mod.__dict__.update(fortran_source.f2py_namespace("fortran_multigrid"))
    else:
        for k, v in vars(kernels.fortran_multigrid).items():
            if not k.startswith("_"):
                setattr(mod, k, v.__func__ if isinstance(v, staticmethod) else v)
    sys.modules["gmg.fortran_multigrid"] = mod
    import gmg  # the reference package (core/gmg/__init__.py)
    gmg.fortran_multigrid = mod
    return kernels
