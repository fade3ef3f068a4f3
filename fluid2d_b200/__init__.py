"""fluid2d_b200: a B200 (sm_100a) implementation of Fluid2d's per-timestep hot path
behind the reference's Python API.

    import fluid2d_b200
    fluid2d_b200.activate()        # puts the flat modules of fluid2d_b200/core on sys.path
    from fluid2d import Fluid2d    # ... exactly like a reference experiment script
    from param import Param
    from grid import Grid

(equivalently: PYTHONPATH=<repo>/fluid2d_b200/core python my_experiment.py, the analogue
of the reference's `source ~/.fluid2d/activate.sh`).
"""
import os
import sys

CORE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "core")


def activate():
    """make `from fluid2d import Fluid2d`, `from param import Param`, ... resolve to this package"""
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (repo, CORE):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, repo)
    sys.path.insert(0, CORE)


def api():
    """namespace with the three classes an experiment script starts from"""
    import types
    activate()
    from param import Param
    from grid import Grid
    from fluid2d import Fluid2d
    return types.SimpleNamespace(Param=Param, Grid=Grid, Fluid2d=Fluid2d, name="fluid2d_b200")
