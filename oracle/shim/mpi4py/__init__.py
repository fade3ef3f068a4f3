"""Single-rank stand-in for mpi4py, used ONLY to import the reference's Python in the
build container (tests/golden/make_golden.py).  Test infrastructure, not product."""
from . import MPI  # noqa: F401
