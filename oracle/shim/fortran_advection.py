"""Stand-in for the reference's f2py module `fortran_advection`: forwards to the CPU oracle
(oracle/kernels.py).  Used ONLY to import the reference's Python in the build
container (tests/golden/make_golden.py).  Test infrastructure, not product."""
from oracle.kernels import fortran_advection as _impl

globals().update({k: v.__func__ if isinstance(v, staticmethod) else v
                  for k, v in vars(_impl).items() if not k.startswith('_')})
