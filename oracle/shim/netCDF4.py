"""Placeholder for the netCDF4 package (not installed in the build container), so that
reference modules which import it at the top (restart.py, plotting helpers) can be imported
by tests/golden/make_golden.py and tools/experiment_compat.py.  Opening a dataset fails
loudly.  Test infrastructure, not product."""


class Dataset(object):
    def __init__(self, *args, **kwargs):
        raise ImportError("netCDF4 is not installed here (oracle/shim/netCDF4.py is a placeholder)")
