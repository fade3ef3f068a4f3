"""Stand-in for the reference's core/output.py (netCDF4 is not installed in the build
container): swallows the history / diagnostics writes.  Used ONLY by
tests/golden/make_golden.py.  Test infrastructure, not product."""


class Output(object):
    def __init__(self, param, grid, diag, flxlist=None):
        self.hisfile = "none"
        self.diagfile = "none"
        self.flxfile = "none"
        self.tnexthis = 1e30
        self.tnextdiag = 1e30
        self.diags_log = []
        self.diag = diag

    def do(self, data, t, kt):
        self.diags_log.append((kt, t, dict(self.diag)))

    def dump_diag(self):
        pass

    def join(self):
        pass
