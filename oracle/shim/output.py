"""Stand-in for the reference's core/output.py (netCDF4 is not installed in the build
container): keeps the reference's history / diagnostics CLOCKS (output.py:56-57, 85-91 --
Fluid2d.loop steers its time step by `tnexthis` when param.exacthistime is set,
fluid2d.py:243-255) and swallows the writes.  Used ONLY by tests/golden/make_golden.py and
tools/experiment_compat.py.  Test infrastructure, not product."""


class Output(object):
    def __init__(self, param, grid, diag, flxlist=None):
        self.hisfile = "none"
        self.diagfile = "none"
        self.flxfile = "none"
        self.freq_his = param.freq_his
        self.freq_diag = param.freq_diag
        self.tnexthis = 0.
        self.tnextdiag = 0.
        self.diags_log = []
        self.diag = diag

    def do(self, data, t, kt):
        if t >= self.tnextdiag:
            self.tnextdiag += self.freq_diag
            self.diags_log.append((kt, t, dict(self.diag)))
        if t >= self.tnexthis:
            self.tnexthis += self.freq_his

    def dump_diag(self):
        pass

    def join(self):
        pass
