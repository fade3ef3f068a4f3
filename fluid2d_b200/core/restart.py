"""Restart: run Fluid2d with restarts (reference: core/restart.py).

In a user script, `Restart(param, grid, f2d)` replaces `f2d.loop()`: if no restart exists in
the experiment directory the run starts from scratch, otherwise from the last restart found;
`param.tend` is then the LENGTH of integration of this job, not its final time
(restart.py:14-17).  The job is split in `param.ninterrestart` sub-intervals with a restart
written at the end of each (same file, overwritten during a job).  The history, diagnostics
and flux files of job NN carry the job index in their names (restart.py:75-89).

One file per rank, <expname>_<NN>_restart_<rank>: every variable of varname_list in double,
halos included, and tend, t, dt, kt, tnextdiag, tnexthis -- netCDF attributes / variables as
in the reference when netCDF4 is importable, the same record in an .npz otherwise.
"""
import glob
import os

import numpy as np

try:
    from netCDF4 import Dataset
    HAVE_NETCDF = True
except ImportError:
    HAVE_NETCDF = False


class Restart(object):
    def __init__(self, param, grid, f2d, launch=True):
        self.list_param = ['expname', 'expdir', 'nbproc', 'myrank', 'tend', 'varname_list', 'ninterrestart',
                           'diag_fluxes']
        param.copy(self, self.list_param)
        self.list_grid = ['nh', 'nxl', 'nyl']
        grid.copy(self, self.list_grid)
        self.f2d = f2d
        self.ext = 'nc' if HAVE_NETCDF else 'npz'
        self.template = self.expdir+'/%s_%02i_restart_%03i.'+self.ext
        self.get_lastrestart()
        self.timelength = f2d.tend
        if self.myrank == 0:
            print('-'*50)
        if self.lastrestart is not None:
            self.restart_file = self.template % (self.expname, self.lastrestart, self.myrank)
            if self.myrank == 0:
                print(' Restart found')
                print(' Restarting from %s' % self.restart_file)
            tend, t, dt, kt, tnextdiag, tnexthis = self.read(f2d.model.var)
            f2d.tend += round(t)
            f2d.t = t
            f2d.dt = dt
            f2d.kt = kt
            f2d.output.tnextdiag = tnextdiag
            f2d.output.tnexthis = tnexthis
            self.nextrestart = self.lastrestart+1
        else:
            if self.myrank == 0:
                print(' No restart')
                print(' Starting from scratch')
            self.nextrestart = 0
        # the output files of this job carry its index
        out = f2d.output
        oext = os.path.splitext(out.diagfile)[1]
        stem = self.expdir+'/%s_%02i' % (self.expname, self.nextrestart)
        out.diagfile = stem+'_diag'+oext
        out.template = stem+'_his_%03i'+oext
        out.hisfile = out.template % self.myrank
        out.hisfile_joined = stem+'_his'+oext
        if self.diag_fluxes:
            out.flxfile = (stem+'_flx_%03i'+oext) % self.myrank
            out.flxfile_joined = stem+'_flx'+oext
        self.lengthsubint = self.timelength/self.ninterrestart
        if launch:
            self.launchf2d(f2d)

    def launchf2d(self, f2d):
        """run, and write a restart at the end of each of the ninterrestart sub-intervals"""
        for kres in range(self.ninterrestart):
            f2d.tend = f2d.t+self.lengthsubint
            f2d.loop(joinhis=(kres == self.ninterrestart-1), keepplotalive=(kres < self.ninterrestart-1))
            self.restart_file = self.template % (self.expname, self.nextrestart, self.myrank)
            if self.myrank == 0:
                print('writing restart %i in %s' % (kres, self.restart_file))
            self.write(f2d.tend, f2d.t, f2d.dt, f2d.kt, f2d.output.tnextdiag, f2d.output.tnexthis, f2d.model.var)

    launch = launchf2d

    def get_lastrestart(self):
        """index of the last restart written by rank 0 (None: none yet)"""
        files = glob.glob(self.expdir+'/%s_*_restart_000.%s' % (self.expname, self.ext))
        self.lastrestart = None
        for f in files:
            pos = f.find('restart')
            idx = int(f[pos-3:pos-1])
            if self.lastrestart is None or idx > self.lastrestart:
                self.lastrestart = idx

    def write(self, tend, t, dt, kt, tnextdiag, tnexthis, var):
        state = var.dstate.numpy()          # one D2H copy of the whole state, halos included
        if HAVE_NETCDF:
            with Dataset(self.restart_file, 'w') as nc:
                for k, v in (('tend', tend), ('t', t), ('dt', dt), ('kt', kt), ('tnextdiag', tnextdiag),
                             ('tnexthis', tnexthis)):
                    nc.setncattr(k, v)
                nc.createDimension('x', self.nxl)
                nc.createDimension('y', self.nyl)
                for k, v in enumerate(self.varname_list):
                    nc.createVariable(v, 'd', ('y', 'x'))[:, :] = state[k]
        else:
            np.savez(self.restart_file, state=state, varnames=np.array(self.varname_list), tend=tend, t=t,
                     dt=dt, kt=kt, tnextdiag=tnextdiag, tnexthis=tnexthis)

    def read(self, var):
        if HAVE_NETCDF:
            with Dataset(self.restart_file, 'r') as nc:
                rec = [nc.getncattr(k) for k in ('tend', 't', 'dt', 'kt', 'tnextdiag', 'tnexthis')]
                state = np.stack([np.array(nc.variables[v][:, :]) for v in self.varname_list])
        else:
            d = np.load(self.restart_file)
            rec = [d[k][()] for k in ('tend', 't', 'dt', 'kt', 'tnextdiag', 'tnexthis')]
            state = d['state']
        var.dstate.upload_all_from(state)
        tend, t, dt, kt, tnextdiag, tnexthis = rec
        return float(tend), float(t), float(dt), int(kt), float(tnextdiag), float(tnexthis)
