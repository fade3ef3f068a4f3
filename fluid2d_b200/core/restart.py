"""Restart: save / reload the model state and clock (reference: core/restart.py).
One file per rank, <expname>_<NN>_restart_<rank>.npz (.nc when netCDF4 is present is not
needed for a round trip; the npz holds the same record: every variable of
varname_list in double including halos, and tend, t, dt, kt, tnextdiag, tnexthis)."""
import glob
import os

import numpy as np


class Restart(object):
    def __init__(self, param, grid, f2d, launch=True):
        self.list_param = ['expname', 'expdir', 'myrank', 'tend', 'varname_list', 'ninterrestart']
        param.copy(self, self.list_param)
        self.f2d = f2d
        self.template = self.expdir+'/%s_%02i_restart' % (self.expname, 0)
        self.timelength = self.tend
        self.lastrestart = self._latest()
        if self.lastrestart is not None:
            self.read(self.lastrestart)
            f2d.tend += f2d.t
        if launch:
            self.launch()

    def _files(self):
        return sorted(glob.glob(self.expdir+'/%s_*_restart_%03i.npz' % (self.expname, self.myrank)))

    def _latest(self):
        files = self._files()
        if not files:
            return None
        return int(os.path.basename(files[-1]).split('_')[-3])

    def launch(self):
        f2d = self.f2d
        start = 0 if self.lastrestart is None else self.lastrestart+1
        t0 = f2d.t
        for k in range(self.ninterrestart):
            f2d.tend = t0+(k+1)*self.timelength/self.ninterrestart
            f2d.loop(joinhis=(k == self.ninterrestart-1))
            self.write(start+k)

    def write(self, idx):
        f2d = self.f2d
        fname = self.expdir+'/%s_%02i_restart_%03i.npz' % (self.expname, idx, self.myrank)
        state = f2d.model.var.dstate.numpy()
        np.savez(fname, state=state, varnames=np.array(self.varname_list), tend=f2d.tend, t=f2d.t,
                 dt=f2d.dt, kt=f2d.kt, tnextdiag=f2d.output.tnextdiag, tnexthis=f2d.output.tnexthis)

    def read(self, idx):
        f2d = self.f2d
        fname = self.expdir+'/%s_%02i_restart_%03i.npz' % (self.expname, idx, self.myrank)
        d = np.load(fname)
        f2d.model.var.dstate.upload_all_from(d['state'])
        f2d.t, f2d.dt, f2d.kt = float(d['t']), float(d['dt']), int(d['kt'])
        f2d.output.tnextdiag, f2d.output.tnexthis = float(d['tnextdiag']), float(d['tnexthis'])
        return f2d.t, f2d.dt, f2d.kt, f2d.output.tnextdiag, f2d.output.tnexthis
