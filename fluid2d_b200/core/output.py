"""Output: history snapshots and integral diagnostics of a run.

Same role and call pattern as the reference's core/output.py (Output(param, grid, diag,
flxlist), do(data, t, kt), tnexthis / tnextdiag, dump_diag(), join(), hisfile,
diagfile).  A snapshot costs one device-to-host copy of the fields in var_to_save (float32
on disk, like the reference's history files).  netCDF4 is used when it is installed;
otherwise the same records go to <expname>_his.npz / <expname>_diag.npz.
"""
import os

import numpy as np

try:
    import netCDF4  # noqa: F401
    HAVE_NETCDF = True
except Exception:
    HAVE_NETCDF = False


class Output(object):
    def __init__(self, param, grid, diag, flxlist=None):
        self.list_param = ['expname', 'myrank', 'nh', 'nprint', 'var_to_save', 'varname_list', 'expdir',
                           'tracer_list', 'diag_fluxes', 'freq_his', 'freq_diag', 'list_diag']
        param.copy(self, self.list_param)
        self.grid = grid
        self.diag = diag
        if type(self.var_to_save) == str:
            self.var_to_save = [self.var_to_save]
        self.var_to_save = [v for v in self.var_to_save if v in self.varname_list]
        ext = 'nc' if HAVE_NETCDF else 'npz'
        self.template = self.expdir+'/%s_his' % self.expname + '_%03i.' + ext
        if param.nbproc > 1:
            self.hisfile = self.template % self.myrank
        else:
            self.hisfile = '%s/%s_his.%s' % (self.expdir, self.expname, ext)
        self.diagfile = '%s/%s_diag.%s' % (self.expdir, self.expname, ext)
        self.flxfile = '%s/%s_flx.%s' % (self.expdir, self.expname, ext)
        self.tnextdiag = 0.
        self.tnexthis = 0.
        self.first = True
        self.his_t = []
        self.his = {v: [] for v in self.var_to_save}
        self.flxlist = flxlist if self.diag_fluxes else None
        self.flx = {v: [] for v in (self.flxlist or [])}
        self.diag_t = []
        self.diag_kt = []
        self.diag_rec = {}
        self.param_attrs = {k: v for k, v in param.__dict__.items()
                            if isinstance(v, (int, float, str, bool))}

    def do(self, data, t, kt):
        """data['his'] is the model Var (device state); data['diag'] the diags dict"""
        if self.first:
            self.first = False
            if self.list_diag == 'all':
                self.list_diag = list(self.diag.keys())
        if t >= self.tnextdiag:
            self.tnextdiag += self.freq_diag
            self.diag_t.append(t)
            self.diag_kt.append(kt)
            for k in self.list_diag:
                if k in self.diag:
                    self.diag_rec.setdefault(k, []).append(float(np.asarray(self.diag[k]).ravel()[0]))
        if t >= self.tnexthis:
            self.tnexthis += self.freq_his
            var = data['his']
            nh = self.nh
            self.his_t.append(t)
            for v in self.var_to_save:
                self.his[v].append(self._snapshot(var.dstate, var.index(v)))
            if self.flxlist:
                fstate = data['flx']._flx     # the flux stack (output.py:94-95)
                for k, v in enumerate(self.flxlist):
                    self.flx[v].append(self._snapshot(fstate, k))

    def _snapshot(self, dstate, k):
        """interior of field k of a DeviceState as float32: cast and packed on the device
        (f2d_pack_interior_f32), one D2H copy of half the fp64 bytes"""
        import torch
        from runtime import rt
        r = rt()
        nh, ny, nx = self.nh, dstate.ny, dstate.nx
        dev = torch.empty((ny-2*nh, nx-2*nh), dtype=torch.float32, device=r.device)
        r.lib.pack_interior_f32(dstate.rptr(k), r.ptr(dev), nh, ny, nx, r.stream)
        return dev.cpu().numpy()

    def dump_diag(self):
        self._write_diag()
        self._write_his()
        if self.flxlist:
            self._write_his(self.flxfile, self.flxlist, self.flx)

    def _write_diag(self):
        if HAVE_NETCDF:
            from netCDF4 import Dataset
            with Dataset(self.diagfile, 'w') as nc:
                nc.createDimension('t', None)
                nc.createVariable('t', 'f', ('t',))[:] = np.array(self.diag_t)
                nc.createVariable('kt', 'i', ('t',))[:] = np.array(self.diag_kt)
                for k, v in self.diag_rec.items():
                    nc.createVariable(k, 'f', ('t',))[:] = np.array(v)
        else:
            np.savez(self.diagfile, t=np.array(self.diag_t), kt=np.array(self.diag_kt),
                     **{k: np.array(v) for k, v in self.diag_rec.items()})

    def _write_his(self, hisfile=None, names=None, rec=None):
        """history-type file (output.py NcfileIO): the model snapshots, or the flux stack"""
        if not self.his_t:
            return
        g, nh = self.grid, self.nh
        hisfile = hisfile or self.hisfile
        names = names if names is not None else self.var_to_save
        rec = rec if rec is not None else self.his
        if HAVE_NETCDF:
            from netCDF4 import Dataset
            with Dataset(hisfile, 'w') as nc:
                for k, v in self.param_attrs.items():
                    nc.setncattr(k, v*1 if isinstance(v, bool) else v)
                nc.createDimension('t', None)
                nc.createDimension('x', g.nxl-2*nh)
                nc.createDimension('y', g.nyl-2*nh)
                nc.createVariable('x', 'f', ('x',))[:] = g.x1d[nh:-nh]
                nc.createVariable('y', 'f', ('y',))[:] = g.y1d[nh:-nh]
                nc.createVariable('msk', 'i', ('y', 'x'))[:] = g.msk[nh:-nh, nh:-nh]
                nc.createVariable('t', 'f', ('t',))[:] = np.array(self.his_t)
                for v in names:
                    nc.createVariable(v, 'f', ('t', 'y', 'x'))[:] = np.stack(rec[v])
        else:
            np.savez(hisfile, t=np.array(self.his_t), x=g.x1d[nh:-nh], y=g.y1d[nh:-nh],
                     msk=g.msk[nh:-nh, nh:-nh], **{v: np.stack(rec[v]) for v in names})

    def join(self):
        """per-rank history files are left as they are (one file per slab)"""
        pass
