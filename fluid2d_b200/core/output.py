"""Output: history snapshots and integral diagnostics of a run.

Same role and call pattern as the reference's core/output.py (Output(param, grid, diag,
flxlist), do(data, t, kt), tnexthis / tnextdiag, dump_diag(), join(), hisfile, diagfile):

* every rank creates its own history (and flux) file on the first do() and APPENDS one record
  at each `tnexthis`, inside do() -- a run that is killed keeps what it has written, and host
  memory holds one snapshot at a time (output.py:77-95, NcfileIO.write);
* rank 0 creates the diagnostics file and flushes it every 10 records (output.py:118-152);
  dump_diag() writes what is left in the buffer.

A snapshot costs one device-to-host copy of the fields in var_to_save (float32 on disk, like
the reference's history files; cast and packed on the device).  netCDF4 is used when it is
installed (unlimited 't' dimension).  Otherwise the records go to an uncompressed .npz that
grows member by member (a ZIP archive opened in append mode: 't/000003.npy',
'vorticity/000003.npy', ...); load_records() reads either layout back as stacked arrays.
"""
import io
import os
import zipfile

import numpy as np

try:
    import netCDF4  # noqa: F401
    HAVE_NETCDF = True
except Exception:
    HAVE_NETCDF = False


def load_records(path):
    """{name: array} of a history / flux / diagnostics file written by Output: the record
    variables stacked along a leading time axis, the static ones (x, y, msk) as they are"""
    if not path.endswith('.npz'):
        from netCDF4 import Dataset
        with Dataset(path) as nc:
            return {k: np.array(v[:]) for k, v in nc.variables.items()}
    out, recs = {}, {}
    with np.load(path) as z:
        for key in z.files:
            if '/' in key:
                name, idx = key.rsplit('/', 1)
                recs.setdefault(name, []).append((int(idx), z[key]))
            else:
                out[key] = z[key]
    for name, items in recs.items():
        items.sort(key=lambda it: it[0])
        arrs = [a for _, a in items]
        # a diagnostics member is a block of records (1-D): concatenate; a snapshot member (2-D)
        # or a scalar time (0-D) is one record: stack
        out[name] = np.concatenate(arrs) if arrs[0].ndim == 1 else np.stack(arrs)
    return out


class RecordFile(object):
    """append-only file of records along an unlimited 't' axis (NcfileIO of the reference)"""

    def __init__(self, path, names, shape=None, static=None, attrs=None, dtype='f'):
        self.path, self.names, self.shape = path, list(names), shape
        self.static = static or {}
        self.attrs = attrs or {}
        self.dtype = dtype
        self.nrec = 0
        self.created = False

    def create(self):
        d = os.path.dirname(self.path)
        if d and not os.path.isdir(d):
            os.makedirs(d, exist_ok=True)
        if HAVE_NETCDF:
            from netCDF4 import Dataset
            with Dataset(self.path, 'w', format='NETCDF4') as nc:
                for k, v in self.attrs.items():
                    nc.setncattr(k, v*1 if isinstance(v, bool) else v)
                nc.createDimension('t', None)
                dims = ('t',)
                if self.shape is not None:
                    nc.createDimension('y', self.shape[0])
                    nc.createDimension('x', self.shape[1])
                    dims = ('t', 'y', 'x')
                for k, v in self.static.items():
                    v = np.asarray(v)
                    sd = {1: ('x',) if k == 'x' else ('y',), 2: ('y', 'x')}[v.ndim]
                    nc.createVariable(k, 'i' if v.dtype.kind in 'iu' else 'f', sd)[:] = v
                nc.createVariable('t', 'f', ('t',))
                if 'kt' in self.names:
                    nc.createVariable('kt', 'i', ('t',))
                for v in self.names:
                    if v not in ('t', 'kt'):
                        nc.createVariable(v, self.dtype, dims)
        else:
            with zipfile.ZipFile(self.path, 'w', zipfile.ZIP_STORED) as z:
                for k, v in self.static.items():
                    self._put(z, k, np.asarray(v))
        self.created = True

    @staticmethod
    def _put(z, key, arr):
        buf = io.BytesIO()
        np.lib.format.write_array(buf, np.asanyarray(arr), allow_pickle=False)
        z.writestr(key+'.npy', buf.getvalue())

    def append(self, rec, count=1):
        """rec: {name: array}; count > 1: a block of `count` records (1-D arrays of that length)"""
        if not self.created:
            self.create()
        if HAVE_NETCDF:
            from netCDF4 import Dataset
            with Dataset(self.path, 'r+') as nc:
                k = slice(self.nrec, self.nrec+count) if count > 1 else self.nrec
                for name, arr in rec.items():
                    nc.variables[name][k] = arr
        else:
            with zipfile.ZipFile(self.path, 'a', zipfile.ZIP_STORED) as z:
                for name, arr in rec.items():
                    self._put(z, '%s/%06i' % (name, self.nrec), arr)
        self.nrec += count


class Output(object):
    def __init__(self, param, grid, diag, flxlist=None):
        self.list_param = ['expname', 'myrank', 'nh', 'nprint', 'var_to_save', 'varname_list', 'expdir',
                           'tracer_list', 'diag_fluxes', 'freq_his', 'freq_diag', 'list_diag']
        param.copy(self, self.list_param)
        self.grid = grid
        self.diag = diag
        self.nbproc = getattr(param, 'nbproc', 1)
        # output.py:33-39
        if self.var_to_save == 'all':
            self.var_to_save = [v for v in self.varname_list]
        if type(self.var_to_save) == str:
            self.var_to_save = [self.var_to_save]
        for v in self.var_to_save:
            if v not in self.varname_list:
                raise ValueError('%s is not a model variable' % v + ' => modify param.var_to_save')
        ext = 'nc' if HAVE_NETCDF else 'npz'
        self.template = self.expdir+'/%s_his' % self.expname + '_%03i.' + ext
        if self.nbproc > 1:
            self.hisfile = self.template % self.myrank
            self.hisfile_joined = '%s/%s_his.%s' % (self.expdir, self.expname, ext)
            self.flxfile = (self.expdir+'/%s_flx' % self.expname + '_%03i.' + ext) % self.myrank
            self.flxfile_joined = '%s/%s_flx.%s' % (self.expdir, self.expname, ext)
        else:
            self.hisfile = '%s/%s_his.%s' % (self.expdir, self.expname, ext)
            self.flxfile = '%s/%s_flx.%s' % (self.expdir, self.expname, ext)
        self.diagfile = '%s/%s_diag.%s' % (self.expdir, self.expname, ext)
        self.tnextdiag = 0.
        self.tnexthis = 0.
        self.first = True
        self.flxlist = flxlist if self.diag_fluxes else None
        self.kdiag = 0
        self.buffersize = 10          # diagnostics records kept in memory between two flushes (output.py:116)
        self.param_attrs = {k: v for k, v in param.__dict__.items()
                            if isinstance(v, (int, float, str, bool))}

    # ------------------------------------------------------------------ files
    def _create_files(self):
        g, nh = self.grid, self.nh
        shape = (g.nyl-2*nh, g.nxl-2*nh)
        static = {'x': g.x1d[nh:-nh], 'y': g.y1d[nh:-nh], 'msk': np.asarray(g.msk)[nh:-nh, nh:-nh]}
        self.nchis = RecordFile(self.hisfile, ['t']+list(self.var_to_save), shape, static, self.param_attrs)
        self.nchis.create()
        if self.flxlist:
            self.ncflx = RecordFile(self.flxfile, ['t']+list(self.flxlist), shape, static, self.param_attrs)
            self.ncflx.create()
        if self.myrank == 0:
            if self.list_diag == 'all':
                self.list_diag = list(self.diag.keys())
            self.list_diag = [k for k in self.list_diag if k in self.diag]
            self.ncdiag = RecordFile(self.diagfile, ['t', 'kt']+list(self.list_diag))
            self.ncdiag.create()
            self.buffer = np.zeros((self.buffersize, len(self.list_diag)+2))

    def do(self, data, t, kt):
        """data['his'] is the model Var (device state); data['diag'] the diags dict;
        data['flx'] the Fluxes driver"""
        if self.first:
            self.first = False
            self._create_files()
        if t >= self.tnextdiag:
            self.tnextdiag += self.freq_diag
            if self.myrank == 0:
                self._write_diag(t, kt)
        if t >= self.tnexthis:
            self.tnexthis += self.freq_his
            var = data['his']
            rec = {'t': np.float32(t)}
            for v in self.var_to_save:
                rec[v] = self._snapshot(var.dstate, var.index(v))
            self.nchis.append(rec)
            if self.flxlist:
                fstate = data['flx']._flx     # the flux stack (output.py:94-95)
                rec = {'t': np.float32(t)}
                for k, v in enumerate(self.flxlist):
                    rec[v] = self._snapshot(fstate, k)
                self.ncflx.append(rec)

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

    # ------------------------------------------------------------------ diagnostics (rank 0)
    def _write_diag(self, t, kt):
        k = self.kdiag % self.buffersize
        self.buffer[k, 0] = t
        self.buffer[k, 1] = kt
        for j, name in enumerate(self.list_diag):
            self.buffer[k, 2+j] = float(np.asarray(self.diag[name]).ravel()[0])
        self.kdiag += 1
        if self.kdiag % self.buffersize == 0:
            self.dump_diag()

    def dump_diag(self):
        """write the diagnostics records still in the buffer (output.py:140-152)"""
        if self.first or self.myrank != 0:
            return
        n = self.kdiag-self.ncdiag.nrec
        if n <= 0:
            return
        i0 = self.ncdiag.nrec % self.buffersize      # the unflushed records are contiguous in the buffer
        rows = self.buffer[i0:i0+n]
        rec = {'t': rows[:, 0].astype(np.float32), 'kt': rows[:, 1].astype(np.int32)}
        for j, name in enumerate(self.list_diag):
            rec[name] = rows[:, 2+j].astype(np.float32)
        self.ncdiag.append(rec, count=n)

    def join(self):
        """per-rank history files are left as they are (one file per slab; the reference joins
        them with a netCDF tool when nbproc <= 64)"""
        pass
