"""Device-resident model state with a lazily synchronised host mirror.

The reference keeps the model state in one numpy array [nvar, ny, nx] that user
scripts mutate in place through the views returned by Var.get() (variables.py:27-30,
docs/howto.rst).  Here the authoritative copy lives in HBM; the host sees it through a
pinned mirror:

  * DeviceState.host_view(k) / state[k] hand out TrackedArray views of the mirror.
    A TrackedArray tells its DeviceState when it is written (slice assignment, in-place
    operators, ufuncs with out=) so that the field is uploaded before the next kernel
    reads it, and refreshes itself from the device before it is read through numpy
    indexing or ufuncs when the device copy is newer.
  * kernels get raw device pointers through rptr()/wptr(); wptr() marks the host
    mirror of that field stale.

Nothing here computes: it only moves bytes (cudaMemcpy through torch).
"""
import ctypes

import numpy as np
import torch


class _OperandStage(object):
    """Device twins of the plain host arrays that user hooks combine with a state field in place
    (`dxdt[4] += self.forc`).  The host array is page-locked where it lies (cudaHostRegister, once
    per array) and copied by DMA on a side stream every time it is used -- its current contents,
    as numpy would read them -- while the host waits for that one copy only, not for the kernels
    queued on the compute stream."""

    MAX = 8

    def __init__(self):
        self.entries = {}      # (address, nbytes) -> [host array (kept alive), cpu tensor, device tensor, busy event]
        self.stream = None

    def get(self, other, device):
        key = (other.ctypes.data, other.nbytes)
        e = self.entries.get(key)
        if e is None or e[0] is not other:
            if e is not None:
                self._drop(key)
            if len(self.entries) >= self.MAX:
                self._drop(next(iter(self.entries)))
            rt = torch.cuda.cudart()
            if int(rt.cudaHostRegister(other.ctypes.data, other.nbytes, 0)) != 0:
                return None
            e = [other, torch.from_numpy(other), torch.empty(other.shape, dtype=torch.float64, device=device), None]
            self.entries[key] = e
        if self.stream is None:
            self.stream = torch.cuda.Stream(device=device)
        if e[3] is not None:
            self.stream.wait_event(e[3])           # the kernel that read the twin last has finished
        with torch.cuda.stream(self.stream):
            e[2].copy_(e[1], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        ev.synchronize()                           # the DMA has read the user's array: it may change again
        torch.cuda.current_stream().wait_event(ev)
        return e

    def _drop(self, key):
        e = self.entries.pop(key)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaHostUnregister(e[0].ctypes.data)


_operands = _OperandStage()


class TrackedArray(np.ndarray):
    """numpy view of the pinned mirror that reports writes to / pulls reads from HBM"""

    _own = None   # (DeviceState, field index or None for all)

    def __array_finalize__(self, obj):
        if obj is not None:
            self._own = getattr(obj, '_own', None)

    # -- coherence hooks ---------------------------------------------------
    def _before_read(self):
        if self._own is not None:
            self._own[0]._refresh_host(self._own[1])

    def _after_write(self):
        if self._own is not None:
            self._own[0]._host_written(self._own[1])

    def __getitem__(self, key):
        self._before_read()
        return np.ndarray.__getitem__(self, key)

    def __setitem__(self, key, value):
        self._before_read()          # partial writes must land on fresh data
        if isinstance(value, TrackedArray):
            value._before_read()
        np.ndarray.__setitem__(self, key, value)
        self._after_write()

    def _device_inplace(self, ufunc, method, inputs, out, kwargs):
        """`view += a`, `view -= a`, `view *= a` on a WHOLE field whose device copy is current
        (the idiom of user forcing hooks, e.g. `dxdt[4] += self.forc; dxdt[4] *= coef`): done
        by one kernel on the device instead of pulling the field to the host, doing the
        arithmetic there and pushing it back.  Same IEEE result as numpy (one rounding per
        element).  Returns False when the pattern does not apply."""
        if method != '__call__' or kwargs or out is None or len(out) != 1 or out[0] is not self:
            return False
        if ufunc not in (np.add, np.subtract, np.multiply) or len(inputs) != 2 or inputs[0] is not self:
            return False
        own = self._own
        if own is None or own[1] is None:
            return False
        st, k = own
        if self.shape != (st.ny, st.nx) or not st.dev_fresh[k] or st.device.type != 'cuda':
            return False
        if not self.flags['C_CONTIGUOUS'] or self.ctypes.data != st._host[k].ctypes.data:
            return False      # a reversed / strided view of the field: host path
        other = inputs[1]
        from runtime import rt
        r = rt()
        n = st.ny*st.nx
        if isinstance(other, (int, float, np.floating, np.integer)):
            c = float(other)
            if ufunc is np.multiply:
                r.lib.scale(st.wptr(k), c, n, r.stream)
            else:
                return False
            return True
        if isinstance(other, TrackedArray) or not isinstance(other, np.ndarray):
            return False
        if other.shape != self.shape or other.dtype != np.float64:
            return False
        staged = _operands.get(other, st.device) if other.flags['C_CONTIGUOUS'] and other.flags['OWNDATA'] else None
        d = staged[2] if staged is not None else torch.from_numpy(np.ascontiguousarray(other)).to(st.device)
        if ufunc is np.add:
            r.lib.add_scaled(st.wptr(k), 1., r.ptr(d), n, r.stream)
        elif ufunc is np.subtract:
            r.lib.add_scaled(st.wptr(k), -1., r.ptr(d), n, r.stream)
        else:
            r.lib.mul_field(st.wptr(k), r.ptr(d), n, r.stream)
        if staged is not None:
            staged[3] = torch.cuda.Event()
            staged[3].record()                      # the twin is busy until this kernel has run
        else:
            torch.cuda.current_stream().synchronize()   # `d` may be freed once we return
        return True

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        if self._device_inplace(ufunc, method, inputs, out, kwargs):
            return self
        plain = []
        for x in inputs:
            if isinstance(x, TrackedArray):
                x._before_read()
                plain.append(x.view(np.ndarray))
            else:
                plain.append(x)
        written = []
        if out is not None:
            pout = []
            for o in out:
                if isinstance(o, TrackedArray):
                    o._before_read()
                    written.append(o)
                    pout.append(o.view(np.ndarray))
                else:
                    pout.append(o)
            kwargs['out'] = tuple(pout)
        res = getattr(ufunc, method)(*plain, **kwargs)
        for o in written:
            o._after_write()
        if out is not None:
            return out[0] if len(out) == 1 else out
        return res

    def fill(self, value):
        if self._own is not None:
            self._own[0]._wait_upload()
        np.ndarray.fill(self, value)
        self._after_write()

    def copy(self, order='C'):
        self._before_read()
        return np.array(self.view(np.ndarray), order=order, copy=True)


class HostField(np.ndarray):
    """Host array with a device twin that is uploaded on demand.

    For model constants the reference keeps as plain numpy attributes and experiment scripts
    edit in place after the model is built (`model.bref[:, :] = buoy`, Internal_IVP /
    KelvinHelmholtz / Leewave).  Any write through the array or a view of it marks the root
    stale; `device_ptr()` uploads again before handing out the device address."""

    _root = None
    _stale = True
    _dev = None

    def __new__(cls, array):
        obj = np.array(array, dtype=np.float64, order='C', copy=True).view(cls)
        obj._root = None
        obj._stale = True
        obj._dev = None
        return obj

    def __array_finalize__(self, obj):
        if isinstance(obj, HostField):
            self._root = obj if obj._root is None else obj._root

    def _touch(self):
        (self if self._root is None else self._root)._stale = True

    def __setitem__(self, key, value):
        np.ndarray.__setitem__(self, key, value)
        self._touch()

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        plain = [x.view(np.ndarray) if isinstance(x, HostField) else x for x in inputs]
        if out is not None:
            for o in out:
                if isinstance(o, HostField):
                    o._touch()
            kwargs['out'] = tuple(o.view(np.ndarray) if isinstance(o, HostField) else o for o in out)
        res = getattr(ufunc, method)(*plain, **kwargs)
        if out is not None:
            return out[0] if len(out) == 1 else out
        return res

    def fill(self, value):
        np.ndarray.fill(self, value)
        self._touch()

    def device_ptr(self):
        """device address of the twin, uploaded first if the host side was written"""
        from runtime import rt
        r = rt()
        root = self if self._root is None else self._root
        if root._stale or root._dev is None:
            root._dev = r.to_device(root.view(np.ndarray), dtype=np.float64)
            root._stale = False
        return r.ptr(root._dev)


class DeviceState(object):
    """[nvar, ny, nx] float64 in HBM (torch owns the allocation) + pinned host mirror"""

    def __init__(self, nvar, ny, nx, device=None):
        self.nvar, self.ny, self.nx = nvar, ny, nx
        self.device = device if device is not None else torch.device('cuda', torch.cuda.current_device())
        if self.device.type == 'cuda':
            from runtime import rt
            self.dev = rt().alloc((nvar, ny, nx))     # symmetric heap when the domain is decomposed
        else:
            self.dev = torch.zeros((nvar, ny, nx), dtype=torch.float64, device=self.device)
        self._host_t = None        # pinned torch tensor, allocated on first host access
        self._host = None          # numpy view of it
        self.host_fresh = [True]*nvar
        self.dev_fresh = [True]*nvar
        self.fieldbytes = ny*nx*8
        self.shape = (nvar, ny, nx)
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self._upload_evt = None    # CUDA event behind the last asynchronous upload from the mirror
        # Uploads run on a copy stream of their own: the first stale field a kernel asks for starts
        # the upload of EVERY stale field (the one asked for first), one event per field, and the
        # compute stream waits for a field's event only when a kernel is about to touch that
        # field -- so the fields a step needs later cross PCIe while its first kernels run.
        # The order of a batch is the order in which the kernels asked for the fields after the
        # previous batch (a time step asks for them in the same order every step).
        self._copy_stream = None
        self._pending = [None]*nvar    # per field: event of an upload still (possibly) in flight
        self._asked = []               # fields in the order they were asked for since the last batch
        self._asked_before = []        # ... and between the two batches before that

    def _wait_upload(self):
        """an upload from the pinned mirror is asynchronous: the host must not write the mirror
        again while the DMA may still be reading it"""
        if self._upload_evt is not None:
            self._upload_evt.synchronize()
            self._upload_evt = None

    # -- host side -----------------------------------------------------------
    def _ensure_host(self):
        if self._host is None:
            self._host_t = torch.zeros((self.nvar, self.ny, self.nx), dtype=torch.float64)
            if self.device.type == 'cuda':
                self._host_t = self._host_t.pin_memory()
            self._host = self._host_t.numpy()
            # the mirror starts as zeros; anything already on the device is newer
            self.host_fresh = [False]*self.nvar

    def _fields(self, k):
        return range(self.nvar) if k is None else (k,)

    def _refresh_host(self, k=None):
        self._ensure_host()
        self._wait_upload()        # every host write passes through here first (TrackedArray._before_read)
        todo = [f for f in self._fields(k) if not self.host_fresh[f]]
        if not todo:
            return
        for f in todo:
            if not self.dev_fresh[f]:
                raise RuntimeError('DeviceState: field %d stale on both sides' % f)
            self._await_field(f)
            self._host_t[f].copy_(self.dev[f], non_blocking=True)
            self.host_fresh[f] = True
            self.d2h_bytes += self.fieldbytes
        if self.device.type == 'cuda':
            torch.cuda.current_stream().synchronize()

    def _host_written(self, k=None):
        for f in self._fields(k):
            self.host_fresh[f] = True
            self.dev_fresh[f] = False

    def host_view(self, k=None):
        """TrackedArray on field k (or on the whole state), fresh from the device"""
        self._refresh_host(k)
        base = self._host if k is None else self._host[k]
        v = base.view(TrackedArray)
        v._own = (self, k)
        return v

    def __getitem__(self, k):
        if isinstance(k, (int, np.integer)):
            return self.host_view(int(k))
        return self.host_view(None)[k]

    def __setitem__(self, k, value):
        if isinstance(k, (int, np.integer)):
            own = getattr(value, '_own', None)
            if isinstance(value, TrackedArray) and own is not None and own[0] is self and own[1] == int(k) \
                    and value.shape == (self.ny, self.nx) and value.flags['C_CONTIGUOUS'] \
                    and self._host is not None and value.ctypes.data == self._host[int(k)].ctypes.data:
                # `state[k] += a` ends with `state[k] = <the view it just updated in place>`: the
                # in-place operation (host or device) has already recorded who holds the fresh copy
                return
            v = self.host_view(int(k))
            if not (isinstance(value, np.ndarray) and np.shares_memory(value, v)):
                np.ndarray.__setitem__(v, slice(None), np.asarray(value))
            self._host_written(int(k))
        else:
            v = self.host_view(None)
            v[k] = value

    # -- device side ---------------------------------------------------------
    def _await_field(self, f):
        """the compute stream waits for the upload of field f, if one may still be in flight"""
        e = self._pending[f]
        if e is not None:
            torch.cuda.current_stream().wait_event(e)
            self._pending[f] = None

    def to_device(self, k=None):
        want = list(self._fields(k))
        if k is not None and k not in self._asked:
            self._asked.append(k)
        if self.device.type != 'cuda':
            for f in want:
                if not self.dev_fresh[f]:
                    self.dev[f].copy_(self._host_t[f])
                    self.dev_fresh[f] = True
                    self.h2d_bytes += self.fieldbytes
            return
        if any(not self.dev_fresh[f] for f in want):
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=self.device)
            cs = self._copy_stream
            # kernels already enqueued may still read what the copies overwrite
            cs.wait_stream(torch.cuda.current_stream())
            if len(self._asked) > 1 or not self._asked_before:
                self._asked_before = list(self._asked)
            self._asked = [k] if k is not None else []
            stale = [f for f in want if not self.dev_fresh[f]]
            stale += [f for f in self._asked_before if not self.dev_fresh[f] and f not in stale]
            stale += [f for f in range(self.nvar) if not self.dev_fresh[f] and f not in stale]
            with torch.cuda.stream(cs):
                for f in stale:
                    self.dev[f].copy_(self._host_t[f], non_blocking=True)
                    self.dev_fresh[f] = True
                    self.h2d_bytes += self.fieldbytes
                    e = torch.cuda.Event()
                    e.record(cs)
                    self._pending[f] = e
                self._upload_evt = self._pending[stale[-1]]
        for f in want:
            self._await_field(f)

    def rptr(self, k):
        """device address of field k for reading"""
        self.to_device(k)
        return ctypes.c_void_p(self.dev[k].data_ptr())

    def wptr(self, k):
        """device address of field k for (partial) writing: mirror becomes stale"""
        self.to_device(k)
        self.host_fresh[k] = False
        return ctypes.c_void_p(self.dev[k].data_ptr())

    def all_ptr(self, write=False):
        """device address of the whole [nvar,ny,nx] block"""
        self.to_device(None)
        if write:
            self.host_fresh = [False]*self.nvar
        return ctypes.c_void_p(self.dev.data_ptr())

    @property
    def size(self):
        return self.nvar*self.ny*self.nx

    def zero_(self):
        for f in range(self.nvar):
            self._await_field(f)
        self.dev.zero_()
        self.dev_fresh = [True]*self.nvar
        self.host_fresh = [False]*self.nvar if self._host is not None else [True]*self.nvar
        if self._host is None:
            self.host_fresh = [True]*self.nvar   # mirror not yet materialised (zeros anyway)

    def upload_all_from(self, array):
        """replace the whole state by a host array (restart, tests)"""
        self._ensure_host()
        self._host[...] = array
        self.host_fresh = [True]*self.nvar
        self.dev_fresh = [False]*self.nvar
        self.to_device(None)

    def numpy(self):
        """fresh plain-numpy copy of the whole state"""
        self._refresh_host(None)
        return np.array(self._host, copy=True)
