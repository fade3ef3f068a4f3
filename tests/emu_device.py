"""CPU emulation of the C ABI for tests of the HOST layer's numerics.  TEST INFRASTRUCTURE.

`tests/mock_device.py` records which entry points the Python host layer (fluid2d_b200/core)
calls; this module goes one step further and *executes* every call on CPU memory with the
oracle's kernels (oracle/kernels.py, oracle/model.py:MG), one entry point at a time, with the
argument meaning include/f2d_b200.h documents.  Swapping it in for the runtime singleton lets
the `-m "not gpu"` suite run the product's host layer -- models, operators, time schemes,
fluxes driver, diagnostics, dt control -- end to end and compare the fields with the fixtures
frozen from the reference's own Python (tests/golden/*.npz): what is verified is that the host
layer asks for the right operations on the right buffers in the right order, with arithmetic
attached.  It says nothing about the CUDA kernels (the -m gpu tests compare those with the
same oracle through the real library), and nothing under fluid2d_b200/ can reach this module:
the product has no CPU path.
"""
import ctypes
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from oracle import kernels as K          # noqa: E402
from oracle import model as om           # noqa: E402

fa, ff, fo, fd, fm = (K.fortran_advection, K.fortran_fluxes, K.fortran_operators, K.fortran_diag,
                      K.fortran_multigrid)


def _addr(p):
    if p is None:
        return 0
    if isinstance(p, ctypes.c_void_p):
        return p.value or 0
    return int(p)


def f64(p, *shape):
    """numpy view of the doubles at address p (None for a NULL pointer)"""
    a = _addr(p)
    if not a:
        return None
    n = int(np.prod(shape))
    return np.ctypeslib.as_array((ctypes.c_double*n).from_address(a)).reshape(shape)


def i8(p, *shape):
    a = _addr(p)
    if not a:
        return None
    n = int(np.prod(shape))
    return np.ctypeslib.as_array((ctypes.c_int8*n).from_address(a)).reshape(shape)


def _ones_msk(ny, nx):
    return np.ones((ny, nx), dtype=np.int8)


def _corner_of_all_fluid(ny, nx):
    """corner mask of an all-fluid domain: 1 except on the last row and column"""
    m = np.ones((ny, nx), dtype=np.int8)
    m[-1, :] = 0
    m[:, -1] = 0
    return m


class _Handle(object):
    """what an f2d_mg_t* stands for here: the oracle's hierarchy"""

    def __init__(self, mg):
        self.mg = mg
        self.Aplanes = {}


class EmuLib(object):
    """one method per entry point of include/f2d_b200.h that the host layer reaches
    (names without the f2d_ prefix, as fluid2d_b200/_lib.py exposes them)"""

    def __init__(self):
        self.handles = {}
        self.count = 0
        self.called = set()

    def __getattribute__(self, name):
        v = object.__getattribute__(self, name)
        if callable(v) and not name.startswith("_"):
            object.__getattribute__(self, "called").add(name)
        return v

    # ---- bookkeeping
    def abi_version(self):
        return 1

    def last_error(self):
        return b""

    def launch_count(self):
        return self.count

    def launch_count_reset(self):
        self.count = 0

    def reduce_scratch_len(self):
        return 64

    # ---- copies, halo
    def copy(self, dst, src, nbytes, stream):
        ctypes.memmove(_addr(dst), _addr(src), int(nbytes))
        return 0

    def zero(self, dst, nbytes, stream):
        ctypes.memset(_addr(dst), 0, int(nbytes))
        return 0

    def fill_halo(self, x, nh, ny, nx, stream):
        fm.fillhalo(f64(x, ny, nx), nh)
        return 0

    def fill_halo_i8(self, x, nh, ny, nx, stream):
        a = i8(x, ny, nx)
        t = a.astype(np.float64)
        fm.fillhalo(t, nh)
        a[...] = t.astype(np.int8)
        return 0

    def _fill(self, a, nh, mode):
        if mode == 1:
            fm.fillhalo(a, nh)
        elif mode == 2:           # y-slab decomposition: x images only, rows come from the exchange
            a[:, :nh] = a[:, -2*nh:-nh]
            a[:, -nh:] = a[:, nh:2*nh]

    # ---- advection
    def _adv(self, upwind, msk, q, dq, u, v, xflx, yflx, cst5, nh, method, order, ny, nx, fill):
        if nh != 3:
            return 1
        m = i8(msk, ny, nx)
        if m is None:
            m = _ones_msk(ny, nx)
        cst = np.array([cst5[k] for k in range(5)])
        Q, DQ, U, V = f64(q, ny, nx), f64(dq, ny, nx), f64(u, ny, nx), f64(v, ny, nx)
        XF, YF = f64(xflx, ny, nx), f64(yflx, ny, nx)
        if XF is None:
            (fa.adv_upwind if upwind else fa.adv_centered)(m, Q, DQ, U, V, cst, nh, method, order)
        else:
            (ff.adv_upwind if upwind else ff.adv_centered)(m, Q, DQ, U, V, XF, YF, cst, nh, method, order)
        self._fill(DQ, nh, fill)
        return 0

    def adv_upwind(self, *a):
        return self._adv(True, *a[:-1])

    def adv_multi(self, msk, q, dq, ntr, u, v, cst5, nh, upwind, method, order, xbase, xout, coef, ny, nx, fill,
                  stream):
        for k in range(ntr):
            rc = self._adv(bool(upwind), msk, q[k], dq[k], u, v, None, None, cst5, nh, method, order, ny, nx, fill)
            if rc:
                return rc
            if xout is not None:
                f64(xout[k], ny, nx)[...] = f64(xbase[k], ny, nx)+coef*f64(dq[k], ny, nx)
        return 0

    def adv_centered(self, *a):
        return self._adv(False, *a[:-1])

    # ---- operators
    def celltocorner(self, xr, xp, ny, nx, stream):
        fo.celltocorner(f64(xr, ny, nx), f64(xp, ny, nx))
        return 0

    def cornertocell(self, xp, xr, ny, nx, stream):
        fo.cornertocell(f64(xp, ny, nx), f64(xr, ny, nx))
        return 0

    def orthogradient(self, msk, psi, dx, dy, nh, u, v, ny, nx, stream):
        fo.computeorthogradient(i8(msk, ny, nx), f64(psi, ny, nx), dx, dy, nh, f64(u, ny, nx), f64(v, ny, nx))
        return 0

    def mask_orthogradient(self, msk, mskp, psi, dx, dy, nh, u, v, ny, nx, stream):
        m, mp = i8(msk, ny, nx), i8(mskp, ny, nx)
        if m is None:
            m, mp = _ones_msk(ny, nx), _corner_of_all_fluid(ny, nx)
        P = f64(psi, ny, nx)
        P *= mp
        fo.computeorthogradient(m, P, dx, dy, nh, f64(u, ny, nx), f64(v, ny, nx))
        return 0

    def add_diffusion(self, msk, trac, dx, nh, Kdiff, dtrac, ny, nx, fill, stream):
        D = f64(dtrac, ny, nx)
        fo.add_diffusion(i8(msk, ny, nx), f64(trac, ny, nx), dx, nh, Kdiff, D)
        self._fill(D, nh, fill)
        return 0

    def add_torque(self, msk, buoy, dx, nh, gravity, domega, ny, nx, premask, fill, stream):
        m, D = i8(msk, ny, nx), f64(domega, ny, nx)
        if premask:
            D *= m
        fo.add_torque(m, f64(buoy, ny, nx), dx, nh, gravity, D)
        self._fill(D, nh, fill)
        return 0

    def noslip_source(self, msknoslip, psi, y, dx, dy, nh, ny, nx, stream):
        fo.computenoslipsourceterm(i8(msknoslip, ny, nx), f64(psi, ny, nx), f64(y, ny, nx), dx, dy, nh)
        return 0

    # ---- reductions (sequential sums: the Fortran's own order)
    def computedotprod(self, msk, x, y, nh, ny, nx, out, scratch, stream):
        f64(out, 1)[0] = fd.computedotprod(i8(msk, ny, nx), f64(x, ny, nx), f64(y, ny, nx), nh)
        return 0

    def computemax(self, msk, x, nh, ny, nx, out, scratch, stream):
        f64(out, 1)[0] = fd.computemax(i8(msk, ny, nx), f64(x, ny, nx), nh)
        return 0

    def computesum(self, msk, x, nh, ny, nx, out, scratch, stream):
        f64(out, 1)[0] = fd.computesum(i8(msk, ny, nx), f64(x, ny, nx), nh)
        return 0

    def computesumandnorm(self, msk, x, nh, ny, nx, out, scratch, stream):
        f64(out, 2)[:] = fd.computesumandnorm(i8(msk, ny, nx), f64(x, ny, nx), nh)
        return 0

    def computekemaxu(self, msk, u, v, nh, ny, nx, out, scratch, stream):
        f64(out, 2)[:] = fd.computekemaxu(i8(msk, ny, nx), f64(u, ny, nx), f64(v, ny, nx), nh)
        return 0

    def computenorm(self, msk, x, nh, ny, nx, out, scratch, stream):
        f64(out, 1)[0] = fm.computenorm(i8(msk, ny, nx), f64(x, ny, nx), nh)
        return 0

    def domain_sum(self, x, nh, ny, nx, out, scratch, stream):
        f64(out, 1)[0] = np.sum(f64(x, ny, nx)[nh:-nh, nh:-nh])
        return 0

    def diag_euler(self, msk, u, v, w, psi, source, xr, yr, nh, ny, nx, out, scratch, stream):
        m = i8(msk, ny, nx)
        W = f64(w, ny, nx)
        o = f64(out, 8)
        ke, maxu = fd.computekemaxu(m, f64(u, ny, nx), f64(v, ny, nx), nh)
        z, z2 = fd.computesumandnorm(m, W, nh)
        o[0], o[1], o[2], o[3] = maxu, ke, z, z2
        o[4] = fd.computedotprod(m, W, f64(xr, ny, nx), nh)
        o[5] = fd.computedotprod(m, W, f64(yr, ny, nx), nh)
        o[6] = fd.computesum(m, f64(psi, ny, nx), nh)
        o[7] = fd.computedotprod(m, W, f64(source, ny, nx), nh)
        return 0

    # ---- time-scheme combinations: the numpy expressions of timescheme.py
    def ts_axpy(self, y, c, a, n, stream):
        Y = f64(y, n)
        Y += c*f64(a, n)
        return 0

    def ts_xpay(self, out, x, c, a, n, stream):
        f64(out, n)[:] = f64(x, n) + c*f64(a, n)
        return 0

    def ts_xpay2(self, out, x, c, a, b, n, stream):
        f64(out, n)[:] = f64(x, n) + c*(f64(a, n)+f64(b, n))
        return 0

    def ts_rk3ssp_final(self, x, c, a, b, d, n, stream):
        X = f64(x, n)
        X += c*(f64(a, n)+f64(b, n)+4*f64(d, n))
        return 0

    def ts_ab2(self, x, c0, a, c1, b, n, stream):
        X = f64(x, n)
        X += c0*f64(a, n) - c1*f64(b, n)
        return 0

    def ts_ab3(self, x, c0, a, c1, b, c2, d, n, stream):
        X = f64(x, n)
        X += c0*f64(a, n) - c1*f64(b, n) + c2*f64(d, n)
        return 0

    def ts_set_xpay(self, x, xb, c, a, n, stream):
        f64(x, n)[:] = f64(xb, n) + c*f64(a, n)
        return 0

    def ts_asselin(self, xs, c, x, xb, n, stream):
        XS = f64(xs, n)
        XS += c*(f64(x, n)+f64(xb, n)-2*XS)
        return 0

    def ts_am3(self, x, xs, xb, n, stream):
        X = f64(x, n)
        X[:] = (1./12.)*(5.*X + 8.*f64(xs, n)-f64(xb, n))
        return 0

    # ---- elementwise glue
    def mul_field(self, y, a, n, stream):
        Y = f64(y, n)
        Y *= f64(a, n)
        return 0

    def mul_mask(self, y, a, n, stream):
        Y = f64(y, n)
        Y *= i8(a, n)
        return 0

    def scale(self, y, alpha, n, stream):
        Y = f64(y, n)
        Y *= alpha
        return 0

    def add_scaled(self, y, alpha, a, n, stream):
        Y = f64(y, n)
        Y += alpha*f64(a, n)
        return 0

    def add_scaled_mask(self, y, alpha, a, n, stream):
        Y = f64(y, n)
        Y += alpha*i8(a, n)
        return 0

    def set_sum(self, y, a, alpha, b, n, stream):
        f64(y, n)[:] = f64(a, n) + alpha*f64(b, n)
        return 0

    def div_scalar(self, y, d, n, stream):
        Y = f64(y, n)
        Y /= d
        return 0

    def sub_lin2_mask(self, y, pa, a, pb, b, mask, n, stream):
        Y = f64(y, n)
        Y -= (pa*f64(a, n)+pb*f64(b, n))*i8(mask, n)
        return 0

    def sub_lin2(self, y, pa, a, pb, b, n, stream):
        Y = f64(y, n)
        Y -= (f64(a, n)*pa+f64(b, n)*pb)
        return 0

    def sub_devscalar(self, y, dev_scalar, denom, n, stream):
        Y = f64(y, n)
        Y -= f64(dev_scalar, 1)[0]/denom
        return 0

    def sub_devscalar_mask(self, y, dev_scalar, denom, a, n, stream):
        Y = f64(y, n)
        Y -= (f64(dev_scalar, 1)[0]/denom)*i8(a, n)
        return 0

    # ---- thermal-wind model: the numpy expressions of operators.py:330-394
    def extrapolate_bry(self, x, nh, ny, nx, axis, stream):
        X = f64(x, ny, nx)
        if axis == 0:
            X[:, -nh] = 2*X[:, -nh-1]-X[:, -nh-2]
            X[:, nh-1] = 2*X[:, nh]-X[:, nh+1]
        else:
            X[-nh, :] = 2*X[-nh-1, :]-X[-nh-2, :]
            X[nh-1, :] = 2*X[nh, :]-X[nh+1, :]
        return 0

    @staticmethod
    def _diffx(a, dx):
        return 0.5*(a[1:-1, 2:]-a[1:-1, :-2])/dx

    @staticmethod
    def _diffz(a, dy):
        return 0.5*(a[2:, 1:-1]-a[:-2, 1:-1])/dy

    def tw_torque(self, msk, b, V, dx, dy, gravity, f0, y, ny, nx, stream):
        Y = f64(y, ny, nx)
        Y[1:-1, 1:-1] = self._diffx(f64(b, ny, nx), dx)*gravity
        Y[1:-1, 1:-1] -= self._diffz(f64(V, ny, nx), dy)*f0
        Y *= i8(msk, ny, nx)
        return 0

    def tw_coriolis(self, msk, u, f0, y, ny, nx, stream):
        Y, U = f64(y, ny, nx), f64(u, ny, nx)
        Y[:, 1:] = - 0.5*f0*(U[:, :-1]+U[:, 1:])
        Y *= i8(msk, ny, nx)
        return 0

    def jacobian(self, msk, x, y, dx, dy, out, ny, nx, stream):
        X, Y, O = f64(x, ny, nx), f64(y, ny, nx), f64(out, ny, nx)
        O[:, :] = 0.
        O[1:-1, 1:-1] = self._diffx(X, dx)*self._diffz(Y, dy)-self._diffz(X, dy)*self._diffx(Y, dx)
        O *= i8(msk, ny, nx)
        return 0

    def negative_part(self, out, x, n, stream):
        X = f64(x, n)
        f64(out, n)[:] = np.where(X > 0, 0., X)
        return 0

    # ---- fluxes driver, output
    def flx_cellvel(self, u, v, uc, vc, nh, ny, nx, fill, stream):
        U, V, UC, VC = f64(u, ny, nx), f64(v, ny, nx), f64(uc, ny, nx), f64(vc, ny, nx)
        UC[nh:-nh, nh:-nh] = (0.5*(U+np.roll(U, 1, axis=1)))[nh:-nh, nh:-nh]
        VC[nh:-nh, nh:-nh] = (0.5*(V+np.roll(V, 1, axis=0)))[nh:-nh, nh:-nh]
        self._fill(UC, nh, fill)
        self._fill(VC, nh, fill)
        return 0

    def flx_split(self, rev, irr, fwd, bwd, cff, sign, n, stream):
        F, B = f64(fwd, n), f64(bwd, n)
        sb = sign*B
        r = cff*(F+sb)
        i = cff*(F-sb)
        f64(rev, n)[:] = r
        f64(irr, n)[:] = i
        return 0

    def pack_interior_f32(self, x, out, nh, ny, nx, stream):
        m, n = ny-2*nh, nx-2*nh
        o = np.ctypeslib.as_array((ctypes.c_float*(m*n)).from_address(_addr(out))).reshape(m, n)
        o[...] = f64(x, ny, nx)[nh:-nh, nh:-nh].astype(np.float32)
        return 0

    # ---- multigrid
    def mg_create(self, h, cornermask, ny, nx, dx, dy, omega, hydroepsilon, Rd, stream):
        cm = f64(cornermask, ny, nx).copy()
        mg = om.MG(cm, nx-6, ny-6, dx, dy, omega=omega, hydroepsilon=hydroepsilon, Rd=(Rd if Rd > 0 else None))
        key = len(self.handles)+1
        self.handles[key] = _Handle(mg)
        h._obj.value = key
        return 0

    def mg_set_relaxation(self, h, mode):
        self._h(h).mg.relaxation = 'tridiagonal' if mode else 'default'
        return 0

    def mg_create_slab(self, h, comm, cornermask, ny_loc, nx, dx, dy, omega, hydroepsilon, Rd, stream):
        glob = self._assemble(f64(cornermask, ny_loc, nx))
        mg = om.MG(glob, nx-6, glob.shape[0]-6, dx, dy, omega=omega, hydroepsilon=hydroepsilon,
                   Rd=(Rd if Rd > 0 else None))
        key = len(self.handles)+1
        self.handles[key] = _Handle(mg)
        self.handles[key].slab = (ny_loc, nx)
        h._obj.value = key
        return 0

    def _h(self, h):
        return self.handles[_addr(h)]

    def mg_destroy(self, h):
        self.handles.pop(_addr(h), None)
        return 0

    def mg_slab_levels(self, h):
        return 1 if getattr(self._h(h), "slab", None) else 0

    def mg_nlevels(self, h):
        return self._h(h).mg.nlevs

    def mg_level_shape(self, h, lev, ny, nx):
        m, n = self._h(h).mg.sizes[lev]
        ny._obj.value, nx._obj.value = m+6, n+6
        return 0

    def mg_level_matrix_mode(self, h, lev):
        return 0

    def mg_level_ptr(self, h, lev, which):
        H = self._h(h)
        mg = H.mg
        if which == 0:
            return mg.msk[lev].ctypes.data
        if which == 1:     # 5 planes [5][ny][nx]
            if lev not in H.Aplanes:
                H.Aplanes[lev] = np.ascontiguousarray(np.moveaxis(mg.A[lev], 2, 0))
            return H.Aplanes[lev].ctypes.data
        return (mg.x, mg.b, mg.r)[which-2][lev].ctypes.data

    def mg_set_graphs(self, h, enable):
        return 0

    def mg_vcycle(self, h, lev1, stream):
        self._h(h).mg.vcycle(lev1)
        return 0

    def mg_fcycle(self, h, lev1, stream):
        self._h(h).mg.fcycle(lev1)
        return 0

    def _shape(self, h):
        H = self._h(h)
        return getattr(H, "slab", None) or H.mg.x[0].shape

    def mg_two_vcycle(self, h, psi, rhs, stream):
        H = self._h(h)
        mg = H.mg
        shape = self._shape(h)
        if getattr(H, "slab", None):
            P, R = self._assemble(f64(psi, *shape)), self._assemble(f64(rhs, *shape))
            mg.two_vcycle(P, R)
            f64(psi, *shape)[...] = self._local_rows(P, shape[0])
            return 0
        mg.two_vcycle(f64(psi, *shape), f64(rhs, *shape))
        return 0

    def mg_solve(self, h, psi, rhs, tol, maxite, nite, res, stream):
        H = self._h(h)
        mg = H.mg
        shape = self._shape(h)
        try:
            if getattr(H, "slab", None):
                P, R = self._assemble(f64(psi, *shape)), self._assemble(f64(rhs, *shape))
                n, r = mg.solve(P, R, maxite=maxite, tol=tol)
                f64(psi, *shape)[...] = self._local_rows(P, shape[0])
            else:
                n, r = mg.solve(f64(psi, *shape), f64(rhs, *shape), maxite=maxite, tol=tol)
        except RuntimeError:
            return 4
        if nite is not None:
            nite._obj.value = int(n)
        if res is not None:
            res._obj.value = float(r)
        return 0

    def invert_vorticity(self, h, msk, mskp, w, psi, u, v, work, rhsp, psi_island, full, perio, area, dx, dy,
                         nh, nite, res, scratch, stream):
        ny, nx = self._shape(h)
        n = ny*nx
        self.celltocorner(w, work, ny, nx, stream)
        if _addr(rhsp):
            self.add_scaled(work, -1., rhsp, n, stream)
        if full:
            err = self.mg_solve(h, psi, work, 1e-11, 4, nite, res, stream)
            if err:
                return err
            if perio:
                P = f64(psi, ny, nx)
                total = np.array([np.sum(P[nh:-nh, nh:-nh])])
                if self.nranks > 1:
                    self.comm_allreduce(None, total.ctypes.data, 1, 0, stream)
                P -= total[0]/area
        else:
            self.mg_two_vcycle(h, psi, work, stream)
            if nite is not None:
                nite._obj.value = 1
            if res is not None:
                res._obj.value = 0.
        if not _addr(psi_island):
            rc = self.mask_orthogradient(msk, mskp, psi, dx, dy, nh, u, v, ny, nx, stream)
        else:
            self.mul_mask(psi, mskp, n, stream)
            self.add_scaled(psi, 1., psi_island, n, stream)
            rc = self.orthogradient(msk, psi, dx, dy, nh, u, v, ny, nx, stream)
        stage = getattr(self, '_uv_stage', {}).pop(_addr(h), None)
        if stage is not None and not rc:
            # f2d_mg_set_uv_stage: out = base + c*([extra +] tendency), numpy's rounding sequence
            ub, vb, ue, ve, uo, vo, c = stage
            for d, b, e, o in ((u, ub, ue, uo), (v, vb, ve, vo)):
                D = f64(d, ny, nx)
                f64(o, ny, nx)[...] = f64(b, ny, nx)+c*((f64(e, ny, nx)+D) if _addr(e) else D)
        return rc

    def mg_set_uv_stage(self, h, ub, vb, ue, ve, uo, vo, c):
        if not hasattr(self, '_uv_stage'):
            self._uv_stage = {}
        self._uv_stage[_addr(h)] = (ub, vb, ue, ve, uo, vo, c)
        return 0

    # ---- y-slab decomposition (npx = 1, npy = ranks): the communicator entry points over a
    # gloo process group (one CPU process per rank); the slab multigrid is emulated by
    # assembling the global problem on every rank, running the oracle's hierarchy on it and
    # keeping the local rows -- Jacobi smoothing is order independent, so this is what the
    # distributed cycles compute
    nranks = 1
    rank = 0

    def _gather(self, a):
        import torch.distributed as dist
        mine = torch.from_numpy(np.ascontiguousarray(a).copy())
        parts = [torch.zeros_like(mine) for _ in range(self.nranks)]
        dist.all_gather(parts, mine)
        return [p.numpy() for p in parts]

    def _assemble(self, loc, nh=3):
        """local slabs [ny_loc, nx] of every rank -> the global array (south to north): the
        interiors, with the outer halo rows of the first and last rank"""
        parts = self._gather(loc)
        return np.concatenate([parts[0][:nh]]+[p[nh:-nh] for p in parts]+[parts[-1][-nh:]], axis=0)

    def _local_rows(self, glob, ny_loc, nh=3):
        j0 = self.rank*(ny_loc-2*nh)
        return glob[j0:j0+ny_loc]

    def comm_barrier(self, comm, stream):
        if self.nranks > 1:
            import torch.distributed as dist
            dist.barrier()
        return 0

    def comm_rank(self, comm):
        return self.rank

    def comm_size(self, comm):
        return self.nranks

    def fill_halo_x(self, x, nh, ny, nx, stream):
        self._fill(f64(x, ny, nx), nh, 2)
        return 0

    def comm_exchange_y(self, comm, x, nh, ny, nx, stream):
        X = f64(x, ny, nx)
        if self.nranks == 1:
            X[:nh] = X[-2*nh:-nh]
            X[-nh:] = X[nh:2*nh]
            return 0
        parts = self._gather(np.stack([X[nh:2*nh], X[-2*nh:-nh]]))     # my bottom / top interior rows
        X[:nh] = parts[(self.rank-1) % self.nranks][1]
        X[-nh:] = parts[(self.rank+1) % self.nranks][0]
        return 0

    def comm_allreduce(self, comm, vals, n, maxmask, stream):
        V = f64(vals, n)
        parts = self._gather(V)
        for k in range(n):
            col = [p[k] for p in parts]
            if (maxmask >> k) & 1:
                V[k] = max(col)
            else:
                acc = col[0]
                for c in col[1:]:
                    acc = acc+c
                V[k] = acc
        return 0


class EmuRuntime(object):
    """stands where runtime.Runtime stands; state lives in CPU tensors"""

    def __init__(self):
        self.lib = EmuLib()
        self.device = torch.device("cpu")
        self.scratch = torch.zeros(64, dtype=torch.float64)
        self.out = torch.zeros(16, dtype=torch.float64)
        self.out_host = torch.zeros(16, dtype=torch.float64)
        self.comm = None
        self.nranks = 1
        self.rank = 0

    def ensure_comm(self, nranks, fieldbytes):
        """y-slabs: one CPU process per rank under torchrun, gloo process group"""
        if nranks == 1 or self.comm is not None:
            return
        import runtime
        import torch.distributed as dist
        runtime.ensure_dist()
        if dist.get_world_size() != nranks:
            raise RuntimeError("param.npy = %d but %d processes were launched" % (nranks, dist.get_world_size()))
        self.rank = self.lib.rank = dist.get_rank()
        self.nranks = self.lib.nranks = nranks
        self.comm = ctypes.c_void_p(1)      # a non-null token: the host layer only tests it and passes it on

    def exchange_y(self, ptr, ny, nx, nh=3):
        self.lib.comm_exchange_y(self.comm, ptr, nh, ny, nx, self.stream)

    def alloc(self, shape, dtype=torch.float64):
        return torch.zeros(shape, dtype=dtype)

    @property
    def stream(self):
        return None

    def ptr(self, t):
        return ctypes.c_void_p(t.data_ptr()) if t is not None else None

    def to_device(self, a, dtype=None):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype).copy())

    def read_out(self, n):
        return self.out[:n].tolist()


def install():
    """activate the host layer on top of the emulated library; returns (api, runtime)"""
    import fluid2d_b200
    api = fluid2d_b200.api()
    import runtime
    import devarray
    import mock_device
    emu = EmuRuntime()
    runtime._rt = emu
    del mock_device._REGISTRY[:]
    mock_device.patch_devicestate(devarray)
    return api, emu


def uninstall():
    import mock_device
    mock_device.uninstall()
