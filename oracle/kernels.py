"""ctypes front-end of the CPU ORACLE (oracle/f2d_oracle.c) -- test infrastructure only.

It exposes the same call surface as the reference's five f2py modules
(`fortran_advection`, `fortran_fluxes`, `fortran_operators`, `fortran_diag`,
`gmg.fortran_multigrid`; f2py lower-cases names and hides the trailing m,n
arguments -- SURVEY.md section 2.2) so that

  * the reference's own Python (operators.py, gmg/level.py ...) can be imported in
    the build container on top of it (oracle/shim, tests/golden/make_golden.py), and
  * the oracle's restated orchestration (oracle/model.py) calls the very same
    functions.

Arrays are numpy, C-ordered, [ny, nx]; float64 fields, int8 masks.  f2py's
`intent(inplace)` semantics (arrays mutated in place, logical indexing preserved)
are what these wrappers provide.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, "_build", "libf2d_oracle.so")
# tests/test_contraction_headroom.py points this at a variant of the same source compiled WITH
# FMA contraction, to see how far contraction alone moves a run (never set otherwise)
_ALT = os.environ.get("F2D_ORACLE_LIB")

c_dp = ctypes.POINTER(ctypes.c_double)
c_bp = ctypes.POINTER(ctypes.c_int8)
c_int = ctypes.c_int
c_dbl = ctypes.c_double


def build(force=False):
    """Compile oracle/f2d_oracle.c with gcc (see oracle/Makefile for the flags)."""
    src = os.path.join(_HERE, "f2d_oracle.c")
    if _ALT:
        return _ALT
    if (not force and os.path.exists(_LIBPATH)
            and os.path.getmtime(_LIBPATH) >= os.path.getmtime(src)):
        return _LIBPATH
    subprocess.check_call(["make", "-s", "-C", _HERE, "_build/libf2d_oracle.so"])
    return _LIBPATH


_lib = None


class _Flushing(object):
    """attribute proxy: call the C function, then copy converted in/out operands back"""

    def __init__(self, cdll):
        self._cdll = cdll
        self._cache = {}

    def __getattr__(self, name):
        f = self._cache.get(name)
        if f is None:
            cf = getattr(self._cdll, name)

            def f(*args, _cf=cf):
                r = _cf(*args)
                if _writebacks:
                    _flush()
                return r
            self._cache[name] = f
        return f


def lib():
    global _lib
    if _lib is None:
        cdll = ctypes.CDLL(build())
        _declare(cdll)
        _lib = _Flushing(cdll)
    return _lib


def _declare(L):
    L.f2d_oracle_set_reduce_mode.argtypes = [c_int]
    L.f2d_oracle_get_reduce_mode.restype = c_int
    L.f2d_oracle_num_threads.restype = c_int
    L.f2d_oracle_set_num_threads.argtypes = [c_int]
    for name in ("f2d_oracle_adv_upwind", "f2d_oracle_adv_centered"):
        f = getattr(L, name)
        f.restype = c_int
        f.argtypes = [c_bp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp,
                      c_int, c_int, c_int, c_int, c_int]
    L.f2d_oracle_computeorthogradient.argtypes = [c_bp, c_dp, c_dbl, c_dbl, c_int,
                                                  c_dp, c_dp, c_int, c_int]
    L.f2d_oracle_celltocorner.argtypes = [c_dp, c_dp, c_int, c_int]
    L.f2d_oracle_cornertocell.argtypes = [c_dp, c_dp, c_int, c_int]
    L.f2d_oracle_add_diffusion.argtypes = [c_bp, c_dp, c_dbl, c_int, c_dbl, c_dp,
                                           c_int, c_int]
    L.f2d_oracle_computenoslipsourceterm.restype = c_dbl
    L.f2d_oracle_computenoslipsourceterm.argtypes = [c_bp, c_dp, c_dp, c_dbl, c_dbl,
                                                     c_int, c_int, c_int]
    L.f2d_oracle_add_torque.argtypes = [c_bp, c_dp, c_dbl, c_int, c_dbl, c_dp,
                                        c_int, c_int]
    L.f2d_oracle_computedotprod.restype = c_dbl
    L.f2d_oracle_computedotprod.argtypes = [c_bp, c_dp, c_dp, c_int, c_int, c_int]
    L.f2d_oracle_computemax.restype = c_dbl
    L.f2d_oracle_computemax.argtypes = [c_bp, c_dp, c_int, c_int, c_int]
    L.f2d_oracle_computesum.restype = c_dbl
    L.f2d_oracle_computesum.argtypes = [c_bp, c_dp, c_int, c_int, c_int]
    L.f2d_oracle_computesumandnorm.argtypes = [c_bp, c_dp, c_int, c_int, c_int,
                                               c_dp, c_dp]
    L.f2d_oracle_computenormmaxu.argtypes = [c_bp, c_dp, c_int, c_int, c_int,
                                             c_dp, c_dp]
    L.f2d_oracle_computekemaxu.argtypes = [c_bp, c_dp, c_dp, c_int, c_int, c_int,
                                           c_dp, c_dp]
    L.f2d_oracle_computekemaxuv.argtypes = [c_bp, c_dp, c_dp, c_int, c_int, c_int,
                                            c_dp, c_dp, c_dp]
    L.f2d_oracle_computekewithpsi.restype = c_dbl
    L.f2d_oracle_computekewithpsi.argtypes = [c_bp, c_dp, c_dp, c_int, c_int, c_int]
    L.f2d_oracle_computenorm.restype = c_dbl
    L.f2d_oracle_computenorm.argtypes = [c_bp, c_dp, c_int, c_int, c_int]
    L.f2d_oracle_computeinner.restype = c_dbl
    L.f2d_oracle_computeinner.argtypes = [c_bp, c_dp, c_dp, c_int, c_int, c_int]
    L.f2d_oracle_smoothtwicewitha.argtypes = [c_bp, c_dp, c_int, c_dp, c_dp, c_dbl,
                                              c_int, c_int, c_dp]
    L.f2d_oracle_computeresidualwitha.argtypes = [c_bp, c_dp, c_int, c_dp, c_dp, c_dp,
                                                  c_int, c_int]
    L.f2d_oracle_fillhalo.argtypes = [c_dp, c_int, c_int, c_int]
    L.f2d_oracle_interpolate.argtypes = [c_bp, c_bp, c_dp, c_int, c_dp,
                                         c_int, c_int, c_int, c_int]
    L.f2d_oracle_restrict.argtypes = [c_bp, c_dp, c_int, c_dp, c_int, c_int, c_int, c_int]
    L.f2d_oracle_coarsenmatrix.argtypes = [c_dp, c_dp, c_bp, c_bp, c_int,
                                           c_int, c_int, c_int, c_int]
    L.f2d_oracle_halotobuffer.argtypes = [c_dp] * 9 + [c_int, c_int, c_int]
    L.f2d_oracle_buffertohalo.argtypes = [c_dp] * 9 + [c_int, c_int, c_int]
    L.f2d_oracle_buffertodomain.argtypes = [c_dp, c_dp, c_int, c_int, c_int, c_int,
                                            c_int, c_int, c_int]
    L.f2d_oracle_smoothtridiag.argtypes = [c_bp, c_dp, c_int, c_dp, c_dp, c_int, c_int]
    L.f2d_oracle_axpy1.argtypes = [c_dp, c_dp, c_dbl, c_dp, ctypes.c_size_t]
    L.f2d_oracle_axpy2.argtypes = [c_dp, c_dp, c_dbl, c_dp, c_dp, ctypes.c_size_t]
    L.f2d_oracle_axpy3.argtypes = [c_dp, c_dbl, c_dp, c_dp, c_dp, ctypes.c_size_t]


# ---------------------------------------------------------------------------
# argument marshalling
# ---------------------------------------------------------------------------
_writebacks = []


def _d(a, write=False):
    """float64 C-contiguous pointer.  Operands of another dtype/layout are converted
    (f2py would do the same); a converted in/out operand is copied back after the
    call by _flush() (f2py's intent(inplace) would have cast the caller's array in
    place -- the values the caller sees are the same)."""
    if a.dtype != np.float64 or not a.flags.c_contiguous:
        src = a
        a = np.ascontiguousarray(a, dtype=np.float64)
        if write:
            _writebacks.append((src, a))
    return a.ctypes.data_as(c_dp), a


def _flush():
    while _writebacks:
        dst, tmp = _writebacks.pop()
        dst[...] = tmp


def _b(a):
    if a.dtype != np.int8 or not a.flags.c_contiguous:
        a = np.ascontiguousarray(a, dtype=np.int8)
    return a.ctypes.data_as(c_bp), a


def _A(A):
    """matrix [m][n][nd]; accepts the non-contiguous A[:, :, :5] view of a 9-array."""
    m, n, k = A.shape
    if A.dtype != np.float64:
        raise TypeError("A must be float64")
    s = A.strides
    if s[2] == 8 and s[1] % 8 == 0 and s[0] == n * s[1] and s[1] // 8 >= k:
        return A.ctypes.data_as(c_dp), s[1] // 8, A
    A = np.ascontiguousarray(A)
    return A.ctypes.data_as(c_dp), k, A


_NULL = ctypes.cast(None, c_dp)


class fortran_advection:
    """core/fortran_advection.f90"""

    @staticmethod
    def adv_upwind(msk, x, y, u, v, cst, nh, method, order):
        m, n = x.shape
        pm, _k0 = _b(msk)
        px, _k1 = _d(x)
        py, _k2 = _d(y, True)
        pu, _k3 = _d(u)
        pv, _k4 = _d(v)
        pc, _k5 = _d(np.asarray(cst))
        err = lib().f2d_oracle_adv_upwind(pm, px, py, pu, pv, _NULL, _NULL, pc,
                                          int(nh), int(method), int(order), m, n)
        if err:
            # the Fortran prints this and STOPs the process (fortran_advection.f90:30-34)
            raise SystemExit("NHALO = 3 is compulsory with UP5 / catastrophic ABORT!!!!")

    @staticmethod
    def adv_centered(msk, x, y, u, v, cst, nh, method, order):
        m, n = x.shape
        pm, _k0 = _b(msk)
        px, _k1 = _d(x)
        py, _k2 = _d(y, True)
        pu, _k3 = _d(u)
        pv, _k4 = _d(v)
        pc, _k5 = _d(np.asarray(cst))
        err = lib().f2d_oracle_adv_centered(pm, px, py, pu, pv, _NULL, _NULL, pc,
                                            int(nh), int(method), int(order), m, n)
        if err:
            raise SystemExit("NHALO = 3 is compulsory with UP5 / catastrophic ABORT!!!!")


class fortran_fluxes:
    """core/fortran_fluxes.f90 (advection + the two face-flux outputs)"""

    @staticmethod
    def adv_upwind(msk, x, y, u, v, xflx, yflx, cst, nh, method, order):
        m, n = x.shape
        pm, _k0 = _b(msk)
        px, _k1 = _d(x)
        py, _k2 = _d(y, True)
        pu, _k3 = _d(u)
        pv, _k4 = _d(v)
        pfx, _k6 = _d(xflx, True)
        pfy, _k7 = _d(yflx, True)
        pc, _k5 = _d(np.asarray(cst))
        err = lib().f2d_oracle_adv_upwind(pm, px, py, pu, pv, pfx, pfy, pc,
                                          int(nh), int(method), int(order), m, n)
        if err:
            raise SystemExit("NHALO = 3 is compulsory with UP5 / catastrophic ABORT!!!!")

    @staticmethod
    def adv_centered(msk, x, y, u, v, xflx, yflx, cst, nh, method, order):
        m, n = x.shape
        pm, _k0 = _b(msk)
        px, _k1 = _d(x)
        py, _k2 = _d(y, True)
        pu, _k3 = _d(u)
        pv, _k4 = _d(v)
        pfx, _k6 = _d(xflx, True)
        pfy, _k7 = _d(yflx, True)
        pc, _k5 = _d(np.asarray(cst))
        err = lib().f2d_oracle_adv_centered(pm, px, py, pu, pv, pfx, pfy, pc,
                                            int(nh), int(method), int(order), m, n)
        if err:
            raise SystemExit("NHALO = 3 is compulsory with UP5 / catastrophic ABORT!!!!")


class fortran_operators:
    """core/fortran_operators.f90"""

    @staticmethod
    def computeorthogradient(msk, psi, dx, dy, nh, u, v):
        m, n = psi.shape
        pm, _k0 = _b(msk)
        pp, _k1 = _d(psi)
        pu, _k2 = _d(u, True)
        pv, _k3 = _d(v, True)
        lib().f2d_oracle_computeorthogradient(pm, pp, float(dx), float(dy), int(nh),
                                              pu, pv, m, n)

    @staticmethod
    def celltocorner(xr, xp):
        m, n = xp.shape
        pr, _k0 = _d(xr)
        pp, _k1 = _d(xp, True)
        lib().f2d_oracle_celltocorner(pr, pp, m, n)

    @staticmethod
    def cornertocell(xp, xr):
        m, n = xr.shape
        pp, _k0 = _d(xp)
        pr, _k1 = _d(xr, True)
        lib().f2d_oracle_cornertocell(pp, pr, m, n)

    @staticmethod
    def add_diffusion(msk, trac, dx, nh, Kdiff, dtrac):
        m, n = trac.shape
        pm, _k0 = _b(msk)
        pt, _k1 = _d(trac)
        pd, _k2 = _d(dtrac, True)
        lib().f2d_oracle_add_diffusion(pm, pt, float(dx), int(nh), float(Kdiff), pd, m, n)

    @staticmethod
    def computenoslipsourceterm(msk, x, y, dx, dy, nh):
        m, n = x.shape
        pm, _k0 = _b(msk)
        px, _k1 = _d(x)
        py, _k2 = _d(y, True)
        return lib().f2d_oracle_computenoslipsourceterm(pm, px, py, float(dx), float(dy),
                                                        int(nh), m, n)

    @staticmethod
    def add_torque(msk, buoy, dx, nh, gravity, domega):
        m, n = buoy.shape
        pm, _k0 = _b(msk)
        pb, _k1 = _d(buoy)
        pd, _k2 = _d(domega, True)
        lib().f2d_oracle_add_torque(pm, pb, float(dx), int(nh), float(gravity), pd, m, n)


class fortran_diag:
    """core/fortran_diag.f90"""

    @staticmethod
    def computedotprod(msk, x, y, nh):
        m, n = x.shape
        pm, _k0 = _b(msk)
        px, _k1 = _d(x)
        py, _k2 = _d(y)
        return lib().f2d_oracle_computedotprod(pm, px, py, int(nh), m, n)

    @staticmethod
    def computemax(msk, x, nh):
        m, n = x.shape
        pm, _k0 = _b(msk)
        px, _k1 = _d(x)
        return lib().f2d_oracle_computemax(pm, px, int(nh), m, n)

    @staticmethod
    def computesum(msk, x, nh):
        m, n = x.shape
        pm, _k0 = _b(msk)
        px, _k1 = _d(x)
        return lib().f2d_oracle_computesum(pm, px, int(nh), m, n)

    @staticmethod
    def computesumandnorm(msk, x, nh):
        m, n = x.shape
        pm, _k0 = _b(msk)
        px, _k1 = _d(x)
        y, y2 = c_dbl(), c_dbl()
        lib().f2d_oracle_computesumandnorm(pm, px, int(nh), m, n,
                                           ctypes.byref(y), ctypes.byref(y2))
        return y.value, y2.value

    @staticmethod
    def computenormmaxu(msk, x, nh):
        m, n = x.shape
        pm, _k0 = _b(msk)
        px, _k1 = _d(x)
        y, y2 = c_dbl(), c_dbl()
        lib().f2d_oracle_computenormmaxu(pm, px, int(nh), m, n,
                                         ctypes.byref(y), ctypes.byref(y2))
        return y.value, y2.value

    @staticmethod
    def computekemaxu(msk, u, v, nh):
        m, n = u.shape
        pm, _k0 = _b(msk)
        pu, _k1 = _d(u)
        pv, _k2 = _d(v)
        ke, mx = c_dbl(), c_dbl()
        lib().f2d_oracle_computekemaxu(pm, pu, pv, int(nh), m, n,
                                       ctypes.byref(ke), ctypes.byref(mx))
        return ke.value, mx.value

    @staticmethod
    def computekemaxuv(msk, u, v, nh):
        m, n = u.shape
        pm, _k0 = _b(msk)
        pu, _k1 = _d(u)
        pv, _k2 = _d(v)
        ke, mu, mv = c_dbl(), c_dbl(), c_dbl()
        lib().f2d_oracle_computekemaxuv(pm, pu, pv, int(nh), m, n, ctypes.byref(ke),
                                        ctypes.byref(mu), ctypes.byref(mv))
        return ke.value, mu.value, mv.value

    @staticmethod
    def computekewithpsi(msk, omega, psi, nh):
        m, n = psi.shape
        pm, _k0 = _b(msk)
        po, _k1 = _d(omega)
        pp, _k2 = _d(psi)
        return lib().f2d_oracle_computekewithpsi(pm, po, pp, int(nh), m, n)


_scratch = {}


def _scratch_for(shape):
    s = _scratch.get(shape)
    if s is None:
        s = np.zeros(shape)
        _scratch[shape] = s
    return s


class fortran_multigrid:
    """core/gmg/fortran_multigrid.f90"""

    @staticmethod
    def smoothtwicewitha(msk, A, x, b, coef, yo=None):
        # `yo` ([3, n] in the reference) is the Fortran's rolling buffer; the oracle
        # keeps sweep 1 in a full-size scratch instead (see f2d_oracle.c).
        m, n = x.shape
        pm, _k0 = _b(msk)
        pA, nd, _k1 = _A(A)
        px, _k2 = _d(x, True)
        pb, _k3 = _d(b)
        ps, _k4 = _d(_scratch_for((m, n)), True)
        lib().f2d_oracle_smoothtwicewitha(pm, pA, nd, px, pb, float(coef), m, n, ps)

    @staticmethod
    def smoothtridiag(msk, A, x, b):
        m, n = x.shape
        pm, _k0 = _b(msk)
        pA, nd, _k1 = _A(A)
        px, _k2 = _d(x, True)
        pb, _k3 = _d(b)
        lib().f2d_oracle_smoothtridiag(pm, pA, nd, px, pb, m, n)

    @staticmethod
    def computeresidualwitha(msk, A, x, b, y):
        m, n = x.shape
        pm, _k0 = _b(msk)
        pA, nd, _k1 = _A(A)
        px, _k2 = _d(x)
        pb, _k3 = _d(b)
        py, _k4 = _d(y, True)
        lib().f2d_oracle_computeresidualwitha(pm, pA, nd, px, pb, py, m, n)

    @staticmethod
    def fillhalo(x, nh):
        m, n = x.shape
        px, _k0 = _d(x, True)
        lib().f2d_oracle_fillhalo(px, int(nh), m, n)

    @staticmethod
    def interpolate(msk1, msk2, x2, nh, x1):
        m1, n1 = x1.shape
        m2, n2 = x2.shape
        p1, _k0 = _b(msk1)
        p2, _k1 = _b(msk2)
        px2, _k2 = _d(x2)
        px1, _k3 = _d(x1, True)
        lib().f2d_oracle_interpolate(p1, p2, px2, int(nh), px1, m2, n2, m1, n1)

    @staticmethod
    def restrict(msk2, x1, nh, x2):
        m1, n1 = x1.shape
        m2, n2 = x2.shape
        p2, _k0 = _b(msk2)
        px1, _k1 = _d(x1)
        px2, _k2 = _d(x2, True)
        lib().f2d_oracle_restrict(p2, px1, int(nh), px2, m2, n2, m1, n1)

    @staticmethod
    def coarsenmatrix(Afine, msk1, msk2, nh):
        m1, n1 = msk1.shape
        m2, n2 = msk2.shape
        Af = np.ascontiguousarray(Afine, dtype=np.float64)
        assert Af.shape == (m1, n1, 9)
        Ac = np.zeros((m2, n2, 9))
        p1, _k0 = _b(msk1)
        p2, _k1 = _b(msk2)
        lib().f2d_oracle_coarsenmatrix(Af.ctypes.data_as(c_dp), Ac.ctypes.data_as(c_dp),
                                       p1, p2, int(nh), m1, n1, m2, n2)
        return Ac

    @staticmethod
    def computenorm(msk, x, nh):
        m, n = x.shape
        pm, _k0 = _b(msk)
        px, _k1 = _d(x)
        return lib().f2d_oracle_computenorm(pm, px, int(nh), m, n)

    @staticmethod
    def computeinner(msk, x, y, nh):
        m, n = x.shape
        pm, _k0 = _b(msk)
        px, _k1 = _d(x)
        py, _k2 = _d(y)
        return lib().f2d_oracle_computeinner(pm, px, py, int(nh), m, n)

    @staticmethod
    def halotobuffer(x, b0, b1, b2, b3, b4, b5, b6, b7):
        m, n = x.shape
        nh = b0.shape[0]
        ps = [_d(x)[0]] + [_d(b, True)[0] for b in (b0, b1, b2, b3, b4, b5, b6, b7)]
        lib().f2d_oracle_halotobuffer(*ps, nh, m, n)

    @staticmethod
    def buffertohalo(x, b0, b1, b2, b3, b4, b5, b6, b7):
        m, n = x.shape
        nh = b0.shape[0]
        ps = [_d(x, True)[0]] + [_d(b)[0] for b in (b0, b1, b2, b3, b4, b5, b6, b7)]
        lib().f2d_oracle_buffertohalo(*ps, nh, m, n)

    @staticmethod
    def buffertodomain(b, x, nh, m1, n1):
        mp, np_, m, n = b.shape
        pb, _k0 = _d(b)
        px, _k1 = _d(x, True)
        lib().f2d_oracle_buffertodomain(pb, px, int(nh), int(m1), int(n1), m, n, mp, np_)


def axpy1(out, x, c, d0):
    lib().f2d_oracle_axpy1(_d(out, True)[0], _d(x)[0], float(c), _d(d0)[0], out.size)


def axpy2(out, x, c, d0, d1):
    lib().f2d_oracle_axpy2(_d(out, True)[0], _d(x)[0], float(c), _d(d0)[0], _d(d1)[0],
                           out.size)


def axpy3(x, c, d0, d1, d2):
    lib().f2d_oracle_axpy3(_d(x, True)[0], float(c), _d(d0)[0], _d(d1)[0], _d(d2)[0],
                           x.size)
