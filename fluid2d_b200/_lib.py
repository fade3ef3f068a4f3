"""ctypes binding of libf2d_b200.so (the C ABI declared in include/f2d_b200.h).

The prototypes are read from the header itself, so the binding cannot drift from the
ABI.  There is NO fallback: if the shared library is missing or a symbol is absent the
import fails loudly -- the product never computes on the CPU.
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), "include", "f2d_b200.h")
LIBDIR = os.path.join(HERE, "lib")

_CTYPES = {
    "int": ctypes.c_int,
    "double": ctypes.c_double,
    "size_t": ctypes.c_size_t,
    "long long": ctypes.c_longlong,
    "unsigned int": ctypes.c_uint,
    "void": None,
    "f2d_stream_t": ctypes.c_void_p,
}


def parse_header(path=HEADER):
    """-> {name: (restype, [argtypes], [argnames])} for every function declared"""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(f2d_\w+)\s*\(([^;{}]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3)
        if "typedef" in ret:
            continue
        argtypes, argnames = [], []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                mm = re.match(r"(.*?)(\w+)$", a)
                ctype_s, aname = mm.group(1).strip(), mm.group(2)
                argtypes.append(_ctype(ctype_s))
                argnames.append(aname)
        protos[name] = (_ctype(ret), argtypes, argnames)
    return protos


def _ctype(s):
    s = s.replace("const", "").strip()
    s = " ".join(s.split())
    nptr = s.count("*")
    base = s.replace("*", "").strip()
    if nptr:
        if base == "char" and nptr == 1:
            return ctypes.c_char_p
        return ctypes.c_void_p          # every pointer travels as an address
    return _CTYPES[base]


class F2DError(RuntimeError):
    def __init__(self, code, msg, fn):
        RuntimeError.__init__(self, "%s failed (code %d): %s" % (fn, code, msg))
        self.code = code


class Library(object):
    """Loaded libf2d_b200.so; integer-returning entry points raise F2DError on failure."""

    ERR_NH, ERR_ARG, ERR_CUDA, ERR_DIVERGE = 1, 2, 3, 4

    def __init__(self, strict=False):
        name = "libf2d_b200_strict.so" if strict else "libf2d_b200.so"
        self.path = os.path.join(LIBDIR, name)
        if not os.path.exists(self.path):
            raise ImportError(
                "fluid2d_b200: %s not found -- build it with `python -m fluid2d_b200.build` "
                "(there is no CPU fallback)" % self.path)
        self.cdll = ctypes.CDLL(self.path)
        self.protos = parse_header()
        self.raw = {}
        for fn, (ret, argtypes, _names) in self.protos.items():
            f = getattr(self.cdll, fn)      # AttributeError if the .so lacks the symbol
            f.restype = ret
            f.argtypes = argtypes
            self.raw[fn] = f
            if ret is ctypes.c_int and fn not in ("f2d_abi_version", "f2d_mg_nlevels",
                                                  "f2d_mg_level_matrix_mode", "f2d_mg_slab_levels", "f2d_mg_tail_level",
                                                  "f2d_comm_rank", "f2d_comm_size"):
                setattr(self, fn[4:], self._checked(fn, f))
            else:
                setattr(self, fn[4:], f)
        if self.abi_version() != 1:
            raise ImportError("fluid2d_b200: ABI version mismatch")

    def _checked(self, fn, f):
        last_error = self.cdll.f2d_last_error
        last_error.restype = ctypes.c_char_p

        def call(*args):
            rc = f(*args)
            if rc != 0:
                raise F2DError(rc, last_error().decode(), fn)
            return rc
        call.__name__ = fn
        return call


_libs = {}


def lib(strict=False):
    if strict not in _libs:
        _libs[strict] = Library(strict)
    return _libs[strict]
