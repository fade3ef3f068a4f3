"""Run the reference's Fortran SOURCE TEXT without a Fortran compiler.  TEST INFRASTRUCTURE,
build container only (reads /root/reference/core/*.f90; nothing is copied into the repo).

No gfortran exists in this image, so oracle/f2d_oracle.c was restated from the Fortran by
reading.  This module closes part of that gap mechanically: it translates the five kernel
files, statement by statement, into Python and executes the result, so that every routine of
the C restatement can be compared with what the reference's own source computes
(tests/test_oracle_vs_fortran_source.py, bit for bit on small arrays).  A slip of
transcription -- an index, a sign, a loop bound, the order of a sum -- shows up as a
difference.  What this is NOT: the gfortran binary.  The semantics of the language are
modelled here, by the same author:

  * default REAL literals (`37./60.`, `0.5`) are single precision and are evaluated in single
    precision until they meet a REAL*8 operand (numpy float32 / float64 scalars follow the
    same promotion rule); `d` exponents give doubles;
  * a scalar assignment converts to the declared type of the left-hand side (REAL*8, REAL,
    INTEGER by truncation, INTEGER*1);
  * INTEGER / INTEGER truncates towards zero; `x**n` with an integer n is a repeated product;
  * arrays are 1-based (or carry their declared lower bounds), x(j,i) is element [j-1][i-1]
    of the numpy array the caller passed (f2py's logical index mapping); an out-of-range
    write raises, an out-of-range read yields a poison value (NaN / -77) and is counted;
  * DATA statements fill in column-major order; no FMA contraction exists in Python.

Supported subset: SUBROUTINE / declarations / DO / IF-ELSEIF-ELSE / one-line IF / assignments
(element, whole-array and section) / CALL / DATA / WRITE (ignored) / STOP / RETURN, free- and
fixed-form continuation lines -- everything the five files use.
"""
import ast
import os
import re

import numpy as np

REFERENCE_CORE = "/root/reference/core"
FILES = {
    "fortran_advection": "fortran_advection.f90",
    "fortran_fluxes": "fortran_fluxes.f90",
    "fortran_operators": "fortran_operators.f90",
    "fortran_diag": "fortran_diag.f90",
    "fortran_multigrid": "gmg/fortran_multigrid.f90",
}


def available():
    return os.path.isfile(os.path.join(REFERENCE_CORE, FILES["fortran_multigrid"]))


class FortranStop(Exception):
    pass


# ---------------------------------------------------------------------------
# run-time support of the translated code
# ---------------------------------------------------------------------------
class FArr(object):
    """numpy array addressed with Fortran subscripts (declared lower bounds, inclusive slices)"""

    def __init__(self, a, lows=None):
        self.a = a
        self.lows = tuple(lows) if lows is not None else (1,)*a.ndim

    def _key(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        if len(idx) != self.a.ndim:
            raise IndexError("rank mismatch")
        out = []
        for k, (i, lo, n) in enumerate(zip(idx, self.lows, self.a.shape)):
            if isinstance(i, slice):
                start = 0 if i.start is None else int(i.start)-lo
                stop = n if i.stop is None else int(i.stop)-lo+1
                if start < 0 or stop > n:
                    raise IndexError("section out of bounds")
                out.append(slice(start, stop))
            else:
                j = int(i)-lo
                if j < 0 or j >= n:
                    raise IndexError("subscript %d of dimension %d out of bounds [%d, %d]" % (int(i), k+1, lo, lo+n-1))
                out.append(j)
        return tuple(out)

    oob_reads = 0     # reads past the end of an array (counted; they yield a poison value)

    def __getitem__(self, idx):
        try:
            return self.a[self._key(idx)]
        except IndexError:
            # The reference does read one line past its arrays in places (adv_centered updates
            # its running mask sums with msk(j+4,i) on the last row, fortran_advection.f90:263)
            # and never uses what it read.  Fortran does not check; here the read yields a
            # poison value, so a use of it would show in the outputs.
            if any(isinstance(i, slice) for i in (idx if isinstance(idx, tuple) else (idx,))):
                raise
            FArr.oob_reads += 1
            return self.a.dtype.type(-77) if self.a.dtype.kind == "i" else self.a.dtype.type(np.nan)

    def __setitem__(self, idx, value):
        self.a[self._key(idx)] = value


def f4(x):
    return np.float32(x)


def f8(x):
    return np.float64(x)


def _isint(x):
    return isinstance(x, (int, np.integer)) and not isinstance(x, bool)


def fdiv(a, b):
    if _isint(a) and _isint(b):
        q = abs(int(a))//abs(int(b))
        return q if (int(a) >= 0) == (int(b) >= 0) else -q
    return a/b


def fpow(a, b):
    if _isint(b):
        n = int(b)
        if n < 0:
            return 1/fpow(a, -n)
        r = a if n > 0 else (1 if _isint(a) else type(a)(1))
        for _ in range(n-1):
            r = r*a
        return r
    return a**b


def frange(a, b, c=1):
    a, b, c = int(a), int(b), int(c)
    return range(a, b+1, c) if c > 0 else range(a, b-1, c)


def cast_int(x):
    return int(x)          # truncation towards zero


INTRINSICS = {
    "abs": abs, "max": max, "min": min, "sqrt": np.sqrt, "log": np.log, "cosh": np.cosh, "exp": np.exp,
    "tanh": np.tanh, "int": cast_int, "dble": np.float64, "real": np.float32, "float": np.float32,
    "mod": lambda a, b: a-b*fdiv(a, b) if _isint(a) and _isint(b) else np.fmod(a, b),
    "sign": lambda a, b: abs(a) if b >= 0 else -abs(a),
}
CASTS = {"real8": "f8", "real4": "f4", "int": "cast_int", "int1": "np.int8"}
DTYPES = {"real8": np.float64, "real4": np.float32, "int": np.int64, "int1": np.int8}


# ---------------------------------------------------------------------------
# source -> logical lines
# ---------------------------------------------------------------------------
def logical_lines(text):
    out = []
    for raw in text.splitlines():
        line = raw.rstrip()
        if not line.strip():
            continue
        m = re.match(r"\s*!f2py\s+intent\(out\)\s*::\s*(.*)$", line, re.I)
        if m:
            out.append("f2pyout "+m.group(1).strip())
            continue
        if line.lstrip().startswith("!") or line[0] in "cC*":
            continue
        # inline comment (no string in these files contains '!', except inside write())
        if "!" in line and "write" not in line.lower():
            line = line[:line.index("!")].rstrip()
            if not line.strip():
                continue
        fixed_cont = re.match(r"^     [^\s0]", line) is not None and line[5] in "$&+"
        body = line[6:] if fixed_cont else line.strip()
        lead_cont = body.startswith("&")
        if lead_cont:
            body = body[1:]
        if out and (fixed_cont or lead_cont or out[-1].endswith("&")):
            prev = out[-1][:-1] if out[-1].endswith("&") else out[-1]
            out[-1] = prev.rstrip()+" "+body.strip()
        else:
            out.append(body.strip())
    return [ln[:-1].rstrip() if ln.endswith("&") else ln for ln in out]


# ---------------------------------------------------------------------------
# expressions
# ---------------------------------------------------------------------------
DOTOPS = {"ne": "!=", "eq": "==", "lt": "<", "le": "<=", "gt": ">", "ge": ">=", "and": " and ", "or": " or ",
          "not": " not ", "true": " True ", "false": " False "}
TOKEN = re.compile(r"\s*(?:(?P<name>[A-Za-z_]\w*)|(?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[eEdD][+-]?\d+)?)|"
                   r"(?P<op>\*\*|==|!=|<=|>=|/=|[-+*/(),:<>=])|(?P<str>'[^']*'|\"[^\"]*\"))")


class _Rewrite(ast.NodeTransformer):
    def visit_BinOp(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Div):
            return ast.copy_location(ast.Call(ast.Name("fdiv", ast.Load()), [node.left, node.right], []), node)
        if isinstance(node.op, ast.Pow):
            return ast.copy_location(ast.Call(ast.Name("fpow", ast.Load()), [node.left, node.right], []), node)
        return node


def expr(src, arrays):
    """Fortran expression -> Python expression text"""
    s = re.sub(r"\.(ne|eq|lt|le|gt|ge|and|or|not|true|false)\.", lambda m: " "+DOTOPS[m.group(1).lower()]+" ", src,
               flags=re.I)
    s = s.replace("/=", "!=")
    out, closers, pos = [], [], 0
    while pos < len(s):
        if s[pos].isspace():
            out.append(" ")
            pos += 1
            continue
        m = TOKEN.match(s, pos)
        if not m:
            kw = re.match(r"(and|or|not|True|False)\b", s[pos:])
            raise SyntaxError("cannot tokenise %r at %r" % (src, s[pos:pos+10])) if not kw else None
        pos = m.end()
        if m.group("name"):
            name = m.group("name")
            low = name.lower()
            if low in ("and", "or", "not") or name in ("True", "False"):
                out.append(" "+name+" ")
                continue
            nxt = re.match(r"\s*\(", s[pos:])
            if nxt and low in arrays:
                pos += nxt.end()
                out.append(low+"[")
                closers.append("]")
            elif nxt:
                pos += nxt.end()
                if low not in INTRINSICS:
                    raise SyntaxError("unknown function %s in %r" % (name, src))
                out.append("INTRINSICS['%s'](" % low)
                closers.append(")")
            else:
                out.append(low)
        elif m.group("num"):
            num = m.group("num")
            if re.fullmatch(r"\d+", num):
                out.append(num)
            elif re.search(r"[dD]", num):
                out.append("f8(%s)" % re.sub(r"[dD]", "e", num))
            else:
                out.append("f4(%s)" % num)
        elif m.group("op"):
            op = m.group("op")
            if op == "(":
                closers.append(")")
                out.append("(")
            elif op == ")":
                out.append(closers.pop())
            else:
                out.append(op)
        else:
            out.append(m.group("str"))
    text = "".join(out).strip()
    tree = ast.parse(text, mode="eval")
    tree = ast.fix_missing_locations(_Rewrite().visit(tree))
    return ast.unparse(tree)


# ---------------------------------------------------------------------------
# declarations
# ---------------------------------------------------------------------------
def split_top(s, sep=","):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == sep and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    parts.append(cur)
    return [p.strip() for p in parts if p.strip()]


DECL = re.compile(r"^(integer\*1|integer|real\*8|real|double precision)\s*(.*?)::\s*(.*)$", re.I)


def parse_decl(line, table):
    m = DECL.match(line)
    if not m:
        return False
    kind = {"integer*1": "int1", "integer": "int", "real*8": "real8", "real": "real4",
            "double precision": "real8"}[m.group(1).lower()]
    attrs, names = m.group(2), m.group(3)
    dims = None
    dm = re.search(r"dimension\s*\((.*)\)", attrs, re.I)
    if dm:
        dims = split_top(dm.group(1))
    for item in split_top(names):
        nm = re.match(r"([A-Za-z_]\w*)\s*(?:\((.*)\))?$", item)
        own = split_top(nm.group(2)) if nm.group(2) else dims
        table[nm.group(1).lower()] = (kind, own)
    return True


# ---------------------------------------------------------------------------
# subroutine -> python function
# ---------------------------------------------------------------------------
def translate_subroutine(name, args, body):
    table = {}
    stmts = []
    outs = []
    for ln in body:
        low = ln.lower()
        if low.startswith("f2pyout "):
            outs += [a.strip().lower() for a in ln[8:].split(",")]
            continue
        if low.startswith("implicit") or parse_decl(ln, table):
            continue
        stmts.append(ln)
    arrays = {k for k, (kind, dims) in table.items() if dims}
    args = [a.lower() for a in args]
    lines = ["def %s(%s):" % (name, ", ".join(args))]
    ind = 1

    def emit(s):
        lines.append("    "*ind+s)

    # dummy arrays: wrap; locals: allocate
    def bounds(dims):
        lows, sizes = [], []
        for d in dims:
            if ":" in d:
                lo, hi = d.split(":")
                lows.append(expr(lo, arrays))
                sizes.append("(%s)-(%s)+1" % (expr(hi, arrays), expr(lo, arrays)))
            else:
                lows.append("1")
                sizes.append(expr(d, arrays))
        return lows, sizes

    for a in args:
        if a in arrays:
            lows, _ = bounds(table[a][1])
            emit("%s = FArr(%s, (%s,))" % (a, a, ", ".join(lows)))
        elif a in table:
            emit("%s = %s(%s)" % (a, CASTS[table[a][0]], a))
    for k, (kind, dims) in table.items():
        if dims and k not in args:
            lows, sizes = bounds(dims)
            emit("%s = FArr(np.zeros((%s,), dtype=DTYPES['%s']), (%s,))" % (k, ", ".join(sizes), kind, ", ".join(lows)))
    scalars_out = [a for a in args if a not in arrays]
    ret = "return {%s}" % ", ".join("'%s': %s" % (a, a) for a in scalars_out)

    def assign(lhs, rhs):
        lhs = lhs.strip()
        r = expr(rhs, arrays)
        m = re.match(r"([A-Za-z_]\w*)\s*(\(.*\))?$", lhs)
        nm = m.group(1).lower()
        if m.group(2):
            return "%s = %s" % (expr(lhs, arrays), r)
        if nm in arrays:          # whole-array assignment
            return "%s.a[...] = %s" % (nm, r)
        if nm not in table:
            raise SyntaxError("undeclared %s" % nm)
        return "%s = %s(%s)" % (nm, CASTS[table[nm][0]], r)

    def simple(st):
        low = st.lower()
        if low.startswith("write") or low.startswith("print"):
            return "pass"
        if low == "stop":
            return "raise FortranStop()"
        if low == "return":
            return ret
        if low == "continue":
            return "pass"
        m = re.match(r"call\s+(\w+)\s*\((.*)\)$", st, re.I)
        if m:
            return "%s(%s)" % (m.group(1).lower(), ", ".join(
                (a.strip().lower()+".a" if a.strip().lower() in arrays else expr(a, arrays)) for a in split_top(m.group(2))))
        m = re.match(r"data\s+(\w+)\s*/(.*)/$", st, re.I)
        if m:
            nm = m.group(1).lower()
            vals = ", ".join(expr(v, arrays) for v in split_top(m.group(2)))
            return "%s.a[...] = np.array([%s], dtype=%s.a.dtype).reshape(%s.a.shape, order='F')" % (nm, vals, nm, nm)
        # assignment: first '=' at depth 0 that is not part of ==, <=, >=, /=
        depth = 0
        for k, ch in enumerate(st):
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "=" and depth == 0 and st[k+1:k+2] != "=" and st[k-1] not in "<>/=":
                return assign(st[:k], st[k+1:])
        raise SyntaxError("cannot translate %r" % st)

    def matching_paren(s, start):
        depth = 0
        for k in range(start, len(s)):
            if s[k] == "(":
                depth += 1
            elif s[k] == ")":
                depth -= 1
                if depth == 0:
                    return k
        raise SyntaxError("unbalanced %r" % s)

    for st in stmts:
        low = st.lower().strip()
        if re.match(r"end\s*do$", low) or re.match(r"end\s*if$", low):
            ind -= 1
            continue
        m = re.match(r"do\s+(\w+)\s*=\s*(.*)$", st, re.I)
        if m:
            emit("for %s in frange(%s):" % (m.group(1).lower(), ", ".join(expr(p, arrays) for p in split_top(m.group(2)))))
            ind += 1
            continue
        m = re.match(r"(else\s*if|elseif|if)\s*\(", st, re.I)
        if m:
            close = matching_paren(st, m.end()-1)
            cond = expr(st[m.end():close], arrays)
            rest = st[close+1:].strip()
            head = m.group(1).lower().replace(" ", "")
            if rest.lower() == "then":
                if head == "if":
                    emit("if %s:" % cond)
                    ind += 1
                else:
                    ind -= 1
                    emit("elif %s:" % cond)
                    ind += 1
            else:
                emit("if %s:" % cond)
                ind += 1
                emit(simple(rest))
                ind -= 1
            continue
        if low == "else":
            ind -= 1
            emit("else:")
            ind += 1
            continue
        emit(simple(st))
    emit(ret)
    # dimension arguments f2py would hide and fill from the array shapes
    infer = {}
    for a in args:
        if a in arrays and a not in outs:
            for axis, d in enumerate(table[a][1]):
                d = d.strip().lower()
                if d in args and d not in arrays and d not in infer:
                    infer[d] = (a, axis)
    META[name] = {"args": args, "infer": infer, "outs": outs,
                  "dims": {a: table[a][1] for a in args if a in arrays},
                  "kinds": {a: table[a][0] for a in args if a in table}}
    return "\n".join(lines)


META = {}


def translate_file(text):
    lines = logical_lines(text)
    subs, cur = [], None
    for ln in lines:
        m = re.match(r"subroutine\s+(\w+)\s*\((.*)\)$", ln, re.I)
        if m:
            cur = (m.group(1).lower(), split_top(m.group(2)), [])
            continue
        if re.match(r"end(\s+subroutine.*)?$", ln, re.I):
            if cur:
                subs.append(cur)
            cur = None
            continue
        if cur:
            cur[2].append(ln)
    return "\n\n".join(translate_subroutine(*s) for s in subs), [s[0] for s in subs]


_modules = {}


def module(key):
    """namespace {routine name: python function} of one reference Fortran file"""
    if key not in _modules:
        text = open(os.path.join(REFERENCE_CORE, FILES[key])).read()
        code, names = translate_file(text)
        ns = {"np": np, "FArr": FArr, "f4": f4, "f8": f8, "fdiv": fdiv, "fpow": fpow, "frange": frange,
              "cast_int": cast_int, "INTRINSICS": INTRINSICS, "DTYPES": DTYPES, "FortranStop": FortranStop}
        exec(compile(code, "<%s translated>" % FILES[key], "exec"), ns)
        ns["__source__"] = code
        ns["__routines__"] = names
        ns["__meta__"] = {n: META[n] for n in names}
        _modules[key] = ns
    return _modules[key]


def call(key, routine, **given):
    """call a translated routine the way f2py exposes it: arrays and scalars by name, the
    dimension arguments filled in from the array shapes; returns the final values of the
    scalar arguments (intent(out) results among them)"""
    ns = module(key)
    meta = ns["__meta__"][routine]
    values = []
    for a in meta["args"]:
        if a in given:
            values.append(given[a])
        elif a in meta["infer"]:
            arr, axis = meta["infer"][a]
            values.append(given[arr].shape[axis])
        else:
            values.append(0.)          # intent(out) scalar
    with np.errstate(all="ignore"):
        return ns[routine](*values)


def f2py_namespace(key):
    """{routine: callable} with the calling convention of the f2py module the reference
    builds from this file (build.py:12-39): dimension arguments hidden, intent(out) arguments
    returned, intent(inplace) arrays modified in place (an array of another dtype is converted
    and copied back, as f2py's intent(inplace) does)"""
    ns = module(key)

    def make(routine):
        meta = ns["__meta__"][routine]
        visible = [a for a in meta["args"] if a not in meta["infer"] and a not in meta["outs"]]

        def fn(*pos, **kw):
            given = dict(zip(visible, pos))
            given.update({k.lower(): v for k, v in kw.items()})
            back = []
            for a, dims in meta["dims"].items():
                if a in given:
                    want = DTYPES[meta["kinds"][a]]
                    arr = given[a]
                    if not isinstance(arr, np.ndarray) or arr.dtype != want:
                        conv = np.array(arr, dtype=want)
                        if isinstance(arr, np.ndarray):
                            back.append((arr, conv))
                        given[a] = conv
            sizes = {d: given[arr].shape[axis] for d, (arr, axis) in meta["infer"].items()}
            for a in meta["outs"]:
                if a in meta["dims"]:      # intent(out) array: allocated from its declared shape
                    shape = tuple(int(eval(expr(d, set()), dict(sizes))) for d in meta["dims"][a])
                    given[a] = np.zeros(shape, dtype=DTYPES[meta["kinds"][a]])
            values = [given[a] if a in given else (sizes[a] if a in sizes else 0.) for a in meta["args"]]
            with np.errstate(all="ignore"):
                res = ns[routine](*values)
            for dst, tmp in back:
                dst[...] = tmp
            got = [given[a] if a in meta["dims"] else res[a] for a in meta["outs"]]
            if not got:
                return None
            return got[0] if len(got) == 1 else tuple(got)
        fn.__name__ = routine
        return fn
    return {r: make(r) for r in ns["__routines__"]}


if __name__ == "__main__":
    import sys
    for k in ([a for a in sys.argv[1:] if a != "-v"] or FILES):
        m = module(k)
        print("# ---- %s: %s" % (k, ", ".join(m["__routines__"])))
        if "-v" in sys.argv:
            print(m["__source__"])
