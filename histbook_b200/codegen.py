"""Lower a fill spec to one fused CUDA kernel (text) for NVRTC.

Replaces, for the fill hot path, the reference's instruction compiler and interpreter
(histbook/instr.py:241-381 ``walkdown``/``instructions``; histbook/fill.py:85-162 ``_fill``;
histbook/calc/__init__.py:327-357 ``calculate``) and ``Hist._postfill`` (histbook/hist.py:392-504):

* common-subexpression elimination over *all* histograms of a book is done here on the goal
  trees (``spec.tree_key``) -- the reference does it with ``CallGraphNode.grow`` (instr.py:67-91);
* every node becomes one SSA register statement of a per-event device function, evaluated in the
  dtype numpy would use (``ufunc.resolve_dtypes`` is asked at plan time, so NEP-50 weak scalars,
  float32 columns staying float32, int/float comparisons etc. follow numpy by construction);
* every histogram gets an accumulation plan: which (bin, column) receives which term, and where
  the bins live while the kernel runs (shared-memory privatised u32 / f64, or global float64 REDs).

The arithmetic *shape* (operand order, balanced sums, ``x*(1/3)`` instead of ``x/3`` ...) is
whatever histbook's ``expr.py``/``instr.totree`` produced; it is not re-derived here.
"""
import hashlib
import math
import struct

import numpy

from histbook_b200 import spec as hbspec

THREADS = 256
MAX_WIN = 128              # hb_template.cuh HB_MAX_WIN
MAX_AUX = 64               # hb_template.cuh HB_MAX_AUX
VEC = 4                    # events per thread per tile (hb_load4)

MAYBECONSTANTS = {"pi": numpy.pi, "Pi": numpy.pi, "e": numpy.e, "E": numpy.e,
                  "inf": numpy.inf, "Inf": numpy.inf, "infinity": numpy.inf, "Infinity": numpy.inf,
                  "nan": numpy.nan, "NaN": numpy.nan, "Nan": numpy.nan}     # expr.py:328-331

CTYPE = {
    "float64": "double", "float32": "float", "int64": "hb_i64", "uint64": "hb_u64",
    "int32": "int", "uint32": "hb_u32", "int16": "short", "uint16": "unsigned short",
    "int8": "signed char", "uint8": "unsigned char", "bool": "bool",
}
# dtype codes of the C ABI (include/hbfill.h, enum hb_dtype)
DTYPE_CODE = {"float64": 0, "float32": 1, "int64": 2, "int32": 3, "int16": 4, "int8": 5,
              "uint64": 6, "uint32": 7, "uint16": 8, "uint8": 9, "bool": 10}


class UnsupportedError(NotImplementedError):
    pass


def ctype(dtype):
    name = numpy.dtype(dtype).name
    if name not in CTYPE:
        raise UnsupportedError("dtype {0} is not supported by the B200 fill backend".format(name))
    return CTYPE[name]


def lit(value, dtype):
    """Exact C literal of ``value`` converted to ``dtype`` the way numpy converts a scalar."""
    dtype = numpy.dtype(dtype)
    with numpy.errstate(all="ignore"):
        v = numpy.array(value).astype(dtype)
    if dtype == numpy.float64:
        bits, = struct.unpack("<q", struct.pack("<d", float(v)))
        return "__longlong_as_double(%dLL)" % bits if bits != -2**63 else "__longlong_as_double((hb_i64)0x8000000000000000ULL)"
    if dtype == numpy.float32:
        bits, = struct.unpack("<i", struct.pack("<f", float(v)))
        return "__int_as_float(%d)" % bits if bits != -2**31 else "__int_as_float(HB_INT_MIN)"
    if dtype == numpy.bool_:
        return "true" if bool(v) else "false"
    if dtype.kind == "i":
        iv = int(v)
        if iv == -2**63:
            return "((hb_i64)0x8000000000000000ULL)"
        return "((%s)%dLL)" % (ctype(dtype), iv)
    if dtype.kind == "u":
        return "((%s)%dULL)" % (ctype(dtype), int(v))
    raise UnsupportedError(str(dtype))


class Value(object):
    """An SSA value of the per-event program."""
    __slots__ = ("c", "dtype", "const", "kind", "extra")

    def __init__(self, c=None, dtype=None, const=None, kind="arr", extra=None):
        self.c = c            # C expression (a variable name) or None for constants
        self.dtype = numpy.dtype(dtype) if dtype is not None else None
        self.const = const    # python value for ("const", v) leaves
        self.kind = kind      # arr | const | index | group
        self.extra = extra

    @property
    def isconst(self):
        return self.kind == "const"


def _pytype(v):
    """What ufunc.resolve_dtypes wants for a Python literal (NEP 50 weak scalars)."""
    if isinstance(v, bool):
        return numpy.dtype(numpy.bool_)
    if isinstance(v, int):
        return int
    if isinstance(v, float):
        return float
    raise UnsupportedError("constant {0!r} in arithmetic".format(v))


def _resolve(ufunc, args):
    sig = tuple(_pytype(a.const) if a.isconst else a.dtype for a in args) + (None,) * ufunc.nout
    try:
        return ufunc.resolve_dtypes(sig)
    except TypeError:
        # all-Python-scalar calls: numpy would use default dtypes
        sig = tuple(numpy.array(a.const).dtype if a.isconst else a.dtype for a in args) + (None,) * ufunc.nout
        return ufunc.resolve_dtypes(sig)


# unary float functions: library name -> (numpy ufunc, C name for double, C name for float)
_FLOAT1 = {
    "sqrt": (numpy.sqrt, "sqrt", "sqrtf"), "exp": (numpy.exp, "exp", "expf"), "exp2": (numpy.exp2, "exp2", "exp2f"),
    "expm1": (numpy.expm1, "expm1", "expm1f"), "log": (numpy.log, "log", "logf"), "log2": (numpy.log2, "log2", "log2f"),
    "log10": (numpy.log10, "log10", "log10f"), "log1p": (numpy.log1p, "log1p", "log1pf"),
    "sin": (numpy.sin, "sin", "sinf"), "cos": (numpy.cos, "cos", "cosf"), "tan": (numpy.tan, "tan", "tanf"),
    "arcsin": (numpy.arcsin, "asin", "asinf"), "arccos": (numpy.arccos, "acos", "acosf"), "arctan": (numpy.arctan, "atan", "atanf"),
    "sinh": (numpy.sinh, "sinh", "sinhf"), "cosh": (numpy.cosh, "cosh", "coshf"), "tanh": (numpy.tanh, "tanh", "tanhf"),
    "arcsinh": (numpy.arcsinh, "asinh", "asinhf"), "arccosh": (numpy.arccosh, "acosh", "acoshf"), "arctanh": (numpy.arctanh, "atanh", "atanhf"),
}
# bit-exact with numpy (IEEE-754 correctly rounded or exact); everything else in _FLOAT1 is "libm-class"
EXACT_FLOAT1 = {"sqrt"}
_ROUND1 = {"floor": (numpy.floor, "floor", "floorf"), "ceil": (numpy.ceil, "ceil", "ceilf"),
           "trunc": (numpy.trunc, "trunc", "truncf"), "rint": (numpy.rint, "rint", "rintf")}
_FLOAT2 = {"arctan2": (numpy.arctan2, "atan2", "atan2f"), "hypot": (numpy.hypot, "hypot", "hypotf"),
           "copysign": (numpy.copysign, "copysign", "copysignf")}
_CMP = {"numpy.equal": (numpy.equal, "=="), "numpy.not_equal": (numpy.not_equal, "!="),
        "numpy.less": (numpy.less, "<"), "numpy.less_equal": (numpy.less_equal, "<=")}
_ARITH = {"numpy.add": (numpy.add, "hb_add"), "numpy.subtract": (numpy.subtract, "hb_sub"),
          "numpy.multiply": (numpy.multiply, "hb_mul")}
_ERFLIKE = {"erf": "hb_erf", "erfc": "hb_erfc", "lgamma": "hb_lgamma", "gamma": "hb_gamma", "factorial": "hb_factorial"}


class Program(object):
    """Builds the per-event SSA program (with CSE) for one spec and one set of column dtypes."""

    def __init__(self, spec, coltypes):
        """coltypes: field name -> (numpy dtype, is_scalar).  Fields that are missing but are
        ``maybeconstants`` must have been added by the caller as float64 scalars (fill.py:133-134)."""
        self.spec = spec
        self.coltypes = coltypes
        self.cols = []          # [(name, dtype, is_scalar)] in slot order
        self.colslot = {}
        self.body = []          # C statements of the event function
        self.memo = {}
        self.counter = 0
        self.consts = []        # __constant__ array definitions (split edges)
        self.inexact = set()    # library functions used that are not bit-identical to numpy
        self.zeroed = {}        # (weight, weight^2) SSA names -> their NaN-zeroed forms (hist.py:422-427), shared by a book's histograms
        self.nonneg = set()     # SSA names known to be >= +0 or NaN (squares, absolute values): the fixed-point add skips the sign path
        self.square_of = {}     # SSA name of a float64 product x * x -> SSA name of x (the sorted strategy stages x only)
        self.always_ok = set()  # names of `ok` flags that can never become false (bin axes with underflow, overflow and NaN bins)
        self.bin_sites = {}     # SSA name of a float64 value -> closed-low bin axes computed on it: [{i, low, scale, numbins, shift, at}]
        self.split_sites = []   # split axes lowered as compare chains: what generate() needs to re-write them as table look-ups
        self.indicator = {}     # SSA name of where(c, 1, 0) -> C expression of c
        self.gated = {}         # SSA name of the float64 product where(c, 1, 0) * y -> (C expression of c, SSA name of y)
        self.zeroed_of = {}     # SSA name of a float64 weight -> SSA name of its NaN-zeroed form
        self.gate_of = {}       # SSA name of the NaN-zeroed form of a gated weight -> (C expression of c, SSA name of y)

    # ------------------------------------------------------------------ helpers
    def new(self, prefix="t"):
        self.counter += 1
        return "%s%d" % (prefix, self.counter)

    def emit(self, dtype, expr, prefix="t"):
        name = self.new(prefix)
        self.body.append("const %s %s = %s;" % (ctype(dtype), name, expr))
        return Value(name, dtype)

    def cast(self, value, dtype):
        """C expression of ``value`` as ``dtype`` (numpy 'safe' casts only: ints widen, int -> float, bool -> anything)."""
        dtype = numpy.dtype(dtype)
        if value.isconst:
            return lit(value.const, dtype)
        if value.dtype == dtype:
            return value.c
        if dtype == numpy.bool_:
            return "(%s != 0)" % value.c
        return "((%s)%s)" % (ctype(dtype), value.c)

    def tobool(self, value):
        if value.isconst:
            return "true" if value.const else "false"
        if value.dtype == numpy.bool_:
            return value.c
        return "(%s != 0)" % value.c

    # ------------------------------------------------------------------ leaves
    def column(self, name):
        if name not in self.colslot:
            if name not in self.coltypes:
                raise ValueError("required field {0} not found in fill arguments".format(repr(name)))
            dtype, is_scalar = self.coltypes[name]
            ctype(dtype)
            self.colslot[name] = len(self.cols)
            self.cols.append((name, numpy.dtype(dtype), bool(is_scalar)))
        slot = self.colslot[name]
        return Value("c%d" % slot, self.cols[slot][1])

    # ------------------------------------------------------------------ trees
    def lower(self, tree):
        kind = tree[0]
        if kind == "const":
            return Value(const=hbspec.decode_value(tree[1]), kind="const")
        key = hbspec.tree_key(tree)
        if key in self.memo:
            return self.memo[key]
        if kind in ("name", "pred"):
            out = self.column(tree[1])
        elif kind == "bconst":
            v = hbspec.decode_value(tree[2])
            dt = numpy.array(v).dtype                 # numpy.full(length, value), fill.py:112-120
            out = self.emit(dt, lit(v, dt))
        elif kind == "call":
            out = self.call(tree[1], [self.lower(x) for x in tree[2:]])
        else:
            raise AssertionError(tree)
        self.memo[key] = out
        return out

    def call(self, fcn, args):
        if fcn.startswith("histbook."):
            return self.index_op(fcn, args)

        if fcn in _ARITH:
            ufunc, helper = _ARITH[fcn]
            a, b, out = _resolve(ufunc, args)
            if fcn == "numpy.subtract" and out == numpy.bool_:
                raise TypeError("numpy boolean subtract, the `-` operator, is not supported")
            res = self.emit(out, "%s<%s>(%s, %s)" % (helper, ctype(out), self.cast(args[0], a), self.cast(args[1], b)))
            if (fcn == "numpy.multiply" and out.kind == "f" and not args[0].isconst and not args[1].isconst
                    and (args[0].c == args[1].c or (args[0].c in self.nonneg and args[1].c in self.nonneg))):
                self.nonneg.add(res.c)               # x * x (and products of such)
            if (fcn == "numpy.multiply" and out == numpy.float64 and not args[0].isconst and not args[1].isconst
                    and args[0].c == args[1].c and args[0].dtype == numpy.float64):
                self.square_of[res.c] = args[0].c
            if fcn == "numpy.multiply" and out == numpy.float64 and not args[0].isconst and not args[1].isconst:
                # a filter as weight, where(c, 1, 0) * y: the product is y (c true) or y * 0 (hist.py filter(); see stage_values)
                for ind, other in ((args[0], args[1]), (args[1], args[0])):
                    if ind.c in self.indicator and other.dtype == numpy.float64 and other.c not in self.indicator:
                        self.gated[res.c] = (self.indicator[ind.c], other.c)
                        break
            return res

        if fcn == "numpy.true_divide":
            a, b, out = _resolve(numpy.true_divide, args)
            ctype(out)
            return self.emit(out, "(%s / %s)" % (self.cast(args[0], a), self.cast(args[1], b)))

        if fcn in _CMP:
            ufunc, op = _CMP[fcn]
            a, b, out = _resolve(ufunc, args)
            return self.emit(numpy.bool_, "(%s %s %s)" % (self.cast(args[0], a), op, self.cast(args[1], b)))

        if fcn == "numpy.logical_and":
            return self.emit(numpy.bool_, "(%s && %s)" % (self.tobool(args[0]), self.tobool(args[1])))
        if fcn == "numpy.logical_or":
            return self.emit(numpy.bool_, "(%s || %s)" % (self.tobool(args[0]), self.tobool(args[1])))
        if fcn == "numpy.logical_not":
            return self.emit(numpy.bool_, "(!%s)" % self.tobool(args[0]))

        if fcn == "numpy.isin":
            # the right-hand side is a Python set (expr.py:185-193); numpy.isin turns a set into a 0-d
            # object array, so the reference's answer is False for every element (kept for parity).
            if args[1].isconst and isinstance(args[1].const, (set, frozenset)):
                return self.emit(numpy.bool_, "false")
            raise UnsupportedError("numpy.isin with a non-set right-hand side")

        if fcn in _FLOAT1:
            ufunc, cd, cf = _FLOAT1[fcn]
            a, out = _resolve(ufunc, args)
            if fcn not in EXACT_FLOAT1:
                self.inexact.add(fcn)
            name = cd if out == numpy.float64 else cf
            ctype(out)
            if out not in (numpy.float64, numpy.float32):
                raise UnsupportedError("{0} in dtype {1}".format(fcn, out))
            return self.emit(out, "%s(%s)" % (name, self.cast(args[0], a)))

        if fcn in _ROUND1:
            ufunc, cd, cf = _ROUND1[fcn]
            a, out = _resolve(ufunc, args)
            if out.kind in "iub":
                return self.emit(out, self.cast(args[0], a))         # numpy >= 2.1: integer input is returned as is
            if out not in (numpy.float64, numpy.float32):
                raise UnsupportedError("{0} in dtype {1}".format(fcn, out))
            return self.emit(out, "%s(%s)" % (cd if out == numpy.float64 else cf, self.cast(args[0], a)))

        if fcn in _FLOAT2:
            ufunc, cd, cf = _FLOAT2[fcn]
            a, b, out = _resolve(ufunc, args)
            if out not in (numpy.float64, numpy.float32):
                raise UnsupportedError("{0} in dtype {1}".format(fcn, out))
            if fcn != "copysign":
                self.inexact.add(fcn)
            return self.emit(out, "%s(%s, %s)" % (cd if out == numpy.float64 else cf, self.cast(args[0], a), self.cast(args[1], b)))

        if fcn in ("abs", "fabs"):
            a, out = _resolve(numpy.absolute, args)
            res = self.emit(out, "hb_abs<%s>(%s)" % (ctype(out), self.cast(args[0], a)))
            self.nonneg.add(res.c)
            return res
        if fcn == "conj":
            a, out = _resolve(numpy.conjugate, args)
            return self.emit(out, self.cast(args[0], a))
        if fcn == "sign":
            a, out = _resolve(numpy.sign, args)
            if out.kind == "u" or out == numpy.bool_:
                return self.emit(out, "(%s)(%s != 0)" % (ctype(out), self.cast(args[0], a)))
            return self.emit(out, "hb_sign<%s>(%s)" % (ctype(out), self.cast(args[0], a)))
        if fcn in ("max", "fmax", "min", "fmin"):
            ufunc = numpy.maximum if fcn in ("max", "fmax") else numpy.minimum
            a, b, out = _resolve(ufunc, args)
            helper = "hb_max" if fcn in ("max", "fmax") else "hb_min"
            if out == numpy.bool_:
                op = "||" if helper == "hb_max" else "&&"
                return self.emit(out, "(%s %s %s)" % (self.cast(args[0], a), op, self.cast(args[1], b)))
            return self.emit(out, "%s<%s>(%s, %s)" % (helper, ctype(out), self.cast(args[0], a), self.cast(args[1], b)))
        if fcn == "heaviside":
            if len(args) == 1:
                args = args + [Value(const=0.5, kind="const")]
            a, b, out = _resolve(numpy.heaviside, args)
            if out not in (numpy.float64, numpy.float32):
                raise UnsupportedError("heaviside in dtype {0}".format(out))
            return self.emit(out, "hb_heaviside<%s>(%s, %s)" % (ctype(out), self.cast(args[0], a), self.cast(args[1], b)))
        if fcn in ("isnan", "isinf", "isfinite"):
            ufunc = getattr(numpy, fcn)
            a, out = _resolve(ufunc, args)
            x = self.cast(args[0], a)
            if a.kind != "f":
                return self.emit(numpy.bool_, "true" if fcn == "isfinite" else "false")
            big = lit(numpy.inf, a)
            if fcn == "isnan":
                return self.emit(numpy.bool_, "(%s != %s)" % (x, x))
            if fcn == "isinf":
                return self.emit(numpy.bool_, "(hb_abs<%s>(%s) == %s)" % (ctype(a), x, big))
            return self.emit(numpy.bool_, "(hb_abs<%s>(%s) < %s)" % (ctype(a), x, big))
        if fcn in ("mod", "fmod"):
            a, b, out = _resolve(numpy.remainder, args)      # calc/__init__.py:95: both are numpy.remainder
            x, y = self.cast(args[0], a), self.cast(args[1], b)
            if out.kind == "f":
                if out not in (numpy.float64, numpy.float32):
                    raise UnsupportedError("remainder in dtype {0}".format(out))
                return self.emit(out, "hb_remainder(%s, %s)" % (x, y))
            if out.kind == "i":
                return self.emit(out, "hb_remainder_int<%s>(%s, %s)" % (ctype(out), x, y))
            if out.kind == "u":
                return self.emit(out, "hb_remainder_uint<%s>(%s, %s)" % (ctype(out), x, y))
            raise TypeError("remainder of booleans")
        if fcn == "pow":
            a, b, out = _resolve(numpy.power, args)
            x, y = self.cast(args[0], a), self.cast(args[1], b)
            if out.kind == "f":
                if out not in (numpy.float64, numpy.float32):
                    raise UnsupportedError("power in dtype {0}".format(out))
                self.inexact.add("pow")
                return self.emit(out, "%s(%s, %s)" % ("pow" if out == numpy.float64 else "powf", x, y))
            if args[1].isconst and args[1].const < 0:
                raise ValueError("Integers to negative integer powers are not allowed.")
            return self.emit(out, "hb_ipow<%s>(%s, %s)" % (ctype(out), x, y))
        if fcn in ("deg2rad", "rad2deg"):
            ufunc = getattr(numpy, fcn)
            a, out = _resolve(ufunc, args)
            if out not in (numpy.float64, numpy.float32):
                raise UnsupportedError("{0} in dtype {1}".format(fcn, out))
            # numpy loops: x * (pi/180) and x * (180/pi), constant formed in the loop dtype
            if out == numpy.float64:
                k = (numpy.pi / 180.0) if fcn == "deg2rad" else (180.0 / numpy.pi)
            else:
                k = (numpy.float32(numpy.pi) / numpy.float32(180.0)) if fcn == "deg2rad" else (numpy.float32(180.0) / numpy.float32(numpy.pi))
            return self.emit(out, "(%s * %s)" % (self.cast(args[0], a), lit(k, out)))
        if fcn in ("logaddexp", "logaddexp2"):
            a, b, out = _resolve(getattr(numpy, fcn), args)
            if out != numpy.float64:
                raise UnsupportedError("{0} in dtype {1}".format(fcn, out))
            self.inexact.add(fcn)
            return self.emit(out, "hb_%s(%s, %s)" % (fcn, self.cast(args[0], a), self.cast(args[1], b)))
        if fcn == "where":
            # numpy.where(c, yes, no): result dtype by ordinary promotion of yes/no (weak Python scalars)
            yes, no = args[1], args[2]
            if yes.isconst and no.isconst:
                out = numpy.result_type(numpy.array(yes.const).dtype, numpy.array(no.const).dtype)
                if isinstance(yes.const, bool) != isinstance(no.const, bool) or type(yes.const) is not type(no.const):
                    out = numpy.result_type(yes.const, no.const)
            else:
                out = numpy.result_type(yes.const if yes.isconst else yes.dtype, no.const if no.isconst else no.dtype)
            res = self.emit(out, "(%s ? %s : %s)" % (self.tobool(args[0]), self.cast(yes, out), self.cast(no, out)))
            if (yes.isconst and no.isconst and out.kind in "iu" and not isinstance(yes.const, bool) and not isinstance(no.const, bool)
                    and yes.const == 1 and no.const == 0):
                self.indicator[res.c] = self.tobool(args[0])
            return res
        if fcn in _ERFLIKE:
            x = args[0]
            dt = numpy.result_type(x.const if x.isconst else x.dtype, 1.0)
            if dt != numpy.float64:
                raise UnsupportedError("{0} of a {1} value (the reference mixes float32 and float64 inside its approximation)".format(fcn, dt))
            self.inexact.add(fcn)
            return self.emit(numpy.float64, "%s(%s)" % (_ERFLIKE[fcn], self.cast(x, numpy.float64)))

        # calc/__init__.py:353-357: no library entry ("round", "//", "^" ...) -> AssertionError at fill time
        raise AssertionError("no library entry for {0!r}".format(fcn))

    # ------------------------------------------------------------------ index ops (calc/__init__.py:149-325)
    def index_op(self, fcn, args):
        v = args[0]
        if v.isconst:
            raise UnsupportedError("axis expression that is a bare constant call")
        if fcn.startswith("histbook.bin"):
            f = fcn[len("histbook.bin"):]
            numbins, low, high = (a.const for a in args[1:4])
            # values - float(low): numpy picks the dtype; float32 stays float32, ints become float64
            r = numpy.subtract.resolve_dtypes((v.dtype, float, None))[2]
            if r not in (numpy.float64, numpy.float32):
                raise UnsupportedError("bin axis on a {0} value".format(v.dtype))
            scale = float(numbins) / float(high - low)
            name, ok = self.new("i"), self.new("ok")
            self.body.append("bool %s = true;" % ok)
            self.body.append("const int %s = hb_bin_axis<%s, %s>(%s, %s, %s, %d, %s);" % (
                name, ctype(r), ", ".join("true" if c != "_" and c != "H" else "false" for c in f),
                self.cast(v, r), lit(float(low), r), lit(scale, r), int(numbins), ok))
            if f[0] == "U" and f[1] == "O" and f[2] == "N":
                self.always_ok.add(ok)
            if r == numpy.float64 and v.dtype == numpy.float64 and f[3] == "L":
                self.bin_sites.setdefault(v.c, []).append({"i": name, "low": float(low), "scale": float(scale), "numbins": int(numbins),
                                                           "shift": 1 if f[0] == "U" else 0, "nanflow": f[2] == "N", "at": len(self.body) - 1})
            return Value(name, numpy.int32, kind="index", extra=ok)

        if fcn.startswith("histbook.intbin"):
            f = fcn[len("histbook.intbin"):]
            lo, hi = args[1].const, args[2].const
            shift = 1 if f[0] == "U" else 0
            a, b, r = numpy.add.resolve_dtypes((v.dtype, int, None))
            shifted = self.emit(r, "hb_add<%s>(%s, %s)" % (ctype(r), self.cast(v, r), lit(shift - lo, r)))
            name, ok = self.new("i"), self.new("ok")
            self.body.append("bool %s = true;" % ok)
            self.body.append("const int %s = hb_intbin_axis<%s, %s, %s>(%s, %d, %s);" % (
                name, ctype(r), "true" if f[0] == "U" else "false", "true" if f[1] == "O" else "false",
                shifted.c, int(hi - lo), ok))
            return Value(name, numpy.int32, kind="index", extra=ok)

        if fcn.startswith("histbook.split"):
            f = fcn[len("histbook.split"):]
            edges = args[1].const
            earr = numpy.asarray(edges)
            # numpy.digitize -> searchsorted compares in the common dtype of column and edges
            common = numpy.result_type(v.dtype, earr.dtype)
            if common.kind in "iub":
                cmp_t = numpy.dtype(numpy.int64)
            elif common in (numpy.float64, numpy.float32):
                cmp_t = numpy.dtype(numpy.float64)      # float32 -> float64 is exact, comparisons are unchanged
            else:
                raise UnsupportedError("split axis comparing in {0}".format(common))
            bits = numpy.array(edges).astype(cmp_t).view(numpy.int64)
            name, ok = self.new("i"), self.new("ok")
            isnan = "(%s != %s)" % (v.c, v.c) if v.dtype.kind == "f" else "false"
            flags = ", ".join("true" if c != "_" and c != "H" else "false" for c in f)
            self.body.append("bool %s = true;" % ok)
            if len(edges) <= 16:
                # few edges: the count of edges <= v as a chain of compares against literals -- the edges sit in the
                # instruction stream / literal pool, no indexed constant loads (which go through the MIO queue: 7 % of the
                # c4 kernel's stall samples with the bisection over a __constant__ array, profiles/r02_c4_ncu.txt)
                vv = self.new("sv")
                self.body.append("const %s %s = %s;" % (ctype(cmp_t), vv, self.cast(v, cmp_t)))
                lits = [lit(e, cmp_t) for e in numpy.array(edges).astype(cmp_t).tolist()]

                # (a decision tree of nested selects, one compare per edge and half the instructions on paper, compiles to
                #  branches: c4 11.2 instead of 12.1 Gev/s)
                count = " + ".join("(int)(%s <= %s)" % (e, vv) for e in lits)
                onedge = " || ".join("(%s == %s)" % (e, vv) for e in lits)
                self.body.append("const int %s = hb_split_index<%s>(%s, %s, %s, %d, %s);" % (
                    name, flags, count, ("(%s)" % onedge) if f[3] in "_H" else "false", isnan, len(edges), ok))
                if v.dtype == numpy.float64 and cmp_t == numpy.float64:
                    self.split_sites.append({"at": len(self.body) - 1, "name": name, "ok": ok, "v": v.c, "vv": vv, "flags": flags,
                                             "closedhigh": f[3] in "_H", "edges": [float(e) for e in numpy.array(edges).astype(cmp_t).tolist()],
                                             "slow": self.body[-1][len("const int %s = " % name):]})
            else:
                cname = self.new("hb_edges")
                self.consts.append("__constant__ hb_i64 %s[%d] = {%s};" % (
                    cname, len(edges), ", ".join(("%dLL" % b) if b != -2**63 else "(-9223372036854775807LL - 1)" for b in bits.tolist())))
                self.body.append("const int %s = hb_split_axis<%s, %s>(%s, %s, %s, %d, %s);" % (
                    name, ctype(cmp_t), flags, self.cast(v, cmp_t), isnan, cname, len(edges), ok))
            return Value(name, numpy.int32, kind="index", extra=ok)

        if fcn == "histbook.cut":
            name, ok = self.new("i"), self.new("ok")
            self.body.append("bool %s = true;" % ok)
            self.body.append("const int %s = hb_cut_axis<%s>(%s, %s);" % (name, ctype(v.dtype), v.c, ok))
            return Value(name, numpy.int32, kind="index", extra=ok)

        if fcn.startswith("histbook.groupbin"):
            f = fcn[len("histbook.groupbin"):]
            binwidth, origin = args[1].const, args[2].const
            # calc/__init__.py:158-172, every step in the dtype of the first product
            if origin == 0:
                r = numpy.multiply.resolve_dtypes((v.dtype, float, None))[2]
                k = self.emit(r, "(%s * %s)" % (self.cast(v, r), lit(1.0 / float(binwidth), r)))
            else:
                r = numpy.subtract.resolve_dtypes((v.dtype, float, None))[2]
                k = self.emit(r, "(%s - %s)" % (self.cast(v, r), lit(float(origin), r)))
                k = self.emit(r, "(%s * %s)" % (k.c, lit(1.0 / float(binwidth), r)))
            if r not in (numpy.float64, numpy.float32):
                raise UnsupportedError("groupbin axis on a {0} value".format(v.dtype))
            fl, ce = ("floor", "ceil") if r == numpy.float64 else ("floorf", "ceilf")
            if f[1] == "L":
                k = self.emit(r, "%s(%s)" % (fl, k.c))
            else:
                k = self.emit(r, "%s(%s)" % (ce, k.c))
                k = self.emit(r, "(%s - %s)" % (k.c, lit(1, r)))
            k = self.emit(r, "(%s * %s)" % (k.c, lit(float(binwidth), r)))
            if origin != 0:
                k = self.emit(r, "(%s + %s)" % (k.c, lit(float(origin), r)))
            key = self.emit(numpy.float64, "((double)%s)" % k.c, prefix="gk")
            return Value(key.c, numpy.float64, kind="group", extra={"nanflow": f[0] == "N", "keytype": "float", "dtype": r})

        if fcn == "histbook.groupby":
            if v.dtype.kind == "f":
                key = self.emit(numpy.float64, "((double)%s)" % v.c, prefix="gk")
                return Value(key.c, numpy.float64, kind="group", extra={"nanflow": True, "keytype": "float", "dtype": v.dtype})
            key = self.emit(numpy.int64, "((hb_i64)%s)" % v.c, prefix="gk")
            return Value(key.c, numpy.int64, kind="group", extra={"nanflow": True, "keytype": "int", "dtype": v.dtype})

        raise AssertionError("no library entry for {0!r}".format(fcn))


# ================================================================================================
# accumulation plans
# ================================================================================================

class HistPlan(object):
    """Where each histogram's bins live during the kernel and which term goes to which column."""

    def __init__(self, hid, h):
        self.hid = hid
        self.shape = tuple(h["shape"])
        self.ncol = self.shape[-1]
        self.nbins = int(numpy.prod(self.shape[:-1], dtype=numpy.int64)) if len(self.shape) > 1 else 1
        self.index = []          # [(Value index, totbins)]
        self.groups = []         # [Value group]
        self.terms = []          # unique accumulated terms: C expression (double) or None (= count of 1)
        self.term_nonneg = []    # per term: provably >= 0 (or NaN)
        self.dest = []           # per term: list of destination columns
        self.strategy = "global"
        self.smem_f64_off = None   # byte offset of the [nbins][slots] float64 rows
        self.win_off = None        # "window" strategy: byte offset of the runtime-sized hot window region
        self.win_table_off = None  # byte offset of the per-row interval table (8 bytes per row)
        self.win_split = None      # the first `win_split` fixed axes form the row index, the rest the in-row index
        self.win_box = False       # the window is a box [lo_k, lo_k + n_k) on every fixed axis (run-time parameters p.win[...]) instead of per-row intervals
        self.gcap = 0              # group histograms: slab slots [0, gcap) are privatised in shared memory
        self.replicas = 1          # statically privatised copies per block (lane & (replicas - 1) picks one): fewer same-bin lanes per atomic
        self.skip_zero = False     # filter-derived weight: rows whose every term is exactly 0.0 are not accumulated
        self.fx_set = set()        # float64 terms kept as exact 128-bit fixed point (hb_template.cuh HbFx) instead of float64
        self.fx_index = {}         # term index -> index of its scale in p.win[2 + ...]
        self.term_src = []         # per term: SSA name when the term is that float64 value itself, else None
        self.sgroup = None         # "sorted" strategy: the SortGroup this histogram belongs to
        self.sslot = {}            # ... term index -> index into the group's value list

    def add_term(self, expr, col, nonneg=False, src=None):
        for t, e in enumerate(self.terms):
            if e == expr:
                self.dest[t].append(col)
                return
        self.terms.append(expr)
        self.term_nonneg.append(bool(nonneg))
        self.term_src.append(src)
        self.dest.append([col])

    @property
    def f64_terms(self):
        return [t for t, e in enumerate(self.terms) if e is not None]

    @property
    def count_term(self):
        for t, e in enumerate(self.terms):
            if e is None:
                return t
        return None

    # Shared-memory layout of a privatised histogram (or of its hot window), slot-major: for every float64 term one
    # block of `nrows` float64 values (updated with atomicAdd(double) = LDS.64 + ATOMS.CAST.SPIN.64) or -- terms in
    # `fx_set` -- four blocks of `nrows` u32 limbs (exact fixed point, native ATOMS.ADD); then `nrows` native u32
    # counters for an unweighted count.  Slot-major because consecutive instructions of a warp then spread over all
    # banks (c2: 127 Gev/s bin-major, 153 with hand-rolled 128-bit CAS pairs on bin-major rows, 156 slot-major).
    def row_layout(self, options=None):
        """(term index per float64 slot, whether a u32 count block follows)."""
        return list(self.f64_terms), self.count_term is not None

    def slot_units(self):
        """Size of every float64 slot in 8-byte units per row: 1 (float64) or 2 (four u32 limbs)."""
        return [2 if t in self.fx_set else 1 for t in self.f64_terms]

    @property
    def fx(self):
        return len(self.fx_set) > 0

    def row_bytes(self, options=None):
        return 8 * sum(self.slot_units()) + (4 if self.count_term is not None else 0)

    @property
    def nrows(self):
        return self.nbins * max(self.gcap, 1) * self.replicas

    @property
    def win_rows(self):
        tots = [t for v, t in self.index]
        return int(numpy.prod(tots[:self.win_split], dtype=numpy.int64)) if self.win_split else 1

    def win_table_bits(self, window_bytes, options=None):
        """Width of a row-table entry: 32 bits when first index and length fit 8 bits each and the offset 16."""
        if self.win_trail <= 255 and window_bytes // max(self.row_bytes(options), 1) <= 65535 and (options or {}).get("win_table32", True):
            return 32
        return 64

    @property
    def win_trail(self):
        tots = [t for v, t in self.index]
        return int(numpy.prod(tots[self.win_split:], dtype=numpy.int64))

    def smem_bytes(self, options=None):
        return self.nrows * self.row_bytes(options)


def _mulexpr(prog, a, b):
    """C expression (double) of numpy's ``a * b`` for two Values, then cast to float64 (add.at target)."""
    ta, tb, out = _resolve(numpy.multiply, [a, b])
    return "((double)hb_mul<%s>(%s, %s))" % (ctype(out), prog.cast(a, ta), prog.cast(b, tb))


def plan_hist(prog, hid, h):
    """hist.py:392-456 restated as: linear index, row mask, and one (term -> columns) table."""
    hp = HistPlan(hid, h)
    for a in h["group"]:
        hp.groups.append(prog.lower(a["goal"]))
    for a in h["fixed"]:
        if a["goal"] is None:
            continue
        hp.index.append((prog.lower(a["goal"]), int(a["totbins"])))

    w = h["weight"]
    wz = w2z = None
    if w is not None and "w" in w:
        wv, w2v = prog.lower(w["w"]), prog.lower(w["w2"])
        if wv.isconst:
            raise UnsupportedError("constant weight expression")
        if wv.dtype.kind == "f" and (wv.c, w2v.c) in prog.zeroed:
            wz, w2z = prog.zeroed[(wv.c, w2v.c)]             # the same weight in another histogram of the book
        elif wv.dtype.kind == "f":
            # hist.py:422-427: rows with NaN weight contribute 0 to every column
            sel = prog.emit(numpy.bool_, "(%s != %s)" % (wv.c, wv.c))
            wz = prog.emit(wv.dtype, "(%s ? %s : %s)" % (sel.c, lit(0.0, wv.dtype), wv.c))
            w2z = prog.emit(w2v.dtype, "(%s ? %s : %s)" % (sel.c, lit(0.0, w2v.dtype), w2v.c))
            if w2v.c in prog.nonneg:
                prog.nonneg.add(w2z.c)
            if prog.square_of.get(w2v.c) == wv.c and wv.dtype == numpy.float64:
                prog.square_of[w2z.c] = wz.c         # (nan ? 0 : w * w) == (nan ? 0 : w) * (nan ? 0 : w), bit for bit
            prog.zeroed[(wv.c, w2v.c)] = (wz, w2z)
            if wv.dtype == numpy.float64:
                prog.zeroed_of.setdefault(wv.c, wz.c)
                if wv.c in prog.gated:
                    prog.gate_of[wz.c] = prog.gated[wv.c]
        else:
            wz, w2z = wv, w2v

    if w is not None and "w" in w and "where" in hbspec.dumps(w["w"]):
        hp.skip_zero = True          # filter: where(cond, 1, 0) [* weight] is exactly 0 for rejected rows

    for p in h["profile"]:
        sx, sx2 = prog.lower(p["sumx"]), prog.lower(p["sumx2"])
        for val, col in ((sx, p["sumwx"]), (sx2, p["sumwx2"])):
            if val.isconst:
                raise UnsupportedError("constant profile expression")
            nonneg = False
            src = None
            if w is None:
                expr = "((double)%s)" % val.c               # sumx * 1
                nonneg = val.c in prog.nonneg
                src = val.c if val.dtype == numpy.float64 else None
            elif "const" in w:
                expr = _mulexpr(prog, val, Value(lit(float(w["const"]), numpy.float64), numpy.float64))
            else:
                expr = _mulexpr(prog, val, wz)
            hp.add_term(expr, col, nonneg, src)

    if w is None:
        hp.add_term(None, h["sumw"])
    elif "const" in w:
        hp.add_term(lit(float(w["const"]), numpy.float64), h["sumw"])
        hp.add_term(lit(float(w["const2"]), numpy.float64), h["sumw2"])
    else:
        hp.add_term("((double)%s)" % wz.c, h["sumw"], False, wz.c if wz.dtype == numpy.float64 else None)
        hp.add_term("((double)%s)" % w2z.c, h["sumw2"], w2z.c in prog.nonneg, w2z.c if w2z.dtype == numpy.float64 else None)
    return hp


def hist_index(hp):
    """(C expression of the row-major linear bin index over the fixed axes, hist.py:402-403; the axes' `ok` flags)."""
    conds = []
    if hp.index:
        idx = hp.index[0][0].c
        conds.append(hp.index[0][0].extra)
        for v, totbins in hp.index[1:]:
            idx = "(%s) * %d + %s" % (idx, totbins, v.c)
            conds.append(v.extra)
    else:
        idx = "0"
    return idx, conds


# ================================================================================================
# the "sorted" strategy of big books
# ================================================================================================

SORT_THREADS = 1024        # one block per SM
SORT_EPT = "auto"          # events per thread and tile: 3 where everything still fits the shared memory, else 2
SORT_SUM = "chunks"        # how the sorted events of a tile are summed: "chunks" (balanced: a contiguous chunk per lane) | "bins" (sort_split lanes per bin);
                           # c4, one GPU, 1e9 events: chunks of 64 lanes per group 12.1 Gev/s, of 128 lanes 11.9, bins 11.75 (profiles/r02_sweeps.txt)
SORT_MAX_BINS = 8192
SORT_ALIAS_TAILS = True    # the tails of the chunked sums travel through the array of sorted event numbers (27 KB less for c4)
SORT_PIPE = False          # producer / consumer warps with double-buffered tiles (_emit_kernel_sorted_pipe); "auto" = when it fits
SORT_MAX_GROUPS = 12       # (key, rank) tickets of a tile live in registers: SORT_EPT x groups of them per thread
SMEM_MAX = 227 * 1024


class SortGroup(object):
    """Histograms of a book that share one linear bin index (and row mask): their float64 terms are not added event
    by event with atomics; per tile the events are counting-sorted by that index (ONE native shared-memory atomic per
    event: the ticket), and every lane of the group's warps then owns a bin: it sums the bin's run of sorted events
    with plain loads and adds and adds the result to the block's float64 table with one plain, conflict-free
    read-modify-write per tile.  Replaces 152 ``numpy.add.at`` calls per fill of c4 (hist.py:439-456).
    (A balanced variant -- every lane a contiguous chunk of the sorted events, runs that end inside the chunk flushed on
    the spot, the rest handed on as "tails" -- was built first: every iteration some lane of a warp flushes, so the
    whole warp walks the 30-instruction flush path nearly every time, twice the instructions and 2 shared-memory
    wavefronts per event more; profiles/r02_c4_ncu.txt.)"""

    def __init__(self, gid, idx, conds, nbins):
        self.gid = gid
        self.idx = idx
        self.conds = conds
        self.nbins = nbins
        self.members = []
        self.values = []          # distinct float64 term expressions (C, evaluated inside the event function)
        self.srcs = []            # per value: SSA name or None
        self.dests = []           # per value: [(hist id, column, ncol)]
        self.count_dests = []     # [(hist id, column, ncol)] fed by the number of events per bin
        self.stage = []           # per value: ("E", k) staged slot | ("sq", k) square of staged slot k
        self.off = {}             # shared-memory byte offsets: acc, cnt, tick, sorted, tkey, tval

    def add(self, hp):
        self.members.append(hp)
        hp.sgroup = self
        for t, e in enumerate(hp.terms):
            if e is None:
                self.count_dests.extend((hp.hid, col, hp.ncol) for col in hp.dest[t])
                continue
            if e not in self.values:
                self.values.append(e)
                self.srcs.append(hp.term_src[t])
                self.dests.append([])
            j = self.values.index(e)
            hp.sslot[t] = j
            self.dests[j].extend((hp.hid, col, hp.ncol) for col in hp.dest[t])

    def smem_bytes(self, tile, lanes, chunks=False, alias=False):
        nv = len(self.values)
        nvp = (nv + 1) // 2 * 2 if chunks else nv                         # chunks: rows of the table are [bin][values], padded to 16 bytes
        tails = (4 * lanes + lanes * 8 * max(nvp, 2)) if chunks else 0      # tail keys, tail values
        if alias:
            # the tails travel through the group's array of sorted event numbers (nobody reads it any more by then)
            return (8 * nvp * self.nbins + 4 * self.nbins + 4 * (self.nbins + 4) + max(2 * tile + 4 * lanes + 32, tails + 32) + 64) // 16 * 16 + 32
        return (8 * nvp * self.nbins + 4 * self.nbins + 4 * (self.nbins + 4) + 2 * tile + 4 * lanes + 32 + tails + 64) // 16 * 16 + 16


def sort_groups(plans, options):
    """Candidate sort groups: histograms without group axes and with at least one float64 term, keyed by (linear
    index expression, row mask); unweighted histograms with the same key join (their counts are the tickets)."""
    mode = options.get("sort", "auto")
    if not mode or options.get("deterministic", False):
        return []
    if mode == "auto" and _unroll(plans, options) != 0:
        return []                                          # small books: the per-event strategies are faster (c2, c3)
    forced = options.get("strategy", {})
    by_key = {}
    order = []
    for hp in plans:
        if hp.groups or forced.get(hp.hid, forced.get(str(hp.hid), forced.get("*"))) not in (None, "sorted"):
            continue
        if not (0 < hp.nbins <= SORT_MAX_BINS) or not hp.f64_terms:
            continue
        key = hist_index(hp)
        key = (key[0], tuple(key[1]))
        if key not in by_key:
            by_key[key] = SortGroup(len(order), key[0], list(key[1]), hp.nbins)
            order.append(by_key[key])
        by_key[key].add(hp)
    for hp in plans:                                       # count-only histograms ride along
        if hp.groups or hp.f64_terms or hp.sgroup is not None or not (0 < hp.nbins <= SORT_MAX_BINS):
            continue
        if forced.get(hp.hid, forced.get(str(hp.hid), forced.get("*"))) not in (None, "sorted"):
            continue
        key = hist_index(hp)
        key = (key[0], tuple(key[1]))
        if key in by_key:
            by_key[key].add(hp)
    return order


def stage_values(prog, groups, gate=False, max_flags=0, flags=None):
    """Which float64 values of the event function are staged in shared memory per tile (E[k][event]): every distinct
    term of the groups, except
      * squares x * x of a value that is staged itself (recomputed by the summing lane), and
      * ``gate``: the NaN-zeroed form of a filtered weight where(c, 1, 0) * y whose NaN-zeroed y is staged itself -- it is
        that staged value where c holds and 0 elsewhere (c true: 1.0 * y is y; c false: 0.0 * y is a zero of either sign,
        or NaN for an infinite or NaN y, which the NaN-zeroing turns into 0; the sign of a zero never shows in a sum that
        starts from +0.0), so ONE flag bit per event travels in the spare high bits of the sorted event numbers instead of
        eight more bytes per event through shared memory (c4: 24 -> 16 staged bytes per event, one gather less per event
        and group).  ``flags`` collects the C expressions of the conditions, at most ``max_flags`` of them.
    g.stage, per value of group g: ("E", k) staged slot k | ("sq", k) its square | ("gate", k, f) slot k where flag f is
    set | ("gsq", k, f) its square where flag f is set."""
    all_srcs = set(src for g in groups for src in g.srcs if src is not None)
    memo = {}
    flags = [] if flags is None else flags

    def derived(src):
        """src is x * x of a value x that is staged itself (not derived in turn: x**4 of c3 stages x**2)"""
        if src not in memo:
            base = prog.square_of.get(src)
            memo[src] = base is not None and base in all_srcs and not derived(base)
        return memo[src]

    def gated(src):
        """(flag number, SSA name of the staged NaN-zeroed base) if src travels as a flag, else None"""
        if not gate or src not in prog.gate_of:
            return None
        cond, base = prog.gate_of[src]
        bz = prog.zeroed_of.get(base)
        if bz is None or bz == src or bz not in all_srcs or bz in prog.gate_of or derived(bz):
            return None
        if cond not in flags:
            if len(flags) >= max_flags:
                return None
            flags.append(cond)
        return flags.index(cond), bz
    staged = []                 # C expressions, slot order
    names = {}                  # SSA name -> slot
    for g in groups:
        for e, src in zip(g.values, g.srcs):
            if src is not None and (derived(src) or gated(src) is not None):
                continue
            if e not in staged:
                staged.append(e)
                if src is not None:
                    names[src] = len(staged) - 1
    for g in groups:
        g.stage = []
        for e, src in zip(g.values, g.srcs):
            if e in staged:
                g.stage.append(("E", staged.index(e)))
            elif gated(src) is not None:
                f, bz = gated(src)
                g.stage.append(("gate", names[bz], f))
            else:
                base = prog.square_of[src]
                if gated(base) is not None:
                    f, bz = gated(base)
                    g.stage.append(("gsq", names[bz], f))
                else:
                    g.stage.append(("sq", names[base]))
    return staged


def _pipe_group_fixed(g):
    """bytes of a sort group's tables that live for the whole kernel (pipelined variant): [bin][values] sums + counts"""
    nvp = (len(g.values) + 1) // 2 * 2
    return (8 * nvp * g.nbins + 15) // 16 * 16 + (4 * g.nbins + 15) // 16 * 16


def _pipe_group_tick(g):
    return (4 * (g.nbins + 1) + 15) // 16 * 16


def _pipe_pad(tile):
    """halfwords of padding behind every lane's chunk of sorted event numbers: a lane's chunk starts 8 x (an odd number)
    bytes after its neighbour's, so that the summing warp reads four event numbers per lane with ONE conflict-free LDS.64"""
    chunk = tile // 32
    pad = 0
    while (chunk + pad) % 8 != 4:
        pad += 1
    return pad


def _pipe_sorted_bytes(tile, g):
    """a group's array of sorted event numbers (padded chunks + the dump slot of masked events); the summing warp re-uses
    it for the tails of its lanes (32 keys + 32 rows of sums)"""
    nvp = (len(g.values) + 1) // 2 * 2
    return (max(2 * (tile + 32 * _pipe_pad(tile) + 8), 128 + 32 * 8 * nvp) + 15) // 16 * 16


def _unroll(plans, options):
    """How the four events a thread owns per tile are laid out in code: 4 = vector loads + four inlined copies of the event
    function; 1 = vector loads, one copy in a loop; 0 = one copy, one coalesced scalar load per column and event (big
    books: c4's 64 histograms are 600 KB of SASS and 110 registers when unrolled, 200 KB and 60 registers this way)."""
    return int(options.get("unroll", 0 if len(plans) > 16 else 4))


def choose_strategies(plans, options, prog=None, layout=None):
    """Shared-memory privatisation for the histograms that fit, global REDs for the rest.  Returns (bytes of shared
    memory, hot-window histogram or None); with ``prog`` given, big books may get sort groups, described in ``layout``."""
    det = bool(options.get("deterministic", False))
    forced = options.get("strategy", {})
    # big books: histograms that share a bin index are summed per tile from events sorted by that index (SortGroup)
    sgroups, staged, sort_bytes = [], [], 0
    pipe = False
    sort_cap = int(options.get("sort_smem", SMEM_MAX - 1024))
    sort_tile, sort_ept = 0, 2
    sort_lanes = options.get("sort_lanes", "auto")
    group_slots = int(options.get("group_slots", 16))
    if prog is not None:
        sgroups = sort_groups(plans, options)
        sgroups.sort(key=lambda g: -len(g.values) * len(g.members))

        def direct_need(slots):
            """shared memory the histograms outside the groups would like (fixed point for every float64 term)"""
            rest = [hp for hp in plans if hp.sgroup is None or id(hp.sgroup) not in set(id(g) for g in sgroups)]
            return 16 + sum((hp.nbins * (slots if hp.groups else 1) * (16 * len(hp.f64_terms) + (4 if hp.count_term is not None else 0)) + 15) // 16 * 16
                            for hp in rest)
        while sgroups:
            sgroups = sgroups[:int(options.get("sort_max_groups", SORT_MAX_GROUPS))]
            sort_chunks = options.get("sort_sum", SORT_SUM) == "chunks"
            # (filtered weights as flag bits beside the 16-bit sorted event numbers: the chunked sums of the phased kernel only)
            sort_flags = []
            ebits = (int(options.get("sort_threads", SORT_THREADS)) * (3 if options.get("sort_ept", SORT_EPT) == "auto" else int(options.get("sort_ept", SORT_EPT))) - 1).bit_length()
            gate = bool(options.get("sort_gate", True)) and sort_chunks and not bool(options.get("sort_pipe", SORT_PIPE)) and ebits < 16
            staged = stage_values(prog, sgroups, gate=gate, max_flags=16 - ebits, flags=sort_flags)
            # the block's warps are dealt out to the groups (4 warps = 128 lanes per group for c4's 8 groups)
            if sort_lanes == "auto":
                lanes = 32
                # (chunked sums: 128 lanes per group where the block has them -- c4's 8 groups then keep all 32 warps busy in the
                #  summing phase: 17.2 against 16.1 Gev/s with 64, profiles/r02_sweeps.txt; in round 2's second session, with the
                #  event phase 40 % longer, it made no difference)
                while lanes * 2 <= min(int(options.get("sort_max_lanes", 128)) if sort_chunks else 256, int(options.get("sort_threads", SORT_THREADS)) // len(sgroups)):
                    lanes *= 2
            else:
                lanes = int(sort_lanes)
            # pipelined variant: the last len(sgroups) warps of the block sum tile k (one warp per group) while the others
            # evaluate tile k + 1; tick counters, sorted event numbers and staged values exist twice.  Where that does not fit,
            # the same groups are tried with one buffer set and phases behind block barriers before a group is given up.
            want_pipe = bool(options.get("sort_pipe", SORT_PIPE)) and sort_chunks and len(sgroups) <= 12
            auto_lanes = lanes
            fits = False
            # events per thread and tile: bigger tiles mean fewer run ends per event in the summing phase and fewer barriers
            # (c4: 11.5 / 13.9 / 14.6 Gev/s for tiles of 1024 / 2048 / 3072); "auto" takes 3 where EVERYTHING still fits
            ept_opt = options.get("sort_ept", SORT_EPT)
            ept_list = [3, 2] if ept_opt == "auto" else [int(ept_opt)]
            sort_threads = int(options.get("sort_threads", SORT_THREADS))
            for pipe, sort_ept in [(pp, ee) for pp in ([True, False] if want_pipe else [False]) for ee in ept_list]:
                sort_tile, lanes = sort_threads * sort_ept, auto_lanes
                if pipe:
                    lanes = 32
                    pipe_threads = sort_threads
                    pipe_prod = pipe_threads - 32 * len(sgroups)
                    sort_tile = pipe_prod * sort_ept
                    if pipe_prod < 256 or sort_tile > 32768:
                        continue
                all_fit = False
                for slots in (group_slots, min(group_slots, 8)):
                    if pipe:
                        sort_bytes = (sum(_pipe_group_fixed(g) for g in sgroups)
                                      + 2 * (sum(_pipe_group_tick(g) + _pipe_sorted_bytes(sort_tile, g) for g in sgroups) + 16 + 8 * len(staged) * sort_tile) + 64)
                    else:
                        sort_bytes = sum(g.smem_bytes(sort_tile, lanes, sort_chunks, sort_chunks and bool(options.get("sort_alias_tails", SORT_ALIAS_TAILS))) for g in sgroups) + 8 * len(staged) * sort_tile + 64
                    if sort_bytes + direct_need(slots) <= sort_cap:
                        all_fit = True
                        break
                if sort_ept != ept_list[-1] and not all_fit:
                    continue                               # a bigger tile is not worth a table pushed out to global memory
                if sort_bytes + 16 * 1024 <= sort_cap:
                    fits = True
                    break
            if fits:
                sort_lanes, group_slots = lanes, slots
                break
            sgroups = sgroups[:-1]                         # does not fit: the last group goes back to the per-event strategies
        kept = set(id(g) for g in sgroups)
        for i, g in enumerate(sgroups):
            g.gid = i
        for hp in plans:
            if hp.sgroup is not None and id(hp.sgroup) not in kept:
                hp.sgroup, hp.sslot = None, {}
            if hp.sgroup is not None:
                hp.strategy = "sorted"
        if not sgroups:
            staged, sort_bytes, pipe = [], 0, False
    if not sgroups:
        sort_lanes, group_slots = 64, int(options.get("group_slots", 16))
    # Which float64 terms are accumulated as exact fixed point (16 bytes per bin, native u32 atomics, ~55 instructions,
    # no retries) rather than with float64 CAS loops (8 bytes, ~10 instructions, but several times the shared-memory
    # wavefronts, and one more attempt for every other lane of the warp that is in the same bin):
    #   fx = False  none;  True  all (always so in deterministic mode);  "mix"  every second term;  [slots]  explicit;
    #   "auto" (default), statically privatised histograms of small books only:
    #          tables up to 1024 rows: all terms   (Hist(bin("x",100), weight="w"): 55 -> 198 Gev/s),
    #          larger ones: all but the first      (c3: 116 -> 152 Gev/s; 53x53 weighted: 154 -> 183; all: 167),
    #          hot windows: all but the first as well, but only in the "dense" variant of the kernel, which the fill
    #          driver picks when the smaller window still holds 95 % of the events (c2: 156 -> 171 Gev/s although the
    #          window shrinks from 12.7 k to 8.5 k bins; all terms: the window halves and the gain is gone).
    #          The CAS path saturates the LSU wavefront pipe with the issue slots 20-40 % busy, the fixed-point path is
    #          issue-bound: on larger tables a mix loads both pipes.  Measurements: profiles/r01_sweeps.txt.
    mode = True if det else options.get("fx", "auto")
    # (statically privatised tables share 96 KB so that two blocks fit an SM; all-fixed-point kernels and big books run one
    #  1024-thread block per SM and may take 200 KB)
    budget = int(options.get("smem_budget", (200 if (mode is True or _unroll(plans, options) == 0) else 96) * 1024))
    if sgroups:
        budget = min(budget, sort_cap - sort_bytes)
    for hp in plans:
        # group axes: the first `group_slots` slab slots (keys in first-seen order) live in shared memory --
        # without this every event of such a histogram is a contended global RED (c4: 8 per event)
        hp.gcap = group_slots if hp.groups else 0
        hp.replicas = 1 if hp.groups else int(options.get("replicas", 1))
    for hp in plans:
        terms = list(hp.f64_terms)
        if hp.strategy == "sorted":
            terms = []
        elif mode == "mix":
            terms = terms[1::2]
        elif mode == "auto":
            terms = terms if hp.nrows <= 1024 else terms[1:]
        elif isinstance(mode, (list, tuple)):          # explicit slot numbers (tuning)
            terms = [t for j, t in enumerate(terms) if j in mode]
        elif not mode:
            terms = []
        hp.fx_set = set(terms)
    total = sum(len(hp.fx_set) for hp in plans)
    big = _unroll(plans, options) == 0
    if (mode == "auto" and total > int(options.get("fx_auto_max_terms", MAX_WIN - 2 if big else 8))) or (mode is not True and total > MAX_WIN - 2):
        # middle-sized books: the fixed-point code costs more registers and instructions than it saves while the
        # event function is unrolled four times; big books run one event at a time (see _unroll) and take it again
        # (c4, 40 float64 terms on 50..100-bin axes: 5.2 -> 5.9 Gev/s from the single copy, 6.4 with fixed point)
        for hp in plans:
            hp.fx_set = set()
    order = sorted(plans, key=lambda hp: hp.smem_bytes(options))
    used = 16 if any(hp.fx_set for hp in plans) else 0     # word 0: how many terms missed their fixed-point window (HB_FXMISS)
    for hp in order:
        if hp.strategy == "sorted":
            continue
        want = forced.get(hp.hid, forced.get(str(hp.hid), forced.get("*")))
        need = hp.smem_bytes(options)
        if mode == "auto" and hp.fx_set and (want == "smem" or want is None) and used + need > budget:
            hp.fx_set = set()                          # fits as float64 only
            need = hp.smem_bytes(options)
        if want == "global":
            hp.strategy = "global"
        elif (want == "smem" or want is None) and used + need <= budget:
            hp.strategy = "smem"
        else:
            hp.strategy = "global"
        if hp.strategy == "smem":
            hp.smem_f64_off = used
            used += (need + 15) // 16 * 16
    # Hot-window privatisation: ONE histogram that is too large for shared memory keeps the bins where most
    # events land in shared memory; the rest goes to global REDs.  The window is *ragged*: the leading fixed
    # axes are flattened into a row index, and every row owns one contiguous interval of the (flattened)
    # trailing axes, placed at run time (fillable.choose_ragged) from what the histogram already holds.
    # For c2 the populated region of (x+y, r) is a V (|x+y| <= sqrt(2) r): at 6400 bins a rectangle covers
    # 80 % of the events, per-row intervals 94.5 % (tools/window_coverage.py).
    window = None
    if options.get("window", True) and not sgroups:
        max_rows = int(options.get("win_max_rows", 4096))
        max_bytes = int(options.get("win_max_hist_bytes", 16 << 20))
        cands = []
        for hp in plans:
            if hp.strategy != "global" or hp.groups or not hp.index or hp.nbins * hp.ncol * 8 > max_bytes:
                continue
            if forced.get(hp.hid, forced.get(str(hp.hid), forced.get("*"))) not in (None, "window"):
                continue
            tots = [t for v, t in hp.index]
            split = None
            for k in range(len(tots) - 1, -1, -1):
                rows = int(numpy.prod(tots[:k], dtype=numpy.int64)) if k else 1
                trail = int(numpy.prod(tots[k:], dtype=numpy.int64))
                if rows <= max_rows and trail <= 65535:
                    split = k
                    break
            if split is not None:
                hp.win_split = split
                cands.append(hp)
        if cands:
            window = min(cands, key=lambda hp: hp.nbins)
            window.strategy = "window"
            used = (used + 15) // 16 * 16
            window.win_table_off = used
            used += (window.win_rows * 8 + 15) // 16 * 16
            window.win_off = used
        elif options.get("box", True) and not det:
            # Histograms beyond that (c5: 203^3 x 2 float64 = 134 MB): the window is a box, [lo_k, lo_k + n_k) on every
            # axis, placed at run time around the bulk of what the histogram already holds (fillable.choose_box).  A
            # 24^3 box of c5 is 221 KB and takes 45 % of N(0,1)^3 events out of the L2 RED path, which is what bounds
            # that kernel (227e9 float64 REDs/s, the device's scattered-RED rate: profiles/atomics_microbench_r01.txt).
            boxes = [hp for hp in plans if hp.strategy == "global" and not hp.groups and len(hp.index) >= 1 and hp.nbins * hp.ncol * 8 > max_bytes
                     and forced.get(hp.hid, forced.get(str(hp.hid), forced.get("*"))) in (None, "window") and len(hp.index) <= 6]
            if boxes:
                window = max(boxes, key=lambda hp: hp.nbins)
                window.strategy = "window"
                window.win_box = True
                window.win_split = 0
                used = (used + 15) // 16 * 16
                window.win_table_off = used
                window.win_off = used
    shared = {}
    for hp in plans:
        if hp.strategy == "global" and not det:      # (deterministic mode: global histograms add to exact accumulators too)
            hp.fx_set = set()
        if hp.strategy == "window" and mode == "auto":
            # two variants, chosen at fill time from the coverage the window reaches (fillable.fill_hists)
            hp.fx_set = set(hp.f64_terms[1:]) if options.get("win_dense", False) else set()
        if det and len(hp.f64_terms) > 0 and len(hp.fx_set) != len(hp.f64_terms):
            raise UnsupportedError("deterministic mode: every float64 term must be fixed point")
        for t in hp.f64_terms:
            if t in hp.fx_set:
                # one scale (p.win[2 + index]) and one decode per distinct VALUE: in a book many histograms accumulate
                # the same weight, weight^2, profile expression (c4: 48 terms, 12 distinct)
                hp.fx_index[t] = shared.setdefault(hp.terms[t], len(shared))
    nfx = len(shared)
    if nfx > MAX_WIN - 2:
        raise UnsupportedError("more than {0} fixed-point terms in one fill".format(MAX_WIN - 2))
    if not any(hp.strategy in ("smem", "window") for hp in plans) and not any(hp.fx for hp in plans):
        used = 0
    if sgroups and pipe:
        # pipelined variant: [acc | cnt of every group | tick of every group, buffer 0, + dump word | the same, buffer 1 | dump]
        # (zeroed at kernel start), then per buffer [sorted event numbers of every group | staged values]
        used = (used + 15) // 16 * 16
        for g in sgroups:
            g.off["acc"] = used
            used += (8 * ((len(g.values) + 1) // 2 * 2) * g.nbins + 15) // 16 * 16
        for g in sgroups:
            g.off["cnt"] = used
            used += (4 * g.nbins + 15) // 16 * 16
        tick0 = used
        for g in sgroups:
            g.off["tick"] = used
            used += _pipe_group_tick(g)
        tick_dump = used                                  # masked tickets of this buffer add here
        used += 16
        tick_stride = used - tick0
        used += tick_stride                               # buffer 1
        dump_off = used
        used += 16
        zero_end = used
        tile0 = used
        for g in sgroups:
            g.off["sorted"] = used
            used += _pipe_sorted_bytes(sort_tile, g)
        e_off = used
        used += 8 * len(staged) * sort_tile
        tile_stride = used - tile0
        used += tile_stride                               # buffer 1
        sort_layout = {"groups": sgroups, "staged": staged, "zero_end": zero_end, "dump_off": dump_off, "e_off": e_off, "tile": sort_tile,
                       "lanes": 32, "chunks": True, "threads": pipe_threads, "ept": sort_ept,
                       "pipe": {"prod": pipe_prod, "tick_stride": tick_stride, "tile_stride": tile_stride, "tick_dump": tick_dump,
                                "pad": _pipe_pad(sort_tile)}}
        if layout is not None:
            layout.update(sort_layout)
    elif sgroups:
        # layout behind the per-event tables: [acc | cnt | tick of every group] (zeroed at kernel start), then the
        # per-tile scratch [sorted event numbers | tails of every group], then the staged values E[k][event]
        used = (used + 15) // 16 * 16
        for g in sgroups:
            g.off["acc"] = used
            nvp = (len(g.values) + 1) // 2 * 2 if options.get("sort_sum", SORT_SUM) == "chunks" else len(g.values)
            used += (8 * nvp * g.nbins + 15) // 16 * 16
        for g in sgroups:
            g.off["cnt"] = used
            used += (4 * g.nbins + 15) // 16 * 16
        for g in sgroups:
            g.off["tick"] = used
            used += (4 * (g.nbins + 1) + 15) // 16 * 16
        dump_off = used                                   # where masked rows of count-only histograms and masked tickets add (never read)
        used += 16
        zero_end = used
        for g in sgroups:
            g.off["sorted"] = used
            region = (2 * sort_tile + 15) // 16 * 16 + 4 * sort_lanes + 32     # (+ room for the chunk padding of HB_SPOS and the dump slot of masked events)
            tails_bytes = (4 * sort_lanes + 15) // 16 * 16 + 8 * sort_lanes * max((len(g.values) + 1) // 2 * 2, 2)
            if options.get("sort_sum", SORT_SUM) == "chunks" and options.get("sort_alias_tails", SORT_ALIAS_TAILS):
                # tail keys and tail values of the group's lanes re-use the array of sorted event numbers: 27 KB of c4's shared
                # memory, which buys tiles of 3072 events (sort_ept 3)
                g.off["tkey"] = used
                g.off["tval"] = used + (4 * sort_lanes + 15) // 16 * 16
                used += (max(region, tails_bytes) + 15) // 16 * 16
            elif options.get("sort_sum", SORT_SUM) == "chunks":
                used += region
                g.off["tkey"] = used
                used += (4 * sort_lanes + 15) // 16 * 16
                g.off["tval"] = used
                used += 8 * sort_lanes * max((len(g.values) + 1) // 2 * 2, 2)
            else:
                used += region
        e_off = used
        used += 8 * len(staged) * sort_tile
        sort_layout = {"groups": sgroups, "staged": staged, "zero_end": zero_end, "dump_off": dump_off, "e_off": e_off, "tile": sort_tile,
                       "lanes": sort_lanes, "chunks": options.get("sort_sum", SORT_SUM) == "chunks", "threads": int(options.get("sort_threads", SORT_THREADS)),
                       "ept": sort_ept, "flags": sort_flags, "ebits": ebits}
        if layout is not None:
            layout.update(sort_layout)
    return used, window


# ================================================================================================
# kernel text
# ================================================================================================

class Kernel(object):
    """Result of codegen: CUDA source + everything the host needs to launch it."""

    def __init__(self):
        self.source = None
        self.name = "hb_fill_kernel"
        self.cols = None            # [(name, dtype, is_scalar)]
        self.plans = None           # [HistPlan] (fill kernels)
        self.smem_bytes = 0
        self.threads = THREADS
        self.tile = THREADS * VEC
        self.inexact = set()
        self.key = None
        self.group_aux = {}         # fill kernel: hist id -> aux slot of its dense slot map (this fill's key combinations -> slab slot)
        self.goal_aux = {}          # fill kernel: tree key of a group goal -> aux slot of the goal's key table of this fill
        self.group_leader = {}      # fill kernel: hist id -> first histogram with the same group goals (its slot lookup may be shared)
        self.window = None          # hot-window histogram: dict(hid, bpb, wmax, totbins, ncol, col, nf)
        self.group_goals = []       # keys kernel: [(tree key, info dict)] in aux order
        self.fx_terms = []          # [(i, hist id, term index)] of the fixed-point terms; scale of term i in p.win[2 + i]
        self.nfx = 0
        self.fx_aux = None
        self.times_aux = None
        self.naux = 0
        self.deterministic = False
        self.exact = {}
        self.sorted = None          # "sorted" strategy: {"tile", "lanes", "ept", "groups": [...], "staged"}


def _colarg(slot, dtype, is_scalar, loaded):
    src = ("s%d" % slot) if is_scalar else loaded
    if dtype == numpy.bool_:
        return "(%s != 0)" % src
    return src


def _loadtype(dtype):
    return ctype(dtype) if dtype != numpy.bool_ else "unsigned char"


def _emit_kernel(prog, event_lines, smem, merge_lines, threads, minblocks, window=None, prefetch=False, unroll=4, times_aux=None, dynamic=False, evict_first=False, red_evict_last=False):
    """The hand-written skeleton around the generated per-event function: persistent grid-stride loop
    over tiles of HB_THREADS*4 events, one vector load per column per thread (tail / unaligned input
    takes the scalar path), shared-memory zeroing before and block merge after."""
    src = []
    vcols_early = [slot for slot, (name, dtype, is_scalar) in enumerate(prog.cols) if not is_scalar]
    if evict_first:
        src.append("#define HB_STREAM_EVICT_FIRST 1")
    if red_evict_last:
        src.append("#define HB_RED_EVICT_LAST 1")
    src.append('#include "hb_template.cuh"')
    src.append("#define HB_THREADS %d" % threads)
    src.append("#define HB_TILE (HB_THREADS * 4)")
    src.append("#define HB_FXMISS (reinterpret_cast<hb_u32*>(hb_smem))")
    src.extend(prog.consts)
    src.append("")
    sig = ["const HbParams& p", "unsigned char* hb_smem", "const bool live"]
    for slot, (name, dtype, is_scalar) in enumerate(prog.cols):
        sig.append("const %s c%d /* %s */" % (ctype(dtype), slot, name))
    src.append("__device__ __forceinline__ void hb_event(%s)\n{" % ", ".join(sig))
    src.extend("    " + line for line in event_lines)
    src.append("}\n")

    src.append('extern "C" __global__ void __launch_bounds__(HB_THREADS, %d) hb_fill_kernel(const HbParams p)\n{' % minblocks)
    src.append("    extern __shared__ __align__(16) unsigned char hb_smem[];")

    def stamp(which):
        # option block_times (measurement only): per block [start, end of the event loop, end of the merge] in ns
        if times_aux is not None:
            src.append("    if (threadIdx.x == 0 && p.aux[%d] != nullptr) { hb_u64 t; asm volatile(\"mov.u64 %%0, %%%%globaltimer;\" : \"=l\"(t)); "
                       "reinterpret_cast<hb_u64*>(const_cast<void*>(p.aux[%d]))[blockIdx.x * 3 + %d] = t; }" % (times_aux, times_aux, which))
    stamp(0)
    if dynamic:
        # hand-out state (see the tile loop): chunks 0 and 1 of this block are static
        src.append("    __shared__ hb_u64 hb_ring[4];")
        src.append("    __shared__ hb_u32 hb_reads[4];")
        src.append("    __shared__ hb_u32 hb_ticket;")
        src.append("    if (threadIdx.x == 0) {")
        src.append("        const hb_i64 nt = (p.n + HB_TILE - 1) / HB_TILE;")
        src.append("        const hb_i64 t0 = (%s) + blockIdx.x, t1 = t0 + gridDim.x;" % (
            "p.vec_ok ? (nt * 15 / 16) / gridDim.x * gridDim.x : 0" if (dynamic == "tail" and unroll != 0 and vcols_early) else "0"))
        src.append("        hb_ticket = 0u;")
        src.append("        for (int i = 0; i < 4; ++i) { hb_reads[i] = 0u; hb_ring[i] = 0ull; }")
        src.append("        hb_ring[0] = (1ull << 32) | (t0 < nt ? (hb_u32)t0 : HB_SCHED_DONE);")
        src.append("        hb_ring[1] = (2ull << 32) | (t1 < nt ? (hb_u32)t1 : HB_SCHED_DONE);")
        src.append("    }")
    if smem or window is not None:
        # `smem` bytes of statically privatised histograms (and the window's row table), then p.win[0] words of window
        bound = "%d + p.win[0]" % (smem // 4) if window is not None else "%d" % (smem // 4)
        src.append("    for (int i = threadIdx.x; i < %s; i += HB_THREADS) reinterpret_cast<hb_u32*>(hb_smem)[i] = 0u;" % bound)
        if window is not None and window.get("box") is None:
            # per-row interval table of the hot window: [offset into the window | first in-row index << 16 | length]
            src.append("    __syncthreads();")
            tt = "hb_u32" if window["table_bits"] == 32 else "hb_u64"
            src.append("    if (p.aux[%d] != nullptr)" % window["aux"])
            src.append("        for (int i = threadIdx.x; i < %d; i += HB_THREADS)" % window["rows"])
            src.append("            reinterpret_cast<%s*>(hb_smem + %d)[i] = reinterpret_cast<const %s*>(p.aux[%d])[i];" % (tt, window["table_off"], tt, window["aux"]))
        src.append("    __syncthreads();")
    elif dynamic:
        src.append("    __syncthreads();")
    for slot, (name, dtype, is_scalar) in enumerate(prog.cols):
        if is_scalar:
            src.append("    const %s s%d = hb_scalar<%s>(p.scal[%d]);" % (_loadtype(dtype), slot, _loadtype(dtype), slot))
    src.append("    const hb_i64 ntiles = (p.n + HB_TILE - 1) / HB_TILE;")
    vcols = [(slot, dtype) for slot, (name, dtype, is_scalar) in enumerate(prog.cols) if not is_scalar]

    def loads(suffix, first, indent):
        return ["%shb_load4<%s>(p.cols[%d], %s, threadIdx.x, HB_THREADS, v%d%s);" % (indent, _loadtype(dtype), slot, first, slot, suffix)
                for slot, dtype in vcols]

    def events(suffix, indent):
        if unroll == 1:
            # big books: one copy of the event code instead of four (c4: 600 KB of SASS thrash the instruction
            # cache); the four loaded values are picked with selects so that they stay in registers
            out = ["#pragma unroll 1", "%sfor (int k = 0; k < 4; ++k) {" % indent]
            args = []
            for slot, (name, dtype, is_scalar) in enumerate(prog.cols):
                if is_scalar:
                    args.append(_colarg(slot, dtype, is_scalar, None))
                else:
                    v = "v%d%s" % (slot, suffix)
                    out.append("%s    const %s e%d = k == 0 ? %s[0] : (k == 1 ? %s[1] : (k == 2 ? %s[2] : %s[3]));" % (indent, _loadtype(dtype), slot, v, v, v, v))
                    args.append(_colarg(slot, dtype, is_scalar, "e%d" % slot))
            out.append("%s    hb_event(p, hb_smem, true%s);" % (indent, "".join(", " + a for a in args)))
            out.append("%s}" % indent)
            return out
        out = ["#pragma unroll", "%sfor (int k = 0; k < 4; ++k) {" % indent]
        out.append("%s    hb_event(p, hb_smem, true%s);" % (indent, "".join(
            ", " + _colarg(slot, dtype, is_scalar, "v%d%s[k]" % (slot, suffix)) for slot, (name, dtype, is_scalar) in enumerate(prog.cols))))
        out.append("%s}" % indent)
        return out

    def tail(indent):
        out = ["%s// tail / unaligned input: scalar loads; every lane still calls hb_event (warp-cooperative accumulates)" % indent,
               "%sfor (int k = 0; k < 4; ++k) {" % indent,
               "%s    const hb_i64 i = first + (hb_i64)k * HB_THREADS + threadIdx.x;" % indent,
               "%s    const bool live = i < p.n;" % indent,
               "%s    const hb_i64 j = live ? i : 0;" % indent,
               "%s    hb_event(p, hb_smem, live%s);" % (indent, "".join(
                   ", " + _colarg(slot, dtype, is_scalar, "hb_load1<%s>(p.cols[%d], j)" % (_loadtype(dtype), slot))
                   for slot, (name, dtype, is_scalar) in enumerate(prog.cols))),
               "%s}" % indent]
        return out

    if dynamic:
        # Dynamic tiles (option dynamic = "tail" | True; off by default).  Blocks run at different speeds (DRAM / L2
        # distance of their SM, contention in the L2 RED path: with the static grid-stride split the slowest block of c2
        # ends 2 %, of c3 3 %, of c5 20 % after the median -- tools/block_times.py).  "tail": the first 15/16 of the tiles
        # are split statically and the rest is handed out from one global counter (p.sched[0]); True: everything is.
        # Measured (profiles/r01_sweeps.txt): every block then ends within 1 % of the others, but the kernels are not
        # bound block by block -- c2 +1 %, c3 and c4 +-0, c5 -1 % (more registers), c1 at 1e7 events -27 % (two more
        # round trips in a 20 us kernel) -- so the static split stays the default.  No block barrier: the warps of a block draw tickets
        # from a shared counter; ticket k is warp-tile k % W of the block's chunk k / W, and the chunk's tile number is
        # published in a 4-slot ring by the warp that drew the first ticket of the chunk TWO earlier -- its global
        # atomicAdd returns while its own loads are in flight.  That warp first waits for the previous chunk to be
        # published, so the tile numbers of a block increase from chunk to chunk and "no tile left" (HB_SCHED_DONE) is
        # final: a warp leaves at its first such ticket, and all W tickets of every real chunk have been drawn by then.
        # A slot is only overwritten once all W readers of its previous chunk are through (hb_reads), so warps may
        # drift apart freely.  Every warp draws its next ticket while the loads of the current one are in flight.  The
        # last block to finish zeroes the counters for the next launch on this stream.
        def draw(indent, t, c, sb, tl):
            return [x.replace("@", indent) for x in [
                "@if (lane == 0u) {",
                "@    %s = atomicAdd(&hb_ticket, 1u);" % t,
                "@    volatile hb_u64* const slot = &hb_ring[(%s / (HB_THREADS / 32)) & 3u];" % t,
                "@    hb_u64 e = *slot;",
                "@    while ((hb_u32)(e >> 32) != %s / (HB_THREADS / 32) + 1u) { __nanosleep(40); e = *slot; }" % t,
                "@    atomicAdd(&hb_reads[(%s / (HB_THREADS / 32)) & 3u], 1u);" % t,
                "@    %s = (hb_u32)e;" % tl,
                "@}",
                "@%s = __shfl_sync(0xffffffffu, %s, 0);" % (t, t),
                "@%s = __shfl_sync(0xffffffffu, %s, 0);" % (tl, tl),
                "@%s = %s / (HB_THREADS / 32); %s = %s %% (HB_THREADS / 32);" % (c, t, sb, t)]]
        if dynamic == "tail" and unroll != 0 and vcols:
            src.append("    const hb_i64 nstatic = p.vec_ok ? (ntiles * 15 / 16) / gridDim.x * gridDim.x : 0;")
            src.append("    for (hb_i64 tile = blockIdx.x; tile < nstatic; tile += gridDim.x) {      // every one of these is a full tile")
            src.append("        const hb_i64 first = tile * HB_TILE;")
            for slot, dtype in vcols:
                src.append("        %s v%d[4];" % (_loadtype(dtype), slot))
            src.extend(loads("", "first", "        "))
            src.extend(events("", "        "))
            src.append("    }")
        else:
            src.append("    const hb_i64 nstatic = 0;")
        src.append("    const unsigned lane = threadIdx.x & 31u;")
        src.append("    unsigned ticket = 0u, chunk, sub; hb_u32 tile = 0u;")
        src.extend(draw("    ", "ticket", "chunk", "sub", "tile"))
        src.append("    while (true) {")
        src.append("        const bool fetcher = (lane == 0u) && (sub == 0u);")
        src.append("        hb_u32 fetched = 0u;")
        src.append("        if (fetcher) {")
        src.append("            volatile hb_u64* const prev = &hb_ring[(chunk + 1u) & 3u];")
        src.append("            while ((hb_u32)(*prev >> 32) != chunk + 2u) __nanosleep(40);")
        src.append("            fetched = atomicAdd(p.sched, 1u);")
        src.append("        }")
        src.append("        const hb_i64 first = (hb_i64)tile * HB_TILE + (hb_i64)sub * 128;")
        src.append("        const bool vec = tile != HB_SCHED_DONE && p.vec_ok && first + 128 <= p.n;")
        if unroll != 0:
            for slot, dtype in vcols:
                src.append("        %s v%d[4];" % (_loadtype(dtype), slot))
            src.append("        if (vec) {")
            src.extend(x.replace("threadIdx.x, HB_THREADS", "(int)lane, 32") for x in loads("", "first", "            "))
            src.append("        }")
        src.append("        if (fetcher) {")
        src.append("            const unsigned s2 = (chunk + 2u) & 3u, need = (HB_THREADS / 32) * ((chunk + 2u) >> 2);")
        src.append("            while (*reinterpret_cast<volatile hb_u32*>(&hb_reads[s2]) != need) __nanosleep(40);")
        src.append("            const hb_u64 t = (hb_u64)nstatic + 2ull * gridDim.x + fetched;")
        src.append("            *reinterpret_cast<volatile hb_u64*>(&hb_ring[s2]) = ((hb_u64)(chunk + 3u) << 32) | (t < (hb_u64)ntiles ? (hb_u32)t : HB_SCHED_DONE);")
        src.append("        }")
        src.append("        if (tile == HB_SCHED_DONE) break;")
        src.append("        unsigned nticket = 0u, nchunk, nsub; hb_u32 ntile = 0u;")
        src.extend(draw("        ", "nticket", "nchunk", "nsub", "ntile"))
        wtail = [x.replace("(hb_i64)k * HB_THREADS + threadIdx.x", "(hb_i64)k * 32 + lane") for x in tail("            ")]
        if unroll == 0:
            src.extend(x.replace("for (int k = 0; k < 4; ++k) {", "_Pragma(\"unroll 1\") for (int k = 0; k < 4; ++k) {")[4:] for x in wtail)
        else:
            src.append("        if (vec) {")
            src.extend(events("", "            "))
            src.append("        } else {")
            src.extend(wtail)
            src.append("        }")
        src.append("        chunk = nchunk; sub = nsub; tile = ntile;")
        src.append("    }")
    elif prefetch and vcols:
        # register double buffer: the next tile's loads are in flight while this tile's events are accumulated
        for slot, dtype in vcols:
            src.append("    %s v%d[4], v%dn[4];" % (_loadtype(dtype), slot, slot))
        src.append("    hb_i64 tile = blockIdx.x;")
        src.append("    bool full = p.vec_ok && tile < ntiles && tile * HB_TILE + HB_TILE <= p.n;")
        src.append("    if (full) {")
        src.extend(loads("", "tile * HB_TILE", "        "))
        src.append("    }")
        src.append("    for (; tile < ntiles; tile += gridDim.x) {")
        src.append("        const hb_i64 first = tile * HB_TILE;")
        src.append("        const hb_i64 next = tile + gridDim.x;")
        src.append("        const bool nfull = p.vec_ok && next < ntiles && next * HB_TILE + HB_TILE <= p.n;")
        src.append("        if (nfull) {")
        src.extend(loads("n", "next * HB_TILE", "            "))
        src.append("        }")
        src.append("        if (full) {")
        src.extend(events("", "            "))
        src.append("        } else {")
        src.extend(tail("            "))
        src.append("        }")
        src.append("        full = nfull;")
        src.append("#pragma unroll")
        src.append("        for (int k = 0; k < 4; ++k) { %s }" % " ".join("v%d[k] = v%dn[k];" % (slot, slot) for slot, dtype in vcols))
        src.append("    }")
    elif unroll == 0:
        # very big books: ONE copy of the event code, one 8-byte (or narrower) coalesced load per column and event --
        # the 4-event vectors alone would hold 8 registers per column through the whole event function
        src.append("    for (hb_i64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {")
        src.append("        const hb_i64 first = tile * HB_TILE;")
        src.extend(x.replace("for (int k = 0; k < 4; ++k) {", "_Pragma(\"unroll 1\") for (int k = 0; k < 4; ++k) {") for x in tail("        "))
        src.append("    }")
    else:
        src.append("    for (hb_i64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {")
        src.append("        const hb_i64 first = tile * HB_TILE;")
        src.append("        if (p.vec_ok && first + HB_TILE <= p.n) {")
        for slot, dtype in vcols:
            src.append("            %s v%d[4];" % (_loadtype(dtype), slot))
        src.extend(loads("", "first", "            "))
        src.extend(events("", "            "))
        src.append("        } else {")
        src.extend(tail("            "))
        src.append("        }")
        src.append("    }")
    if times_aux is not None:
        src.append("    __syncthreads();")
    stamp(1)
    if merge_lines:
        src.append("    __syncthreads();")
        src.extend("    " + line for line in merge_lines)
    if times_aux is not None:
        src.append("    __syncthreads();")
    stamp(2)
    if dynamic:
        src.append("    if (threadIdx.x == 0) {")
        src.append("        __threadfence();")
        src.append("        if (atomicAdd(p.sched + 1, 1u) == gridDim.x - 1u) { p.sched[0] = 0u; p.sched[1] = 0u; __threadfence(); }")
        src.append("    }")
    src.append("}")
    return "\n".join(src) + "\n"


def _emit_kernel_sorted(prog, event_lines, merge_lines, layout, options):
    """Kernel of a big book with sort groups.  Per tile of HB_TILE events, one 1024-thread block per SM:
      1. every thread evaluates HB_EPT events (one coalesced scalar load per column): histograms outside the groups are
         updated as in the other kernels; the float64 values the groups accumulate are staged in shared memory
         (E[k][event]); per group the event draws a ticket -- ONE native ATOMS.ADD on the group's per-tile bin counter,
         whose return value is the event's rank inside its bin;
      2. one warp per group turns the counters into bin start offsets (and adds them to the group's count table);
      3. every thread writes its events' numbers to start[bin] + rank: the tile's events, sorted by bin, per group;
      4. per group, every one of HB_LANES lanes owns a bin (bins b, b + HB_LANES, ...): it walks the bin's run of sorted
         events, adds the staged values in registers and then adds the sums to the block's float64 table with plain
         loads and stores -- nobody else touches that bin, and consecutive lanes hit consecutive banks.
    No float64 atomics and no fixed-point limbs: c4 went from ~100 shared atomics per event to 8 tickets + 32 counts."""
    groups, staged = layout["groups"], layout["staged"]
    threads, ept, tile, lanes = layout["threads"], layout["ept"], layout["tile"], layout["lanes"]
    G = len(groups)
    assert threads % lanes == 0
    split = int(options.get("sort_split", 2))              # lanes per bin in the summing phase (a power of two, at most 32; c4: 10.9 / 11.4 / 10.8 / 9.0 Gev/s for 1 / 2 / 4 / 8)
    assert split in (1, 2, 4, 8, 16, 32)
    nwg = min(threads // lanes, 15)                        # warp-groups that take sort groups (named barriers 1..15)
    src = []
    src.append('#include "hb_template.cuh"')
    src.append("#define HB_THREADS %d" % threads)
    src.append("#define HB_EPT %d" % ept)
    src.append("#define HB_TILE (HB_THREADS * HB_EPT)")
    src.append("#define HB_LANES %d" % lanes)
    src.append("#define HB_NSG %d" % G)
    src.append("#define HB_FXMISS (reinterpret_cast<hb_u32*>(hb_smem))")
    chunks = bool(layout.get("chunks"))
    chunk = tile // lanes
    odd = chunks and bool(options.get("sort_chunk_odd", True))
    if odd:
        # where sorted position `pos` lives in a group's array of event numbers: the chunks of consecutive lanes must be an odd
        # number of 32-bit words apart, so that the 32 lanes of a warp, each reading its own chunk, hit 32 different banks
        # (chunks of 48: 16-way conflicts, a quarter of the kernel's shared-memory wavefronts).  A chunk LENGTH of an odd number
        # of words (50 event numbers for tiles of 3072: 62 of the 64 lanes work) needs no padding, hence no division by the
        # chunk length in the scatter of every event and group.
        chunk = -(-tile // lanes)
        while chunk % 4 != 2:
            chunk += 1
        src.append("#define HB_SPOS(pos) (pos)")
    elif chunks:
        # (the padded form: chunks of tile / lanes event numbers, two halfwords of padding behind each)
        assert chunk * lanes == tile
        src.append("#define HB_SPOS(pos) ((pos) + (((pos) / %du) << 1))" % chunk)
    else:
        src.append("#define HB_SPOS(pos) (pos)")
    src.extend(prog.consts)
    src.append("")
    # full tiles -- all but the last one of a fill -- run a copy of the event function in which ``live`` is a compile-time
    # true: the dump-word selects of the count atomics, tickets and scatter keys fold away (option sort_full)
    full = bool(options.get("sort_full", True))
    # (filtered weights that travel as flag bits beside the sorted event numbers: stage_values; the chunked sums only)
    flags = list(layout.get("flags") or []) if layout.get("chunks") else []
    ebits = int(layout.get("ebits", 16))
    sig = ["const HbParams& p", "unsigned char* hb_smem", "const bool live_", "const int eid", "hb_u32 (&kr)[HB_NSG]"]
    if flags:
        sig.append("hb_u32& fl")
    for slot, (name, dtype, is_scalar) in enumerate(prog.cols):
        sig.append("const %s c%d /* %s */" % (ctype(dtype), slot, name))
    src.append("template <bool HB_FULL>\n__device__ __forceinline__ void hb_event(%s)\n{" % ", ".join(sig))
    src.append("    const bool live = HB_FULL || live_;")
    src.extend("    " + line for line in event_lines)
    src.append("    // ---- values the sort groups accumulate, staged for this tile")
    # staged values travel in pairs: slots 2h and 2h + 1 are one 16-byte record per event (one STS.128 here, one LDS.128 per
    # gather in the summing phase: a random 16-byte gather costs ~9 wavefronts per warp, two 8-byte ones ~12); an odd last slot
    # is an array of doubles
    def e_pair_off(h):
        return layout["e_off"] + 16 * h * tile

    def e_single_off():
        return layout["e_off"] + 16 * (len(staged) // 2) * tile
    for h in range(len(staged) // 2):
        src.append("    reinterpret_cast<double2*>(hb_smem + %d)[eid] = make_double2(%s, %s);" % (e_pair_off(h), staged[2 * h], staged[2 * h + 1]))
    if len(staged) % 2:
        src.append("    reinterpret_cast<double*>(hb_smem + %d)[eid] = %s;" % (e_single_off(), staged[-1]))

    def gather(ind, used, e, suffix=""):
        """C lines loading the staged slots ``used`` of event ``e`` into v<k><suffix>"""
        out = []
        for h in range(len(staged) // 2):
            if 2 * h in used or 2 * h + 1 in used:
                out.append("%sconst double2 vp%d%s = reinterpret_cast<const double2*>(hb_smem + %d)[%s];" % (ind, h, suffix, e_pair_off(h), e))
                for k, part in ((2 * h, "x"), (2 * h + 1, "y")):
                    if k in used:
                        out.append("%sconst double v%d%s = vp%d%s.%s;" % (ind, k, suffix, h, suffix, part))
        if len(staged) % 2 and len(staged) - 1 in used:
            out.append("%sconst double v%d%s = reinterpret_cast<const double*>(hb_smem + %d)[%s];" % (ind, len(staged) - 1, suffix, e_single_off(), e))
        return out
    if flags:
        # (filtered weights: one flag bit per event and condition, written beside the event number in the scatter)
        src.append("    fl = %s;" % " | ".join("((%s) ? %du : 0u)" % (c, 1 << f) for f, c in enumerate(flags)))
    for g in groups:
        src.append("    // ---- sort group %d: %d bins, %d values, histograms %s" % (g.gid, g.nbins, len(g.values), [hp.hid for hp in g.members]))
        src.append("    {")
        if options.get("sort_branchfree", True):
            src.append("        const bool okr = %s;" % " && ".join(["live"] + g.conds))
            src.append("        const int bin = okr ? (%s) : 0;" % g.idx)
            src.append("        const hb_u32 rank = atomicAdd(reinterpret_cast<hb_u32*>(hb_smem + (okr ? %d + 4 * bin : %d)), 1u);" % (g.off["tick"], layout["dump_off"]))
            src.append("        kr[%d] = okr ? (((hb_u32)bin << 16) | rank) : 0xffffffffu;" % g.gid)
        else:
            src.append("        hb_u32 t = 0xffffffffu;")
            src.append("        if (%s) {" % " && ".join(["live"] + g.conds))
            src.append("            const int bin = %s;" % g.idx)
            src.append("            t = ((hb_u32)bin << 16) | atomicAdd(reinterpret_cast<hb_u32*>(hb_smem + %d) + bin, 1u);" % g.off["tick"])
            src.append("        }")
            src.append("        kr[%d] = t;" % g.gid)
        src.append("    }")
    src.append("}\n")

    src.append('extern "C" __global__ void __launch_bounds__(HB_THREADS, %d) hb_fill_kernel(const HbParams p)\n{' % int(options.get("sort_min_blocks", 1)))
    src.append("    extern __shared__ __align__(16) unsigned char hb_smem[];")
    src.append("    for (int i = threadIdx.x; i < %d; i += HB_THREADS) reinterpret_cast<hb_u32*>(hb_smem)[i] = 0u;" % (layout["zero_end"] // 4))
    for cname, off, entries in layout.get("split_tables", []):
        # (bin index -> digitize count of the split axes on the same value: _apply_split_tables)
        src.append("    for (int i = threadIdx.x; i < %d; i += HB_THREADS) reinterpret_cast<int*>(hb_smem + %d)[i] = %s[i];" % (4 * entries, off, cname))
    if int(options.get("sort_stagger_us", 0)) > 0:
        # two blocks share an SM: the second starts late, so that its phases fall between the first one's
        src.append("    if (blockIdx.x >= (gridDim.x + 1) / 2 || (p.flags & 2)) { for (int i = 0; i < %d; ++i) __nanosleep(1000); }" % int(options["sort_stagger_us"]))
    src.append("    __syncthreads();")
    for slot, (name, dtype, is_scalar) in enumerate(prog.cols):
        if is_scalar:
            src.append("    const %s s%d = hb_scalar<%s>(p.scal[%d]);" % (_loadtype(dtype), slot, _loadtype(dtype), slot))
    src.append("    const hb_i64 ntiles = (p.n + HB_TILE - 1) / HB_TILE;")
    src.append("    const int hb_wg = threadIdx.x / HB_LANES, hb_l = threadIdx.x % HB_LANES;")
    src.append("    for (hb_i64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {")
    src.append("        const hb_i64 first = tile * HB_TILE;")
    for k in range(ept):
        src.append("        hb_u32 kr%d[HB_NSG];" % k)
        if flags:
            src.append("        hb_u32 fl%d;" % k)
    src.append("        // 1. evaluate, stage, draw tickets (the columns of all HB_EPT events are loaded first: their latencies overlap)")
    for k in range(ept):
        src.append("        const hb_i64 i%d = first + %d * HB_THREADS + threadIdx.x;" % (k, k))
        src.append("        const bool live%d = i%d < p.n;" % (k, k))
        src.append("        const hb_i64 j%d = live%d ? i%d : 0;" % (k, k, k))
        for slot, (name, dtype, is_scalar) in enumerate(prog.cols):
            if not is_scalar:
                src.append("        const %s ld%d_%d = hb_load1<%s>(p.cols[%d], j%d);" % (_loadtype(dtype), k, slot, _loadtype(dtype), slot, k))
    def event_calls(ind, tmpl):
        for k in range(ept):
            src.append("%shb_event<%s>(p, hb_smem, live%d, %d * HB_THREADS + (int)threadIdx.x, kr%d%s%s);" % (ind, tmpl, k, k, k, (", fl%d" % k) if flags else "", "".join(
                ", " + _colarg(slot, dtype, is_scalar, "ld%d_%d" % (k, slot))
                for slot, (name, dtype, is_scalar) in enumerate(prog.cols))))
    if full:
        src.append("        if (first + HB_TILE <= p.n) {")
        event_calls("            ", "true")
        src.append("        } else {")
        event_calls("            ", "false")
        src.append("        }")
    else:
        event_calls("        ", "false")
    src.append("        __syncthreads();")
    src.append("        // 2. bin counters -> bin start offsets (one warp per group)")
    for g in groups:
        cnt = "reinterpret_cast<hb_u32*>(hb_smem + %d)" % g.off["cnt"] if g.count_dests else "nullptr"
        src.append("        if ((threadIdx.x >> 5) == %d) hb_sort_scan(reinterpret_cast<hb_u32*>(hb_smem + %d), %d, %s);" % (
            g.gid % (threads // 32), g.off["tick"], g.nbins, cnt))
    src.append("        __syncthreads();")
    src.append("        // 3. scatter the event numbers to start[bin] + rank")
    for k in range(ept):
        # what the scatter writes: the event's number in the tile, its flag bits above it
        if flags:
            src.append("        const unsigned short en%d = (unsigned short)((%d * HB_THREADS + threadIdx.x) | (fl%d << %d));" % (k, k, k, ebits))
        else:
            src.append("        const unsigned short en%d = (unsigned short)(%d * HB_THREADS + threadIdx.x);" % (k, k))

    def scatter(ind, full_tile):
        for k in range(ept):
            for g in groups:
                if full_tile and all(c in prog.always_ok for c in g.conds):
                    # (a full tile, and no axis of this group can mask a row: every ticket is a real one)
                    src.append("%s{ const hb_u32 sp = reinterpret_cast<const hb_u32*>(hb_smem + %d)[kr%d[%d] >> 16] + (kr%d[%d] & 0xffffu); reinterpret_cast<unsigned short*>(hb_smem + %d)[HB_SPOS(sp)] = en%d; }" % (
                        ind, g.off["tick"], k, g.gid, k, g.gid, g.off["sorted"], k))
                elif options.get("sort_branchfree", True):
                    # (an event that is masked for this group writes its number to a dump slot behind the array)
                    src.append("%s{ const bool v = kr%d[%d] != 0xffffffffu; const hb_u32 sp = reinterpret_cast<const hb_u32*>(hb_smem + %d)[v ? (kr%d[%d] >> 16) : 0u] + (kr%d[%d] & 0xffffu); reinterpret_cast<unsigned short*>(hb_smem + %d)[v ? HB_SPOS(sp) : %du] = en%d; }" % (
                        ind, k, g.gid, g.off["tick"], k, g.gid, k, g.gid, g.off["sorted"], tile + 2 * lanes + 4, k))
                else:
                    src.append("%sif (kr%d[%d] != 0xffffffffu) { const hb_u32 sp = reinterpret_cast<const hb_u32*>(hb_smem + %d)[kr%d[%d] >> 16] + (kr%d[%d] & 0xffffu); reinterpret_cast<unsigned short*>(hb_smem + %d)[HB_SPOS(sp)] = en%d; }" % (
                        ind, k, g.gid, g.off["tick"], k, g.gid, k, g.gid, g.off["sorted"], k))
    if full and any(all(c in prog.always_ok for c in g.conds) for g in groups):
        src.append("        if (first + HB_TILE <= p.n) {")
        scatter("            ", True)
        src.append("        } else {")
        scatter("            ", False)
        src.append("        }")
    else:
        scatter("        ", False)
    src.append("        __syncthreads();")
    if options.get("sort_prefetch", True):
        # the columns of this block's next tile: one L2 prefetch per 32-byte sector, sent before the summing phase (the loads at
        # the top of a tile wait for DRAM otherwise: 6.7 % of the kernel's stall samples, profiles/r02_c4_ncu.txt)
        src.append("        { const hb_i64 nfirst = (tile + gridDim.x) * HB_TILE;")
        src.append("          if (nfirst + HB_TILE <= p.n) {")
        for slot, (name, dtype, is_scalar) in enumerate(prog.cols):
            if is_scalar:
                continue
            per = max(1, 32 // dtype.itemsize)                  # threads per sector
            for k in range(ept):
                src.append("              if ((threadIdx.x %% %du) == 0u) hb_prefetch_l2(reinterpret_cast<const char*>(p.cols[%d]) + (nfirst + %d * HB_THREADS + threadIdx.x) * %d);" % (
                    per, slot, k, dtype.itemsize))
        src.append("          } }")
    if chunks:
        # the balanced variant: every lane walks a contiguous chunk of the sorted events; a run that ends inside the chunk is
        # flushed on the spot (it ends in no other lane's chunk), a run that goes on in the next lane is a "tail", added
        # after the group's barrier by the first lane that holds a tail of that bin
        src.append("        // 4. chunked sums over the sorted events")
        for g in groups:
            nv = len(g.values)
            wg = g.gid % nwg
            used_slots = sorted(set(st_[1] for st_ in g.stage))
            src.append("        if (hb_wg == %d) {   // sort group %d" % (wg, g.gid))
            src.append("            const hb_u32* const st = reinterpret_cast<const hb_u32*>(hb_smem + %d);" % g.off["tick"])
            src.append("            const unsigned short* const so = reinterpret_cast<const unsigned short*>(hb_smem + %d);" % g.off["sorted"])
            src.append("            double* const acc = reinterpret_cast<double*>(hb_smem + %d);" % g.off["acc"])
            src.append("            int* const tkeys = reinterpret_cast<int*>(hb_smem + %d);" % g.off["tkey"])
            src.append("            double* const tvals = reinterpret_cast<double*>(hb_smem + %d);" % g.off["tval"])
            src.append("            const hb_u32 ng = st[%d];" % g.nbins)
            src.append("            const hb_u32 p0 = (hb_u32)hb_l * %du, p1 = min(p0 + %du, ng);" % (chunk, chunk))
            # every position of this lane's chunk has pos / chunk == hb_l: HB_SPOS(pos) == pos + 2 * hb_l, no division per event
            src.append("            const unsigned short* const sop = so + %s;" % ("0" if odd else "2 * hb_l"))
            src.append("            " + " ".join("double a%d = 0.0;" % j for j in range(nv)))
            src.append("            int tkey = -1;")
            nvp = (nv + 1) // 2 * 2

            def rowop(ptr, fmt_pair, fmt_single):
                """statements over the 16-byte pairs of one [values] row at ``ptr`` (a double*)"""
                out = ["double2* const r = reinterpret_cast<double2*>(%s); double2 q;" % ptr]
                for h in range(nvp // 2):
                    out.append((fmt_pair if 2 * h + 1 < nv else fmt_single).format(h=h, a=2 * h, b=2 * h + 1))
                return " ".join(out)
            # (rows of the block's table are [bin][values]: a flush is 3 LDS.128 + 3 STS.128 for c4's six sums instead of 6 + 6
            #  64-bit ones -- the flush path is walked by the whole warp whenever one lane's run ends, which is most iterations)
            flush = "{ " + rowop("acc + key * %d" % nvp, "q = r[{h}]; q.x += a{a}; q.y += a{b}; r[{h}] = q;", "q = r[{h}]; q.x += a{a}; r[{h}] = q;") + " " + " ".join("a%d = 0.0;" % j for j in range(nv)) + " }"
            # (sort_skip_sum: a measurement aid -- the sums are skipped, the float64 columns stay empty)
            src.append("            if (p0 < ng%s) {" % (" && p.n < 0" if options.get("sort_skip_sum") else ""))
            src.append("                int key = hb_sort_find(st, %d, p0);" % g.nbins)
            src.append("                hb_u32 runend = st[key + 1];")
            advance = ["if (pos%s >= runend) {       // the run of `key` ends in this lane's chunk, and in no other lane's",
                       "    " + flush,
                       "    do { ++key; runend = st[key + 1]; } while (pos%s >= runend);",
                       "}"]

            def accumulate(ind, sfx):
                """C lines adding the gathered values v<k><sfx> of one event (flag bits in er<sfx>) to the running sums"""
                squared = set()
                for j, st_ in enumerate(g.stage):
                    kind, k = st_[0], st_[1]
                    if kind in ("sq", "gsq") and k not in squared:
                        squared.add(k)
                        src.append("%sconst double q%d%s = hb_mul<double>(v%d%s, v%d%s);" % (ind, k, sfx, k, sfx, k, sfx))
                    if kind == "E":
                        src.append("%sa%d += v%d%s;" % (ind, j, k, sfx))
                    elif kind == "sq":
                        src.append("%sa%d += q%d%s;" % (ind, j, k, sfx))
                    elif kind == "gate":
                        src.append("%sa%d += (er%s & %du) ? v%d%s : 0.0;" % (ind, j, sfx, 1 << (ebits + st_[2]), k, sfx))
                    else:
                        assert kind == "gsq"
                        src.append("%sa%d += (er%s & %du) ? q%d%s : 0.0;" % (ind, j, sfx, 1 << (ebits + st_[2]), k, sfx))
            emask = (1 << ebits) - 1 if flags else 0xffff
            pair = odd and chunk % 2 == 0 and bool(options.get("sort_pair", False))
            if pair:
                # two events per round: ONE 32-bit load brings both event numbers (the chunk starts on an even position), both
                # gathers are in flight before the first run-end test -- the loop is a chain of dependent shared-memory loads
                # (short-scoreboard stalls lead the kernel's stall reasons, profiles/r02_c4_ncu.txt)
                src.append("                hb_u32 pos = p0;")
                src.append("                for (; pos + 1u < p1; pos += 2u) {")
                src.append("                    const unsigned ee = *reinterpret_cast<const unsigned*>(sop + pos);")
                src.append("                    const unsigned er_a = ee & 0xffffu, er_b = ee >> 16, e_a = er_a & %du, e_b = er_b & %du;" % (emask, emask))
                src.extend(gather("                    ", used_slots, "e_a", "_a"))
                src.extend(gather("                    ", used_slots, "e_b", "_b"))
                src.extend("                    " + x % "" if "%s" in x else "                    " + x for x in advance)
                accumulate("                    ", "_a")
                src.extend("                    " + x % " + 1u" if "%s" in x else "                    " + x for x in advance)
                accumulate("                    ", "_b")
                src.append("                }")
                src.append("                if (pos < p1) {")
            else:
                src.append("                for (hb_u32 pos = p0; pos < p1; ++pos) {")
            src.extend("                    " + x % "" if "%s" in x else "                    " + x for x in advance)
            src.append("                    const unsigned er = sop[pos], e = er & %du;" % emask)
            src.extend(gather("                    ", used_slots, "e"))
            accumulate("                    ", "")
            src.append("                }")
            src.append("                if (runend > p1) tkey = key;       // the run goes on in the next lane: a tail")
            src.append("                else { %s }" % flush)
            src.append("            }")
            if g.off["tkey"] == g.off["sorted"]:
                # (the tails re-use the sorted event numbers: every lane of the group must be done reading them)
                src.append('            asm volatile("bar.sync %%0, %%1;" :: "r"(%d), "r"(HB_LANES) : "memory");' % (1 + wg))
            src.append("            tkeys[hb_l] = tkey;")
            src.append("            if (tkey >= 0) { %s }" % rowop("tvals + hb_l * %d" % nvp, "r[{h}] = make_double2(a{a}, a{b});", "r[{h}] = make_double2(a{a}, 0.0);"))
            src.append('            asm volatile("bar.sync %%0, %%1;" :: "r"(%d), "r"(HB_LANES) : "memory");' % (1 + wg))
            src.append("            for (int b = hb_l; b <= %d; b += HB_LANES) reinterpret_cast<hb_u32*>(hb_smem + %d)[b] = 0u;" % (g.nbins, g.off["tick"]))
            src.append("            if (tkey >= 0 && (hb_l == 0 || tkeys[hb_l - 1] != tkey)) {")
            src.append("                const int key = tkey;")
            src.append("                for (int m = hb_l + 1; m < HB_LANES && tkeys[m] == key; ++m) { %s }" % rowop("tvals + m * %d" % nvp, "q = r[{h}]; a{a} += q.x; a{b} += q.y;", "q = r[{h}]; a{a} += q.x;"))
            src.append("                " + flush)
            src.append("            }")
            src.append("        }")
    src.append("        // 4. every lane of a group's warps sums the runs of its bins")
    for g in ([] if chunks else groups):
        nv = len(g.values)
        wg = g.gid % nwg
        used_slots = sorted(set(k for kind, k in g.stage))

        def element(ind, e):
            return gather(ind, used_slots, e, "_" + e)

        def adds(ind, e):
            out = []
            for j, (kind, k) in enumerate(g.stage):
                if kind == "E":
                    out.append("%sa%d += v%d_%s;" % (ind, j, k, e))
                else:
                    out.append("%sa%d += hb_mul<double>(v%d_%s, v%d_%s);" % (ind, j, k, e, k, e))
            return out
        src.append("        if (hb_wg == %d) {   // sort group %d" % (wg, g.gid))
        src.append("            const hb_u32* const st = reinterpret_cast<const hb_u32*>(hb_smem + %d);" % g.off["tick"])
        src.append("            const unsigned short* const so = reinterpret_cast<const unsigned short*>(hb_smem + %d);" % g.off["sorted"])
        src.append("            double* const acc = reinterpret_cast<double*>(hb_smem + %d);" % g.off["acc"])
        # HB_SPLIT lanes share a bin: lane q of them takes the events q, q + HB_SPLIT, ... of the bin's run, a butterfly of
        # shuffles adds the partial sums, lane 0 adds them to the table.  One lane per bin left the warp that holds the
        # fullest bin working alone while the block waited at the barrier (x*y of c4: 11 % of a tile in one bin, the
        # barrier the top stall reason at 5 warps per issue -- profiles/r02_c4_ncu.txt).
        src.append("            for (int w0 = 0; w0 < %d; w0 += HB_LANES) {" % (g.nbins * split))
        src.append("                const int w = w0 + hb_l;")
        src.append("                const bool mine = w < %d;" % (g.nbins * split))
        src.append("                const int b = mine ? (w / %d) : 0;" % split)
        src.append("                const hb_u32 end = mine ? st[b + 1] : 0u;")
        src.append("                hb_u32 pos = mine ? st[b] + (hb_u32)(w %% %d) : 0u;" % split)
        src.append("                const bool any = mine && st[b] != end;")
        src.append("                " + " ".join("double a%d = 0.0;" % j for j in range(nv)))
        src.append("                for (; pos + %du < end; pos += %du) {       // two events per round: their loads are in flight together" % (split, 2 * split))
        src.append("                    const unsigned e0 = so[pos], e1 = so[pos + %du];" % split)
        src.extend(element("                    ", "e0"))
        src.extend(element("                    ", "e1"))
        src.extend(adds("                    ", "e0"))
        src.extend(adds("                    ", "e1"))
        src.append("                }")
        src.append("                if (pos < end) {")
        src.append("                    const unsigned e0 = so[pos];")
        src.extend(element("                    ", "e0"))
        src.extend(adds("                    ", "e0"))
        src.append("                }")
        d = 1
        while d < split:
            src.append("                " + " ".join("a%d += __shfl_xor_sync(0xffffffffu, a%d, %d);" % (j, j, d) for j in range(nv)))
            d *= 2
        src.append("                if (any && (w %% %d) == 0) { %s }      // this lane alone adds to bin b" % (split, " ".join("acc[%d + b] += a%d;" % (j * g.nbins, j) for j in range(nv))))
        src.append("            }")
        src.append('            asm volatile("bar.sync %%0, %%1;" :: "r"(%d), "r"(HB_LANES) : "memory");' % (1 + wg))
        src.append("            for (int b = hb_l; b <= %d; b += HB_LANES) reinterpret_cast<hb_u32*>(hb_smem + %d)[b] = 0u;" % (g.nbins, g.off["tick"]))
        src.append("        }")
    src.append("        __syncthreads();")
    src.append("    }")
    src.append("    __syncthreads();")
    src.extend("    " + line for line in merge_lines)
    for g in groups:
        src.append("    // merge sort group %d" % g.gid)
        src.append("    for (int b = threadIdx.x; b < %d; b += HB_THREADS) {" % g.nbins)
        for j, dests in enumerate(g.dests):
            if chunks:
                src.append("        { const double v = reinterpret_cast<const double*>(hb_smem + %d)[b * %d + %d];" % (g.off["acc"], (len(g.values) + 1) // 2 * 2, j))
            else:
                src.append("        { const double v = reinterpret_cast<const double*>(hb_smem + %d)[b];" % (g.off["acc"] + 8 * j * g.nbins))
            src.append("          if (v != 0.0) { %s } }" % " ".join("hb_red_f64(p.hist[%d] + (hb_i64)b * %d + %d, v);" % (hid, ncol, col) for hid, col, ncol in dests))
        if g.count_dests:
            src.append("        { const hb_u32 c = reinterpret_cast<const hb_u32*>(hb_smem + %d)[b];" % g.off["cnt"])
            src.append("          if (c != 0u) { const double v = (double)c; %s } }" % " ".join(
                "hb_red_f64(p.hist[%d] + (hb_i64)b * %d + %d, v);" % (hid, ncol, col) for hid, col, ncol in g.count_dests))
        src.append("    }")
    src.append("}")
    return "\n".join(src) + "\n"


def _emit_kernel_sorted_pipe(prog, event_lines, merge_lines, layout, options):
    """The sorted strategy as a two-stage pipeline inside one 1024-thread block (``sort_pipe``).

    In ``_emit_kernel_sorted`` the phases of a tile follow one another behind block barriers: the event phase is bound by the
    issue slots (56 % of the instructions, a quarter of the shared-memory wavefronts), the summing phase by the shared-memory
    pipe (a third of the instructions, 63 % of the wavefronts), and half of the warps idle through it -- 44 % of the time
    (profiles/r02_c4_ncu.txt).  Here the block's warps are split by role: the first HB_PROD threads (*producers*) evaluate
    tile k + 1 -- stage, tickets, scan, scatter, exactly as before -- while one *consumer* warp per sort group sums tile k.
    Tick counters, sorted event numbers and staged values exist twice; named barriers hand the buffers over
    (producers ``bar.arrive`` FULL_b, consumers ``bar.sync`` it, and the other way round for EMPTY_b).  The consumer of a
    group is ONE warp: every lane walks a chunk of HB_TILE / 32 sorted events, runs that end inside the chunk are flushed
    on the spot, tails are combined by a segmented reduction over shuffles -- no tail arrays, no barrier inside the group.
    The count tables belong to the producers (scan), the float64 tables to the consumers: no table has two writers."""
    groups, staged = layout["groups"], layout["staged"]
    threads, ept, tile = layout["threads"], layout["ept"], layout["tile"]
    pp = layout["pipe"]
    prod, tick_stride, tile_stride, tick_dump, padh = pp["prod"], pp["tick_stride"], pp["tile_stride"], pp["tick_dump"], pp["pad"]
    G = len(groups)
    chunk = tile // 32
    assert chunk * 32 == tile and prod + 32 * G == threads and prod % 32 == 0
    BAR_PROD, BAR_FULL, BAR_EMPTY = 1, 2, 4
    src = []
    src.append('#include "hb_template.cuh"')
    src.append("#define HB_THREADS %d" % threads)
    src.append("#define HB_PROD %d" % prod)
    src.append("#define HB_EPT %d" % ept)
    src.append("#define HB_TILE (HB_PROD * HB_EPT)")
    src.append("#define HB_LANES 32")
    src.append("#define HB_NSG %d" % G)
    src.append("#define HB_FXMISS (reinterpret_cast<hb_u32*>(hb_smem))")
    if padh:
        src.append("#define HB_SPOS(pos) ((pos) + ((pos) / %du) * %du)" % (chunk, padh))
    else:
        src.append("#define HB_SPOS(pos) (pos)")
    src.append("#define HB_BAR_SYNC(id, n) asm volatile(\"bar.sync %0, %1;\" :: \"r\"(id), \"r\"(n) : \"memory\")")
    src.append("#define HB_BAR_ARRIVE(id, n) asm volatile(\"bar.arrive %0, %1;\" :: \"r\"(id), \"r\"(n) : \"memory\")")
    src.extend(prog.consts)
    src.append("// per sort group: tick counters, sorted event numbers (both relative to the buffer), sums, counts (-1: none), bins")
    src.append("__constant__ int hb_goff[%d][6] = { %s };" % (
        G, ", ".join("{ %d, %d, %d, %d, %d, 0 }" % (g.off["tick"], g.off["sorted"], g.off["acc"], g.off["cnt"] if g.count_dests else -1, g.nbins) for g in groups)))
    src.append("")
    sig = ["const HbParams& p", "unsigned char* hb_smem", "unsigned char* const hb_tk", "unsigned char* const hb_tl", "const bool live", "const int eid",
           "hb_u32 (&kr)[HB_NSG]"]
    for slot, (name, dtype, is_scalar) in enumerate(prog.cols):
        sig.append("const %s c%d /* %s */" % (ctype(dtype), slot, name))
    src.append("__device__ __forceinline__ void hb_event(%s)\n{" % ", ".join(sig))
    src.extend("    " + line for line in event_lines)
    src.append("    // ---- values the sort groups accumulate, staged for this tile (hb_tl: this tile's buffer)")

    def e_pair_off(h):
        return layout["e_off"] + 16 * h * tile

    def e_single_off():
        return layout["e_off"] + 16 * (len(staged) // 2) * tile
    for h in range(len(staged) // 2):
        src.append("    reinterpret_cast<double2*>(hb_tl + %d)[eid] = make_double2(%s, %s);" % (e_pair_off(h), staged[2 * h], staged[2 * h + 1]))
    if len(staged) % 2:
        src.append("    reinterpret_cast<double*>(hb_tl + %d)[eid] = %s;" % (e_single_off(), staged[-1]))

    def gather(ind, used, e, sfx=""):
        out = []
        for h in range(len(staged) // 2):
            if 2 * h in used or 2 * h + 1 in used:
                out.append("%sconst double2 vp%d%s = reinterpret_cast<const double2*>(hb_tl + %d)[%s];" % (ind, h, sfx, e_pair_off(h), e))
                for k, part in ((2 * h, "x"), (2 * h + 1, "y")):
                    if k in used:
                        out.append("%sconst double v%d%s = vp%d%s.%s;" % (ind, k, sfx, h, sfx, part))
        if len(staged) % 2 and len(staged) - 1 in used:
            out.append("%sconst double v%d%s = reinterpret_cast<const double*>(hb_tl + %d)[%s];" % (ind, len(staged) - 1, sfx, e_single_off(), e))
        return out
    for g in groups:
        src.append("    // ---- sort group %d: %d bins, %d values, histograms %s" % (g.gid, g.nbins, len(g.values), [hp.hid for hp in g.members]))
        src.append("    {")
        src.append("        const bool okr = %s;" % " && ".join(["live"] + g.conds))
        src.append("        const int bin = okr ? (%s) : 0;" % g.idx)
        src.append("        const hb_u32 rank = atomicAdd(reinterpret_cast<hb_u32*>(hb_tk + (okr ? %d + 4 * bin : %d)), 1u);" % (g.off["tick"], tick_dump))
        src.append("        kr[%d] = okr ? (((hb_u32)bin << 16) | rank) : 0xffffffffu;" % g.gid)
        src.append("    }")
    src.append("}\n")

    src.append('extern "C" __global__ void __launch_bounds__(HB_THREADS, 1) hb_fill_kernel(const HbParams p)\n{')
    src.append("    extern __shared__ __align__(16) unsigned char hb_smem[];")
    src.append("    for (int i = threadIdx.x; i < %d; i += HB_THREADS) reinterpret_cast<hb_u32*>(hb_smem)[i] = 0u;" % (layout["zero_end"] // 4))
    src.append("    __syncthreads();")
    for slot, (name, dtype, is_scalar) in enumerate(prog.cols):
        if is_scalar:
            src.append("    const %s s%d = hb_scalar<%s>(p.scal[%d]);" % (_loadtype(dtype), slot, _loadtype(dtype), slot))
    src.append("    const hb_i64 ntiles = (p.n + HB_TILE - 1) / HB_TILE;")
    src.append("    // tiles blockIdx.x, blockIdx.x + gridDim.x, ...: hb_nk of them; the k-th one goes through buffer k & 1")
    src.append("    const hb_i64 hb_nk = ntiles > (hb_i64)blockIdx.x ? (ntiles - (hb_i64)blockIdx.x + (hb_i64)gridDim.x - 1) / (hb_i64)gridDim.x : 0;")
    src.append("    if (threadIdx.x < HB_PROD) {")
    src.append("        // ===== producers: evaluate, stage, draw tickets; scan; scatter")
    src.append("        for (hb_i64 k = 0; k < hb_nk; ++k) {")
    src.append("            const int hb_b = (int)(k & 1);")
    src.append("            unsigned char* const hb_tk = hb_smem + hb_b * %d;" % tick_stride)
    src.append("            unsigned char* const hb_tl = hb_smem + hb_b * %d;" % tile_stride)
    src.append("            const hb_i64 first = ((hb_i64)blockIdx.x + k * (hb_i64)gridDim.x) * HB_TILE;")
    for q in range(ept):
        src.append("            hb_u32 kr%d[HB_NSG];" % q)
    for q in range(ept):
        src.append("            const hb_i64 i%d = first + %d * HB_PROD + threadIdx.x;" % (q, q))
        src.append("            const bool live%d = i%d < p.n;" % (q, q))
        src.append("            const hb_i64 j%d = live%d ? i%d : 0;" % (q, q, q))
        for slot, (name, dtype, is_scalar) in enumerate(prog.cols):
            if not is_scalar:
                src.append("            const %s ld%d_%d = hb_load1<%s>(p.cols[%d], j%d);" % (_loadtype(dtype), q, slot, _loadtype(dtype), slot, q))
    src.append("            if (k >= 2) HB_BAR_SYNC(%d + hb_b, HB_THREADS);      // the consumers are done with this buffer (tile k - 2)" % BAR_EMPTY)
    for q in range(ept):
        src.append("            hb_event(p, hb_smem, hb_tk, hb_tl, live%d, %d * HB_PROD + (int)threadIdx.x, kr%d%s);" % (q, q, q, "".join(
            ", " + _colarg(slot, dtype, is_scalar, "ld%d_%d" % (q, slot))
            for slot, (name, dtype, is_scalar) in enumerate(prog.cols))))
    src.append("            HB_BAR_SYNC(%d, HB_PROD);" % BAR_PROD)
    src.append("            if ((threadIdx.x >> 5) < %du) {        // one warp per group: bin counters -> bin start offsets" % G)
    src.append("                const int sg = (int)(threadIdx.x >> 5);")
    src.append("                hb_sort_scan(reinterpret_cast<hb_u32*>(hb_tk + hb_goff[sg][0]), hb_goff[sg][4], hb_goff[sg][3] >= 0 ? reinterpret_cast<hb_u32*>(hb_smem + hb_goff[sg][3]) : nullptr);")
    src.append("            }")
    src.append("            HB_BAR_SYNC(%d, HB_PROD);" % BAR_PROD)
    dump_slot = tile + 32 * padh + 4
    for q in range(ept):
        for g in groups:
            src.append("            { const bool v = kr%d[%d] != 0xffffffffu; const hb_u32 sp = reinterpret_cast<const hb_u32*>(hb_tk + %d)[v ? (kr%d[%d] >> 16) : 0u] + (kr%d[%d] & 0xffffu); reinterpret_cast<unsigned short*>(hb_tl + %d)[v ? HB_SPOS(sp) : %du] = (unsigned short)(%d * HB_PROD + threadIdx.x); }" % (
                q, g.gid, g.off["tick"], q, g.gid, q, g.gid, g.off["sorted"], dump_slot, q))
    src.append("            __threadfence_block();")
    src.append("            HB_BAR_ARRIVE(%d + hb_b, HB_THREADS);                 // tile k is ready" % BAR_FULL)
    src.append("        }")
    src.append("    } else {")
    src.append("        // ===== consumers: warp g sums the sorted events of group g")
    src.append("        const int hb_g = ((int)threadIdx.x - HB_PROD) >> 5, hb_l = (int)(threadIdx.x & 31u);")
    src.append("        for (hb_i64 k = 0; k < hb_nk; ++k) {")
    src.append("            const int hb_b = (int)(k & 1);")
    src.append("            unsigned char* const hb_tk = hb_smem + hb_b * %d;" % tick_stride)
    src.append("            unsigned char* const hb_tl = hb_smem + hb_b * %d;" % tile_stride)
    src.append("            HB_BAR_SYNC(%d + hb_b, HB_THREADS);" % BAR_FULL)
    # groups of one shape (bins, staged slots) share ONE copy of the summing code: a consumer warp keeps its group for the whole
    # kernel, so the group's offsets are three warp-uniform words from a __constant__ table (c4: eight groups, one copy)
    classes = {}
    for g in groups:
        classes.setdefault((g.nbins, tuple(g.stage)), []).append(g)
    for (nbins, stage), members in classes.items():
        nv = len(stage)
        nvp = (nv + 1) // 2 * 2
        used_slots = sorted(set(k for kind, k in stage))

        def rowop(ptr, fmt_pair, fmt_single):
            out = ["double2* const r = reinterpret_cast<double2*>(%s); double2 q;" % ptr]
            for h in range(nvp // 2):
                out.append((fmt_pair if 2 * h + 1 < nv else fmt_single).format(h=h, a=2 * h, b=2 * h + 1))
            return " ".join(out)
        flush = "{ " + rowop("acc + key * %d" % nvp, "q = r[{h}]; q.x += a{a}; q.y += a{b}; r[{h}] = q;", "q = r[{h}]; q.x += a{a}; r[{h}] = q;") + " " + " ".join("a%d = 0.0;" % j for j in range(nv)) + " }"
        mask = sum(1 << g.gid for g in members)
        cond = "true" if len(members) == G else "((%du >> hb_g) & 1u) != 0u" % mask
        src.append("            if (%s) {   // sort groups %s: %d bins, %d values" % (cond, [g.gid for g in members], nbins, nv))
        src.append("                hb_u32* const st = reinterpret_cast<hb_u32*>(hb_tk + hb_goff[hb_g][0]);")
        src.append("                unsigned char* const sob = hb_tl + hb_goff[hb_g][1];")
        src.append("                const unsigned short* const so = reinterpret_cast<const unsigned short*>(sob);")
        src.append("                double* const acc = reinterpret_cast<double*>(hb_smem + hb_goff[hb_g][2]);")
        src.append("                const hb_u32 ng = st[%d];" % nbins)
        src.append("                const hb_u32 p0 = (hb_u32)hb_l * %du, p1 = min(p0 + %du, ng);" % (chunk, chunk))
        src.append("                " + " ".join("double a%d = 0.0;" % j for j in range(nv)))
        src.append("                int tkey = -1;")
        src.append("                if (p0 < ng%s) {" % (" && p.n < 0" if options.get("sort_skip_sum") else ""))
        src.append("                    int key = hb_sort_find(st, %d, p0);" % nbins)
        src.append("                    hb_u32 runend = st[key + 1];")
        # the event numbers of a lane's chunk are contiguous: four of them per LDS.64, their staged values gathered before the
        # first add -- the summing warp is alone on its group, so its speed is the latency of this chain (one event at a time
        # the eight consumer warps were busy 97 % of the kernel and the producers waited a third of theirs,
        # profiles/r02_c4_pipe_ncu.txt)
        batch = max(1, min(4, 24 // max(2 * len(used_slots), 1))) if chunk % 4 == 0 else 1
        src.append("                    const unsigned short* const sol = so + (hb_u32)hb_l * %du;" % (chunk + padh))
        src.append("                    hb_u32 pos = p0;")

        def one(ind, q, sfx):
            out = ["%sif (pos + %du >= runend) {       // the run of `key` ends in this lane's chunk, and in no other lane's" % (ind, q)]
            out.append("%s    %s" % (ind, flush))
            out.append("%s    do { ++key; runend = st[key + 1]; } while (pos + %du >= runend);" % (ind, q))
            out.append("%s}" % ind)
            for j, (kind, k) in enumerate(stage):
                if kind == "E":
                    out.append("%sa%d += v%d%s;" % (ind, j, k, sfx))
                else:
                    out.append("%sa%d += hb_mul<double>(v%d%s, v%d%s);" % (ind, j, k, sfx, k, sfx))
            return out
        if batch == 4:
            src.append("                    for (; pos + 4u <= p1; pos += 4u) {")
            src.append("                        const uint2 ev = *reinterpret_cast<const uint2*>(sol + (pos - p0));")
            src.append("                        const unsigned e_0 = ev.x & 0xffffu, e_1 = ev.x >> 16, e_2 = ev.y & 0xffffu, e_3 = ev.y >> 16;")
            for q in range(4):
                src.extend(gather("                        ", used_slots, "e_%d" % q, "_%d" % q))
            for q in range(4):
                src.extend(one("                        ", q, "_%d" % q))
            src.append("                    }")
        src.append("                    for (; pos < p1; ++pos) {")
        src.append("                        const unsigned e = sol[pos - p0];")
        src.extend(gather("                        ", used_slots, "e", ""))
        src.extend(one("                        ", 0, ""))
        src.append("                    }")
        src.append("                    if (runend > p1) tkey = key;       // the run goes on in the next lane: a tail")
        src.append("                    else { %s }" % flush)
        src.append("                }")
        src.append("                __syncwarp();")
        src.append("                // tails: lanes holding the same bin are neighbours; the first of them adds up the others' sums.  They travel")
        src.append("                // through the group's array of sorted event numbers, which nobody reads any more")
        src.append("                int* const tkeys = reinterpret_cast<int*>(sob);")
        src.append("                double* const tvals = reinterpret_cast<double*>(sob + 128);")
        src.append("                tkeys[hb_l] = tkey;")
        src.append("                if (tkey >= 0) { %s }" % rowop("tvals + hb_l * %d" % nvp, "r[{h}] = make_double2(a{a}, a{b});", "r[{h}] = make_double2(a{a}, 0.0);"))
        src.append("                __syncwarp();")
        src.append("                if (tkey >= 0 && (hb_l == 0 || tkeys[hb_l - 1] != tkey)) {")
        src.append("                    const int key = tkey;")
        src.append("                    for (int m = hb_l + 1; m < 32 && tkeys[m] == key; ++m) { %s }" % rowop("tvals + m * %d" % nvp, "q = r[{h}]; a{a} += q.x; a{b} += q.y;", "q = r[{h}]; a{a} += q.x;"))
        src.append("                    " + flush)
        src.append("                }")
        src.append("                for (int b = hb_l; b <= %d; b += 32) st[b] = 0u;" % nbins)
        src.append("            }")
    src.append("            if (k + 2 < hb_nk) { __threadfence_block(); HB_BAR_ARRIVE(%d + hb_b, HB_THREADS); }   // the buffer may be filled again" % BAR_EMPTY)
    src.append("        }")
    src.append("    }")
    src.append("    __syncthreads();")
    src.extend("    " + line for line in merge_lines)
    for g in groups:
        src.append("    // merge sort group %d" % g.gid)
        src.append("    for (int b = threadIdx.x; b < %d; b += HB_THREADS) {" % g.nbins)
        for j, dests in enumerate(g.dests):
            src.append("        { const double v = reinterpret_cast<const double*>(hb_smem + %d)[b * %d + %d];" % (g.off["acc"], (len(g.values) + 1) // 2 * 2, j))
            src.append("          if (v != 0.0) { %s } }" % " ".join("hb_red_f64(p.hist[%d] + (hb_i64)b * %d + %d, v);" % (hid, ncol, col) for hid, col, ncol in dests))
        if g.count_dests:
            src.append("        { const hb_u32 c = reinterpret_cast<const hb_u32*>(hb_smem + %d)[b];" % g.off["cnt"])
            src.append("          if (c != 0u) { const double v = (double)c; %s } }" % " ".join(
                "hb_red_f64(p.hist[%d] + (hb_i64)b * %d + %d, v);" % (hid, ncol, col) for hid, col, ncol in g.count_dests))
        src.append("    }")
    src.append("}")
    return "\n".join(src) + "\n"


def _goal_lookup(g, aux_slot):
    """C lines (event-function scope) that find the position of one group goal's key among the keys THIS fill holds
    (ascending 64-bit patterns, written by the host after discovery: fillable.fill_hists).  Done once per goal and
    event, however many histograms group by it."""
    n = g.c
    out = ["const hb_u64* const gq_%s = reinterpret_cast<const hb_u64*>(p.aux[%d]);" % (n, aux_slot),
           "const int gn_%s = (int)gq_%s[0];" % (n, n)]
    if g.dtype.kind == "f" and not g.extra["nanflow"]:
        # nanflow=False: NaN rows are dropped (calc/__init__.py:189-192)
        out.append("const int gd_%s = (%s != %s) ? -1 : hb_key_find(gq_%s, hb_keybits(%s));" % (n, g.c, g.c, n, g.c))
    else:
        out.append("const int gd_%s = hb_key_find(gq_%s, hb_keybits(%s));" % (n, n, g.c))
    return out


def _group_lookup(hp, aux_slot, leader=None, share=False):
    """C lines that turn the group keys of one event into the slab slot of hist ``hp`` (or reject the row): the
    positions of its goals' keys (``_goal_lookup``) index the histogram's dense slot map of this fill.
    Histograms that group by the same goals form a *signature*; its first histogram (``leader`` None) looks the slot up, the
    others take that slot when the host passes a null map -- their maps are the same bytes (fillable._bind_stores), the usual
    case for the histograms of a book -- and look it up in their own map otherwise: c4 does one map load per event, not eight."""
    sig = "_".join(g.c for g in hp.groups)
    comb = "gd_%s" % hp.groups[0].c
    for g in hp.groups[1:]:
        comb = "(%s) * gn_%s + gd_%s" % (comb, g.c, g.c)
    if leader is not None and share:
        # kernel variant for fills in which every follower's map is its leader's (option group_fast, chosen per fill by
        # fillable.fill_hists): no null test, no second map load -- ~10 instructions per histogram and event on c4
        return ["const bool okg = okg_%s;" % sig, "const int slot = slot_%s;" % sig]
    out = ["const int* const gm = reinterpret_cast<const int*>(p.aux[%d]);" % aux_slot]
    if leader is None:
        out.append("const bool okg = %s;" % " && ".join("gd_%s >= 0" % g.c for g in hp.groups))
        out.append("const int slot = okg ? gm[%s] : -1;" % comb)
        out.append("okg_%s = okg; slot_%s = slot;" % (sig, sig))
    else:
        out.append("const bool okg = okg_%s;" % sig)
        out.append("const int slot = gm == nullptr ? slot_%s : (okg ? gm[%s] : -1);" % (sig, comb))
    return out


def _slot_ptr(hp, region, j, bidx, nrows, const=False):
    """(C type, C expression) of the pointer to bin ``bidx`` of float64 slot ``j`` (j = number of slots: the u32 count
    block) of a histogram whose shared-memory region starts at byte ``region`` and has ``nrows`` rows (int or C expr)."""
    units = hp.slot_units()
    before = 8 * sum(units[:j])
    fx = j < len(units) and hp.f64_terms[j] in hp.fx_set
    ctype_ = "double" if (j < len(units) and not fx) else "hb_u32"
    if isinstance(nrows, int):
        base = "hb_smem + %d" % (region + before * nrows)
    else:
        base = "hb_smem + %d + %d * %s" % (region, before, nrows)
    c = "const " if const else ""
    return ctype_, "(reinterpret_cast<%s%s*>(%s) + %s)" % (c, ctype_, base, bidx)


def _emit_row_adds(hp, options, region, bidx, nrows, indent, grow=None, gexact=None):
    """C lines adding one event's terms to the shared-memory row ``bidx`` of a histogram (region start ``region``,
    ``nrows`` rows).  ``grow`` = the bin's row in the global histogram, ``gexact`` = its exact accumulators
    (deterministic mode): where a term goes that does not fit its fixed-point window."""
    slots, count_u32 = hp.row_layout(options)
    det = bool(options.get("deterministic", False))
    out = []

    def val(t):
        return "1.0" if hp.terms[t] is None else hp.terms[t]
    fxs = [(j, t) for j, t in enumerate(slots) if t in hp.fx_set]
    # fixed point: stage k of every term before stage k+1 of any, so that the returning atomics overlap
    # (the value fxv<i> and its decoded form fxd<i> were computed once at the top of the event function: _emit_fx_decodes)
    for j, t in fxs:
        out.append("%sconst double fv%d = fxv%d;" % (indent, j, hp.fx_index[t]))
        out.append("%sconst HbFx& ff%d = fxd%d;" % (indent, j, hp.fx_index[t]))
        out.append("%shb_u32* const fa%d = %s;" % (indent, j, _slot_ptr(hp, region, j, bidx, nrows)[1]))
    for j, t in fxs:
        out.append("%sconst hb_u32 fc%d = hb_fx_stage0(fa%d, %s, ff%d);" % (indent, j, j, nrows, j))
    # float64 slots: CAS loop (atomicAdd); "exch" = take / put back with ATOMS.EXCH.64 + STS.64 -- fewer wavefronts on
    # paper, slower in every measurement (c2 129 vs 156 Gev/s, c3 88 vs 116, c4 collapses: lanes that meet the marker spin)
    add = "hb_sadd_f64_exch" if options.get("f64_atomic", "cas") == "exch" else "hb_sadd_f64"
    for j, t in enumerate(slots):
        if t not in hp.fx_set:
            out.append("%s%s(%s, %s);" % (indent, add, _slot_ptr(hp, region, j, bidx, nrows)[1], val(t)))
    for j, t in fxs:
        out.append("%sconst hb_u64 fd%d = hb_fx_stage1(fa%d, %s, ff%d, fc%d);" % (indent, j, j, nrows, j, j))
    if count_u32:
        out.append("%shb_sadd_u32(%s, 1u);" % (indent, _slot_ptr(hp, region, len(slots), bidx, nrows)[1]))
    for j, t in fxs:
        out.append("%shb_fx_stage2(fa%d, %s, ff%d, fd%d);" % (indent, j, nrows, j, j))
    for j, t in fxs:
        if det:
            # does not fit the shared-memory field: the lower global window takes it, exactly
            out.append("%sif (!ff%d.fits && !hb_fxg_add1(%s + %d, fv%d, p.win[%d] - HB_FXG_LOWER)) {" % (
                indent, j, gexact, 6 * j + 3, j, 2 + hp.fx_index[t]))
        else:
            out.append("%sif (!ff%d.fits) {" % (indent, j))
        out.append("%s    atomicAdd(HB_FXMISS, 1u);" % indent)
        for col in hp.dest[t]:
            out.append("%s    hb_red_f64(%s + %d, fv%d);" % (indent, grow, col, j))
        out.append("%s}" % indent)
    return out


def _emit_fx_decodes(plans):
    """C lines for the top of the event function: every distinct fixed-point value once, decoded once."""
    seen, out = set(), []
    for hp in plans:
        if hp.strategy not in ("smem", "window"):
            continue
        for t in hp.f64_terms:
            if t in hp.fx_set and hp.fx_index[t] not in seen:
                i = hp.fx_index[t]
                seen.add(i)
                out.append("const double fxv%d = %s;" % (i, hp.terms[t]))
                out.append("const HbFx fxd%d = hb_fx_decode<%s>(fxv%d, p.win[%d]);" % (i, "true" if hp.term_nonneg[t] else "false", i, 2 + i))
    return out


def _emit_row_merge(hp, options, region, bidx, nrows, indent, rowidx=None):
    """C lines adding shared-memory row ``bidx`` to the global row ``row`` (float64 REDs, zero slots skipped); in
    deterministic mode the fixed-point limbs go to the exact accumulators of global row ``rowidx`` as integers."""
    slots, count_u32 = hp.row_layout(options)
    det = bool(options.get("deterministic", False)) and hp.fx
    out = []
    if det:
        out.append("%shb_u64* const grow = p.fxg[%d] + (hb_i64)(%s) * %d;" % (indent, hp.hid, rowidx, 6 * len(slots)))
    for j, t in enumerate(slots):
        ptr = _slot_ptr(hp, region, j, bidx, nrows, const=True)[1]
        if t in hp.fx_set and det:
            out.append("%s{ const hb_u32* const fa = %s;" % (indent, ptr))
            out.append("%s  const hb_u64 ml = ((hb_u64)fa[%s] << 32) | fa[0], mh = ((hb_u64)fa[3 * %s] << 32) | fa[2 * %s];" % (indent, nrows, nrows, nrows))
            out.append("%s  hb_fxg_add128(grow + %d, ml, mh, (hb_u64)((hb_i64)mh >> 63)); }" % (indent, 6 * j))
            continue
        if t in hp.fx_set:
            out.append("%s{ const double v = hb_fx_read(%s, %s, p.win[%d] - 1075);" % (indent, ptr, nrows, 2 + hp.fx_index[t]))
        else:
            out.append("%s{ const double v = *%s;" % (indent, ptr))
        out.append("%s  if (v != 0.0) {" % indent)
        for col in hp.dest[t]:
            out.append("%s    hb_red_f64(row + %d, v);" % (indent, col))
        out.append("%s  } }" % indent)
    if count_u32:
        out.append("%s{ const hb_u32 c = *%s;" % (indent, _slot_ptr(hp, region, len(slots), bidx, nrows, const=True)[1]))
        out.append("%s  if (c != 0u) { const double v = (double)c;" % indent)
        for col in hp.dest[hp.count_term]:
            out.append("%s    hb_red_f64(row + %d, v);" % (indent, col))
        out.append("%s  } }" % indent)
    return out


def _emit_global_adds(hp, options, rowidx, indent):
    """C lines adding one event's terms to the histogram in global memory: float64 REDs, or -- deterministic mode --
    integer adds to the exact accumulators p.fxg (a term that does not fit them is still a float64 RED, counted)."""
    out = []
    det = bool(options.get("deterministic", False))
    out.append("%sdouble* const row = p.hist[%d] + (hb_i64)(%s) * %d;" % (indent, hp.hid, rowidx, hp.ncol))
    if det and hp.fx:
        ns = len(hp.f64_terms)
        out.append("%shb_u64* const grow = p.fxg[%d] + (hb_i64)(%s) * %d;" % (indent, hp.hid, rowidx, 6 * ns))
    for t, e in enumerate(hp.terms):
        if e is not None and det and hp.fx:
            j = hp.f64_terms.index(t)
            out.append("%s{ const double v = %s;" % (indent, e))
            out.append("%s  if (!hb_fxg_add(grow + %d, v, p.win[%d])) {" % (indent, 6 * j, 2 + hp.fx_index[t]))
            out.append("%s    atomicAdd(HB_FXMISS, 1u);" % indent)
            for col in hp.dest[t]:
                out.append("%s    hb_red_f64(row + %d, v);" % (indent, col))
            out.append("%s  } }" % indent)
            continue
        val = "1.0" if e is None else e
        if len(hp.dest[t]) > 1:
            out.append("%s{ const double v = %s;" % (indent, val))
            for col in hp.dest[t]:
                out.append("%s  hb_red_f64(row + %d, v);" % (indent, col))
            out.append("%s}" % indent)
        else:
            out.append("%shb_red_f64(row + %d, %s);" % (indent, hp.dest[t][0], val))
    return out


def split_table(edges, low, scale, numbins):
    """Table that turns the index of a closed-low float64 ``bin`` axis into the digitize count of a ``split`` axis on the SAME
    value: [(edge, count)] per regular bin b, or None where the look-up is not valid.

    T(v) = fl(fl(v - low) * scale) is what hb_bin_axis floors; it is monotone in v (two correctly rounded monotone
    operations), so with k_e = floor(T(e)) for an edge e and b = floor(T(v)):  b > k_e  implies  v > e,  and  b < k_e  implies
    v < e.  Only an edge with k_e == b has to be compared with v itself.  Hence, where every regular bin holds at most one
    such edge, the number of edges <= v is  count[b] + (edge[b] <= v)  with count[b] = #{e: k_e < b} and edge[b] = the edge
    in bin b, +inf if there is none (calc/__init__.py:286-290 computes it as numpy.digitize, hb_split_index from a compare
    chain).  Values whose bin index is not a regular bin (underflow, overflow, NaN, the int32 cast quirk) keep the chain."""
    edges = [float(e) for e in edges]
    if not edges or any(e != e for e in edges) or any(a >= b for a, b in zip(edges, edges[1:])):
        return None
    ks = []
    for e in edges:
        with numpy.errstate(all="ignore"):
            t = (numpy.float64(e) - numpy.float64(low)) * numpy.float64(scale)
        if t != t:
            return None
        ks.append(float(numpy.floor(t)))
    table = []
    for b in range(int(numbins)):
        inside = [e for e, k in zip(edges, ks) if k == float(b)]
        if len(inside) > 1:
            return None
        table.append((inside[0] if inside else float("inf"), sum(1 for k in ks if k < float(b))))
    return table


def _apply_split_tables(prog, layout, smem, cap):
    """Big books: split axes on a value that also has a closed-low float64 bin axis take their index from that axis'
    bin index and ONE 16-byte shared-memory load (``split_table``) instead of a chain of compares against every edge -- 22
    instructions per axis and event become 8 on c4, where the event phase is bound by issue slots, not by shared memory.
    Re-writes the statements in prog.body; returns the new shared-memory size."""
    tables = []                                             # [(constant name, byte offset, entries)]
    shared = {}
    for site in prog.split_sites:
        cands = [b for b in prog.bin_sites.get(site["v"], []) if b["at"] < site["at"]]
        best = None
        for b in sorted(cands, key=lambda b: -b["numbins"]):
            tab = split_table(site["edges"], b["low"], b["scale"], b["numbins"])
            if tab is not None:
                best = (b, tab)
                break
        if best is None:
            continue
        b, tab = best
        key = tuple(tab)
        if key not in shared:
            off = (smem + 15) // 16 * 16
            if off + 16 * len(tab) > cap:
                continue
            cname = "hb_stab%d" % len(tables)
            words = []
            for edge, count in tab:
                bits = int(numpy.array([edge], dtype=numpy.float64).view(numpy.uint64)[0])
                lo, hi = bits & 0xFFFFFFFF, bits >> 32
                words.extend([lo - (1 << 32) if lo >= (1 << 31) else lo, hi - (1 << 32) if hi >= (1 << 31) else hi, int(count), 0])
            prog.consts.append("__constant__ int %s[%d] = {%s};" % (cname, len(words), ", ".join(
                "(-2147483647 - 1)" if w == -(1 << 31) else str(w) for w in words)))
            shared[key] = (cname, off)
            tables.append((cname, off, len(tab)))
            smem = off + 16 * len(tab)
        cname, off = shared[key]
        fast = "hb_split_index<%s>(sq.z + (int)(se <= %s), %s, false, %d, %s)" % (
            site["flags"], site["vv"], ("(se == %s)" % site["vv"]) if site["closedhigh"] else "false", len(site["edges"]), site["ok"])
        # (a bin axis without a NaN bin leaves the index of a NaN value wherever the conversion put it -- bin 0 -- and only
        #  masks the row: the value itself has to be looked at then)
        regular = "sfb < %du" % b["numbins"] if b["nanflow"] else "sfb < %du && %s == %s" % (b["numbins"], site["vv"], site["vv"])
        prog.body[site["at"]] = (
            "int %s; { const unsigned sfb = (unsigned)(%s - %d); if (" + regular.replace("%", "%%") + ") { const int4 sq = reinterpret_cast<const int4*>(hb_smem + %d)[sfb]; "
            "const double se = __hiloint2double(sq.y, sq.x); %s = %s; } else { %s = %s } }") % (
                site["name"], b["i"], b["shift"], off, site["name"], fast, site["name"], site["slow"])
    layout["split_tables"] = tables
    return smem


def generate(spec, coltypes, options=None):
    """The fill kernel of a spec for one set of column dtypes."""
    options = options or {}
    prog = Program(spec, coltypes)
    plans = [plan_hist(prog, hid, h) for hid, h in enumerate(spec["hists"])]
    if len(plans) > 128:
        raise UnsupportedError("more than 128 histograms in one fill")
    layout = {}
    smem, window = choose_strategies(plans, options, prog, layout)
    # (measured on c4: 17.2 Gev/s without, 16.1 with -- the 16-byte look-ups by 32 different bins cost more shared-memory
    #  wavefronts than the 14 instructions they save are worth, the kernel being bound by that pipe; an option, off by default)
    if layout and not layout.get("pipe") and options.get("split_table", False):
        smem = _apply_split_tables(prog, layout, smem, int(options.get("sort_smem", SMEM_MAX - 1024)))
    if len(prog.cols) > 32:
        raise UnsupportedError("more than 32 input fields in one fill")
    # resident threads per SM stay near 1024 whatever the shared-memory footprint: the fixed-point adds are chains
    # of returning atomics and want warps to hide them
    if 0 < smem <= 8 * 1024 and not any(hp.groups for hp in plans):
        # tiny tables (c1: 103 bins): the block merge -- every block REDs every non-zero bin into the same few
        # global addresses -- is what costs; 1024-thread blocks quarter the number of blocks
        # (c1 at 1e8 events: 751 -> 818 Gev/s, profiles/r01_sweeps.txt)
        dthreads, dblocks = 1024, 1
    elif smem <= 37 * 1024:
        dthreads, dblocks = THREADS, 2
    elif smem <= 113 * 1024:
        dthreads, dblocks = 512, 2
    elif _unroll(plans, options) == 0:
        dthreads, dblocks = 768, 1            # the one-event-at-a-time code of big books stays within the 85 registers this allows
    else:
        dthreads, dblocks = 512, 1
    threads = int(options.get("threads", dthreads))
    minblocks = int(options.get("min_blocks", dblocks))
    if layout:
        threads, minblocks = layout["threads"], 1
    smem_cap = smem
    if window is not None:
        # one resident block of 1024 threads per SM that owns 226 of the SM's 227 KB of shared memory: with per-row
        # intervals and float64 rows that holds 99.9 % of c2's events (14.3 k bins); measured (float64 CAS variant, 200 KB)
        # 156 Gev/s vs 152 for two 512-thread blocks with 100 KB each (94 %) and 148 for four 256-thread blocks with
        # 56 KB (80 %); the dense variant gains another 1 % from 200 -> 226 KB (profiles/r01_sweeps.txt).
        threads = int(options.get("win_threads", 1024))
        minblocks = int(options.get("win_min_blocks", 1))
        smem_cap = int(options.get("win_smem", 226 * 1024))
        if smem_cap - smem < 4096:
            window.strategy, window = "global", None
            smem_cap = smem

    nfx_total = 1 + max([hp.fx_index[t] for hp in plans for t in hp.f64_terms if t in hp.fx_set] + [-1])
    group_aux = {}
    group_sig = {}                                     # SSA names of a histogram's group goals -> the first histogram with these goals
    group_leader = {}                                  # histogram -> that first histogram (whose slot it may share)
    goal_aux = {}                                      # tree key of a group goal -> aux slot of its key table
    ev = list(prog.body)
    done_goals = set()
    for hp in plans:
        for a, g in enumerate(hp.groups):
            goalkey = hbspec.tree_key(spec["hists"][hp.hid]["group"][a]["goal"])
            if goalkey not in goal_aux:
                goal_aux[goalkey] = len(goal_aux)
            if g.c not in done_goals:
                done_goals.add(g.c)
                ev.extend(_goal_lookup(g, goal_aux[goalkey]))
    ev.extend(_emit_fx_decodes(plans))
    for hp in plans:
        ev.append("// ---- hist %d: shape %s, %s" % (hp.hid, hp.shape, hp.strategy))
        if hp.strategy == "sorted":
            continue                                   # accumulated by its sort group (below)
        idx, conds = hist_index(hp)
        ev.append("{")
        if hp.groups:
            if len(group_aux) + len(goal_aux) >= MAX_AUX - 4:
                raise UnsupportedError("more than {0} histograms with group axes in one fill".format(MAX_AUX - 4 - len(goal_aux)))
            group_aux[hp.hid] = len(goal_aux) + len(group_aux)
            sig = "_".join(g.c for g in hp.groups)
            if sig not in group_sig:
                group_sig[sig] = hp.hid
                ev.insert(len(ev) - 2, "bool okg_%s = false; int slot_%s = -1;" % (sig, sig))       # (in front of this histogram's block)
                ev.extend("    " + x for x in _group_lookup(hp, group_aux[hp.hid]))
            else:
                group_leader[hp.hid] = group_sig[sig]
                ev.extend("    " + x for x in _group_lookup(hp, group_aux[hp.hid], leader=group_sig[sig], share=bool(options.get("group_fast", False))))
            conds = ["okg", "slot >= 0"] + conds
        conds = ["live"] + conds
        # (lane pairing is for histograms whose every event is a global RED; the few lanes that fall outside a hot
        # window are cheaper as two predicated REDs each than as a warp-wide exchange -- unless asked for: win_pair)
        pair = ((hp.strategy == "global" or (hp.strategy == "window" and options.get("win_pair", not options.get("win_dense", False))))
                and options.get("pair_red", True) and not options.get("deterministic", False)
                and hp.ncol == 2 and len(hp.terms) == 2
                and sorted(hp.dest[0] + hp.dest[1]) == [0, 1])
        if hp.strategy == "window" and hp.win_box:
            # inside the box -> shared memory.  p.win[wbase + 2k] = first index of the box on axis k, [wbase + 2k + 1] = its extent
            wbase = 2 + nfx_total
            ev.append("    const bool okw = %s;" % " && ".join(conds))
            tests, wb = [], None
            for a, (v, totbins) in enumerate(hp.index):
                ev.append("    const unsigned wq%d = (unsigned)(%s - p.win[%d]);" % (a, v.c, wbase + 2 * a))
                tests.append("wq%d < (unsigned)p.win[%d]" % (a, wbase + 2 * a + 1))
                wb = "wq%d" % a if wb is None else "(%s) * (unsigned)p.win[%d] + wq%d" % (wb, wbase + 2 * a + 1, a)
            ev.append("    const bool inside = okw && %s;" % " && ".join(tests))
            ev.append("    if (inside) {")
            ev.append("        const int wb = (int)(%s);" % wb)
            ev.extend(_emit_row_adds(hp, options, hp.win_off, "wb", (smem_cap - smem) // hp.row_bytes(options), "        ",
                                     grow="(p.hist[%d] + (hb_i64)(%s) * %d)" % (hp.hid, idx, hp.ncol),
                                     gexact="(p.fxg[%d] + (hb_i64)(%s) * %d)" % (hp.hid, idx, 6 * len(hp.f64_terms))))
            ev.append("    }")
            conds = ["okw", "!inside"]
        elif hp.strategy == "window":
            # inside the hot window -> shared memory.  Row table entry: x = offset of the row's interval in the
            # window, y = (first in-row index << 16) | length; a row without interval has length 0.
            k = hp.win_split
            wr = "0"
            if k:
                wr = hp.index[0][0].c
                for v, totbins in hp.index[1:k]:
                    wr = "(%s) * %d + %s" % (wr, totbins, v.c)
            wc = hp.index[k][0].c
            for v, totbins in hp.index[k + 1:]:
                wc = "(%s) * %d + %s" % (wc, totbins, v.c)
            ev.append("    const bool okw = %s;" % " && ".join(conds))
            if hp.win_table_bits(smem_cap - smem, options) == 32:
                # packed entry: first in-row index << 24 | length << 16 | offset  (one LDS.32 instead of an LDS.64)
                ev.append("    const hb_u32 we = reinterpret_cast<const hb_u32*>(hb_smem + %d)[okw ? (%s) : 0];" % (hp.win_table_off, wr))
                ev.append("    const unsigned wa = (unsigned)((%s) - (int)(we >> 24));" % wc)
                ev.append("    const bool inside = okw && wa < ((we >> 16) & 0xffu);")
                ev.append("    if (inside) {")
                ev.append("        const int wb = (int)((we & 0xffffu) + wa);")
            else:
                ev.append("    const uint2 we = reinterpret_cast<const uint2*>(hb_smem + %d)[okw ? (%s) : 0];" % (hp.win_table_off, wr))
                ev.append("    const unsigned wa = (unsigned)((%s) - (int)(we.y >> 16));" % wc)
                ev.append("    const bool inside = okw && wa < (we.y & 0xffffu);")
                ev.append("    if (inside) {")
                ev.append("        const int wb = (int)(we.x + wa);")
            # (rows are laid out for the window's full capacity, a compile-time constant: every slot and limb sits at an
            #  immediate offset; p.win[0] still limits what is zeroed to the bins in use)
            ev.extend(_emit_row_adds(hp, options, hp.win_off, "wb", (smem_cap - smem) // hp.row_bytes(options), "        ",
                                     grow="(p.hist[%d] + (hb_i64)(%s) * %d)" % (hp.hid, idx, hp.ncol),
                                     gexact="(p.fxg[%d] + (hb_i64)(%s) * %d)" % (hp.hid, idx, 6 * len(hp.f64_terms))))
            ev.append("    }")
            conds = ["okw", "!inside"]
        if pair:
            # Two adjacent float64 columns (sumw, sumw2): lanes 2k / 2k+1 serve each other's event so that one RED
            # instruction carries both columns of 16 events -- L2 handles the pair as one 16-byte sector update
            # (measured 1.9x over two separate REDs, profiles/atomics_microbench_r01.txt).
            t0 = 0 if hp.dest[0] == [0] else 1
            v0 = "1.0" if hp.terms[t0] is None else hp.terms[t0]
            v1 = "1.0" if hp.terms[1 - t0] is None else hp.terms[1 - t0]
            ev.append("    {")
            ev.append("        const bool okrow = %s;" % " && ".join(conds))
            ev.append("        const int bin = okrow ? (%s) : 0;" % idx)
            # (whole warps usually have nothing outside a well-placed window: skip the exchange then)
            vote = "if (__any_sync(0xffffffffu, okrow)) " if hp.strategy == "window" else ""
            if hp.groups:
                ev.append("        %shb_red_pair(p.hist[%d], okrow ? ((hb_i64)slot * %d + bin) : (hb_i64)-1, %s, %s);" % (vote, hp.hid, hp.nbins, v0, v1))
            else:
                ev.append("        %shb_red_pair(p.hist[%d], okrow ? bin : -1, %s, %s);" % (vote, hp.hid, v0, v1))
            ev.append("    }")
            ev.append("}")
            continue
        if hp.skip_zero:
            nz = ["(%s != 0.0)" % e for e in hp.terms if e is not None]
            if nz and len(nz) == len(hp.terms):
                conds = conds + ["(%s)" % " || ".join(nz)]
        if (layout and hp.strategy == "smem" and not hp.groups and not hp.f64_terms and hp.count_term is not None and hp.replicas == 1
                and options.get("sort_branchfree", True)):
            # big books: a count-only histogram is ONE unconditional native atomic -- a masked row adds to a dump word -- instead
            # of a branch around it (BSSY / BRA / BSYNC were 13 % of the instructions of the event phase, profiles/r02_c4_ncu.txt)
            ev.append("    hb_sadd_u32(reinterpret_cast<hb_u32*>(hb_smem + ((%s) ? %d + 4 * (%s) : %d)), 1u);" % (
                " && ".join(conds), hp.smem_f64_off, idx, layout["dump_off"]))
            ev.append("}")
            continue
        if (layout and hp.strategy == "smem" and hp.groups and not hp.f64_terms and hp.count_term is not None
                and options.get("sort_branchfree", True)):
            # the same for a count-only histogram with group axes: slots below the privatised ones take the unconditional atomic,
            # the (rare) rest the global path
            ev.append("    const bool okr = %s;" % " && ".join(conds))
            if options.get("group_fast", False):
                # kernel variant for fills whose histograms hold no more slots than are privatised (the host checks: fillable.fill_hists)
                ev.append("    hb_sadd_u32(reinterpret_cast<hb_u32*>(hb_smem + (okr ? %d + 4 * (slot * %d + (%s)) : %d)), 1u);" % (
                    hp.smem_f64_off, hp.nbins, idx, layout["dump_off"]))
                ev.append("}")
                continue
            ev.append("    const bool insm = okr && slot < %d;" % hp.gcap)
            ev.append("    hb_sadd_u32(reinterpret_cast<hb_u32*>(hb_smem + (insm ? %d + 4 * (slot * %d + (%s)) : %d)), 1u);" % (
                hp.smem_f64_off, hp.nbins, idx, layout["dump_off"]))
            ev.append("    if (okr && !insm) {")
            ev.append("        const int bin = %s;" % idx)
            ev.extend(_emit_global_adds(hp, options, "(hb_i64)slot * %d + bin" % hp.nbins, "        "))
            ev.append("    }")
            ev.append("}")
            continue
        ev.append("    if (%s) {" % (" && ".join(conds) if conds else "true"))
        ev.append("        const int bin = %s;" % idx)
        if hp.strategy == "smem":
            indent = "        "
            if hp.groups:
                ev.append("        if (slot < %d) {" % hp.gcap)
                ev.append("            const int srowi = slot * %d + bin;" % hp.nbins)
                indent = "            "
            elif hp.replicas > 1:
                ev.append("        const int srowi = (int)(threadIdx.x & %du) * %d + bin;" % (hp.replicas - 1, hp.nbins))
            else:
                ev.append("        const int srowi = bin;")
            grow_idx = "srowi" if hp.groups else "bin"
            ev.extend(_emit_row_adds(hp, options, hp.smem_f64_off, "srowi", hp.nrows, indent,
                                     grow="(p.hist[%d] + (hb_i64)%s * %d)" % (hp.hid, grow_idx, hp.ncol),
                                     gexact="(p.fxg[%d] + (hb_i64)%s * %d)" % (hp.hid, grow_idx, 6 * len(hp.f64_terms))))
            if hp.groups:
                ev.append("        } else {")
                ev.extend(_emit_global_adds(hp, options, "(hb_i64)slot * %d + bin" % hp.nbins, "            "))
                ev.append("        }")
        else:
            ev.extend(_emit_global_adds(hp, options, ("(hb_i64)slot * %d + bin" % hp.nbins) if hp.groups else "bin", "        "))
        ev.append("    }")
        ev.append("}")

    merge = []
    for hp in plans:
        if hp.strategy != "smem":
            continue
        merge.append("// merge hist %d" % hp.hid)
        merge.append("for (int b = threadIdx.x; b < %d; b += HB_THREADS) {" % hp.nrows)
        grow_idx = "b" if hp.replicas == 1 else "(b %% %d)" % hp.nbins
        merge.append("    double* const row = p.hist[%d] + (hb_i64)%s * %d;" % (hp.hid, grow_idx, hp.ncol))
        merge.extend(_emit_row_merge(hp, options, hp.smem_f64_off, "b", hp.nrows, "    ", rowidx=grow_idx))
        merge.append("}")

    if window is not None and window.win_box:
        hp = window
        wbase = 2 + nfx_total
        nrows_w = (smem_cap - smem) // hp.row_bytes(options)
        slots_w = hp.row_layout(options)[0]
        paired = (hp.ncol == 2 and len(slots_w) == 2 and hp.count_term is None and sorted(hp.dest[slots_w[0]] + hp.dest[slots_w[1]]) == [0, 1]
                  and options.get("merge_pair", True))
        merge.append("// merge the box of hist %d: bin wb of the box -> its row in the histogram" % hp.hid)
        merge.append("for (int wb0 = 0; wb0 < p.win[1]; wb0 += HB_THREADS) {")
        merge.append("    const int wb = wb0 + (int)threadIdx.x;")
        merge.append("    const bool wlive = wb < p.win[1];")
        merge.append("    unsigned wrest = wlive ? (unsigned)wb : 0u;")
        merge.append("    hb_i64 grow = 0, gstride = 1;")
        for a in range(len(hp.index) - 1, -1, -1):
            merge.append("    { const unsigned wn = (unsigned)max(p.win[%d], 1); grow += (hb_i64)(wrest %% wn + (unsigned)p.win[%d]) * gstride; wrest /= wn; gstride *= %d; }" % (
                wbase + 2 * a + 1, wbase + 2 * a, hp.index[a][1]))
        if paired:
            t0 = slots_w[0] if hp.dest[slots_w[0]] == [0] else slots_w[1]
            t1 = slots_w[1] if t0 == slots_w[0] else slots_w[0]
            vals = ["*%s" % _slot_ptr(hp, hp.win_off, slots_w.index(t), "(wlive ? wb : 0)", nrows_w, const=True)[1] for t in (t0, t1)]
            merge.append("    const double mv0 = wlive ? %s : 0.0;" % vals[0])
            merge.append("    const double mv1 = wlive ? %s : 0.0;" % vals[1])
            merge.append("    hb_red_pair(p.hist[%d], (wlive && (mv0 != 0.0 || mv1 != 0.0)) ? grow : (hb_i64)-1, mv0, mv1);" % hp.hid)
        else:
            merge.append("    if (wlive) {")
            merge.append("        double* const row = p.hist[%d] + grow * %d;" % (hp.hid, hp.ncol))
            merge.extend(_emit_row_merge(hp, options, hp.win_off, "wb", nrows_w, "        ", rowidx="grow"))
            merge.append("    }")
        merge.append("}")
    elif window is not None:
        hp = window
        merge.append("// merge the hot window of hist %d: one warp per table row, lanes along the row's interval" % hp.hid)
        merge.append("for (int wr = threadIdx.x >> 5; wr < %d; wr += HB_THREADS / 32) {" % hp.win_rows)
        if hp.win_table_bits(smem_cap - smem, options) == 32:
            merge.append("    const hb_u32 wp = reinterpret_cast<const hb_u32*>(hb_smem + %d)[wr];" % hp.win_table_off)
            merge.append("    const uint2 we = make_uint2(wp & 0xffffu, ((wp >> 24) << 16) | ((wp >> 16) & 0xffu));")
        else:
            merge.append("    const uint2 we = reinterpret_cast<const uint2*>(hb_smem + %d)[wr];" % hp.win_table_off)
        nrows_w = (smem_cap - smem) // hp.row_bytes(options)
        slots_w = hp.row_layout(options)[0]
        paired = (hp.ncol == 2 and len(slots_w) == 2 and hp.count_term is None and sorted(hp.dest[slots_w[0]] + hp.dest[slots_w[1]]) == [0, 1]
                  and not options.get("deterministic", False) and options.get("merge_pair", True))
        if paired:
            # two columns per bin: lanes 2k / 2k+1 exchange so that one RED instruction carries both halves of 16 bins'
            # rows (hb_red_pair) -- half the sector operations at L2, where all blocks merge the same addresses at once
            t0 = slots_w[0] if hp.dest[slots_w[0]] == [0] else slots_w[1]
            t1 = slots_w[1] if t0 == slots_w[0] else slots_w[0]
            merge.append("    for (unsigned wa0 = 0u; wa0 < (we.y & 0xffffu); wa0 += 32u) {")
            merge.append("        const unsigned wa = wa0 + (threadIdx.x & 31u);")
            merge.append("        const bool wlive = wa < (we.y & 0xffffu);")
            merge.append("        const int wb = wlive ? (int)(we.x + wa) : 0;")
            vals = []
            for t in (t0, t1):
                j = slots_w.index(t)
                ptr = _slot_ptr(hp, hp.win_off, j, "wb", nrows_w, const=True)[1]
                if t in hp.fx_set:
                    vals.append("hb_fx_read(%s, %d, p.win[%d] - 1075)" % (ptr, nrows_w, 2 + hp.fx_index[t]))
                else:
                    vals.append("*%s" % ptr)
            merge.append("        const double mv0 = wlive ? %s : 0.0;" % vals[0])
            merge.append("        const double mv1 = wlive ? %s : 0.0;" % vals[1])
            merge.append("        const int mrow = (wlive && (mv0 != 0.0 || mv1 != 0.0)) ? (int)(wr * %d + (we.y >> 16) + wa) : -1;" % hp.win_trail)
            merge.append("        hb_red_pair(p.hist[%d], mrow, mv0, mv1);" % hp.hid)
        else:
            merge.append("    for (unsigned wa = threadIdx.x & 31u; wa < (we.y & 0xffffu); wa += 32u) {")
            merge.append("        const int wb = (int)(we.x + wa);")
            merge.append("        double* const row = p.hist[%d] + ((hb_i64)wr * %d + (we.y >> 16) + wa) * %d;" % (hp.hid, hp.win_trail, hp.ncol))
            merge.extend(_emit_row_merge(hp, options, hp.win_off, "wb", nrows_w, "        ",
                                         rowidx="(hb_i64)wr * %d + (we.y >> 16) + wa" % hp.win_trail))
        merge.append("    }")
        merge.append("}")

    k = Kernel()
    winfo = None
    naux = len(group_aux) + len(goal_aux)
    if window is not None:
        bpb = window.row_bytes(options)
        winfo = {"hid": window.hid, "bpb": bpb, "wmax": (smem_cap - smem) // bpb, "rows": window.win_rows, "trail": window.win_trail,
                 "ncol": window.ncol, "col": spec["hists"][window.hid]["sumw"], "aux": naux, "table_off": window.win_table_off,
                 "table_bits": window.win_table_bits(smem_cap - smem, options)}
        if window.win_box:
            winfo.update({"box": [t for v, t in window.index], "box_base": 2 + nfx_total, "rows": 0, "aux": None})
        else:
            naux += 1
    k.fx_terms = [(hp.fx_index[t], hp.hid, t) for hp in plans for t in hp.f64_terms if t in hp.fx_set]
    k.nfx = (1 + max(i for i, hid, t in k.fx_terms)) if k.fx_terms else 0
    if k.fx_terms:
        k.fx_aux = naux                 # aux slot of the device counter [terms that missed their fixed-point window]
        naux += 1
        merge.append("if (threadIdx.x == 0 && *HB_FXMISS != 0u && p.aux[%d] != nullptr)" % k.fx_aux)
        merge.append("    atomicAdd(reinterpret_cast<hb_u64*>(const_cast<void*>(p.aux[%d])), (hb_u64)*HB_FXMISS);" % k.fx_aux)
    k.times_aux = None
    if options.get("block_times", False):
        k.times_aux = naux
        naux += 1
    k.naux = naux
    k.deterministic = bool(options.get("deterministic", False))
    # deterministic mode: per histogram [(scale index, destination column mask)] of its float64 slots, in slot order
    k.exact = {}
    if k.deterministic:
        for hp in plans:
            if hp.fx:
                k.exact[hp.hid] = [(hp.fx_index[t], sum(1 << c for c in hp.dest[t])) for t in hp.f64_terms]
    unroll = _unroll(plans, options)
    if layout and layout.get("pipe"):
        k.source = _emit_kernel_sorted_pipe(prog, ev, merge, layout, options)
    elif layout:
        k.source = _emit_kernel_sorted(prog, ev, merge, layout, options)
    else:
        k.source = _emit_kernel(prog, ev, smem, merge, threads, minblocks, window=winfo, prefetch=bool(options.get("prefetch", False)), unroll=unroll,
                                times_aux=k.times_aux, dynamic=False if options.get("prefetch", False) else options.get("dynamic", False),
                                evict_first=bool(options.get("evict_first", True)), red_evict_last=bool(options.get("red_evict_last", False)))
    k.cols = list(prog.cols)
    k.plans = plans
    k.smem_bytes = smem_cap
    k.window = winfo
    k.threads = threads
    k.tile = layout["tile"] if layout else threads * VEC
    k.sorted = dict((key, layout[key]) for key in ("tile", "lanes", "ept")) if layout else None
    if layout and layout.get("pipe"):
        k.sorted["pipe"] = {"producers": layout["pipe"]["prod"], "consumer_warps": len(layout["groups"])}
    if layout:
        k.sorted["groups"] = [{"nbins": g.nbins, "values": len(g.values), "hists": [hp.hid for hp in g.members]} for g in layout["groups"]]
        k.sorted["staged"] = len(layout["staged"])
        k.sorted["flags"] = len(layout.get("flags") or []) if layout.get("chunks") else 0
    k.inexact = set(prog.inexact)
    k.group_aux = group_aux
    k.group_leader = group_leader
    k.goal_aux = goal_aux
    k.key = hashlib.sha256(k.source.encode()).hexdigest()
    return k


def group_goals(spec):
    """Distinct group-axis goals of a spec: [(tree key, tree)] in first-appearance order."""
    seen, out = set(), []
    for h in spec["hists"]:
        for a in h["group"]:
            key = hbspec.tree_key(a["goal"])
            if key not in seen:
                seen.add(key)
                out.append((key, a["goal"]))
    return out


def generate_range(spec, coltypes, options=None):
    """The pilot kernel of the fixed-point terms: evaluates every such term and records the largest binary
    exponent it takes (finite values only) in aux[0][i] -- fillable._fx_scales turns that into the scale of
    term i.  Accumulates nothing."""
    options = options or {}
    prog = Program(spec, coltypes)
    plans = [plan_hist(prog, hid, h) for hid, h in enumerate(spec["hists"])]
    choose_strategies(plans, options, prog, {})
    terms = [(hp, t) for hp in plans for t in hp.f64_terms if t in hp.fx_set]
    if not terms:
        return None
    ev = list(prog.body)
    nfx = 1 + max(hp.fx_index[t] for hp, t in terms)
    done = set()
    for hp, t in terms:
        i = hp.fx_index[t]
        if i in done:
            continue
        done.add(i)
        ev.append("{ const hb_u32 e = ((hb_u32)__double2hiint(%s) >> 20) & 0x7ffu;" % hp.terms[t])
        ev.append("  const hb_u32 m = __reduce_max_sync(0xffffffffu, (live && e != 0x7ffu) ? e : 0u);")
        ev.append("  if ((threadIdx.x & 31u) == 0u && m > reinterpret_cast<hb_u32*>(hb_smem)[%d]) atomicMax(reinterpret_cast<hb_u32*>(hb_smem) + %d, m); }" % (i, i))
    merge = ["for (int i = threadIdx.x; i < %d; i += HB_THREADS) {" % nfx,
             "    const hb_u32 m = reinterpret_cast<const hb_u32*>(hb_smem)[i];",
             "    if (m != 0u) atomicMax(reinterpret_cast<hb_u32*>(const_cast<void*>(p.aux[0])) + i, m);",
             "}"]
    threads = THREADS
    k = Kernel()
    k.smem_bytes = (nfx * 4 + 15) // 16 * 16
    k.nfx = nfx
    k.source = _emit_kernel(prog, ev, k.smem_bytes, merge, threads, 2)
    k.cols = list(prog.cols)
    k.threads = threads
    k.tile = threads * VEC
    k.fx_terms = [(hp.fx_index[t], hp.hid, t) for hp, t in terms]
    k.key = hashlib.sha256(k.source.encode()).hexdigest()
    return k


def generate_keys(spec, coltypes, options=None):
    """The key-discovery kernel: evaluates only the group-axis expressions and inserts every key
    seen into one device key set per distinct group goal (numpy.unique, calc/__init__.py:150,178,182)."""
    options = options or {}
    goals = group_goals(spec)
    if not goals:
        return None
    if len(goals) > 64:
        raise UnsupportedError("more than 64 distinct group axes in one fill")
    prog = Program(spec, coltypes)
    values = [prog.lower(tree) for key, tree in goals]
    ev = list(prog.body)
    infos = []
    for j, g in enumerate(values):
        cond = "if (live) "
        if g.dtype.kind == "f" and not g.extra["nanflow"]:
            cond = "if (live && !(%s != %s)) " % (g.c, g.c)
        ev.append("%shb_keyset_insert_cached(reinterpret_cast<hb_u64*>(const_cast<void*>(p.aux[%d])), hb_keybits(%s), reinterpret_cast<hb_u64*>(hb_smem) + %d);" % (cond, j, g.c, 256 * j))
        infos.append(dict(g.extra, kind="f" if g.dtype.kind == "f" else "i"))
    threads = int(options.get("threads", THREADS))
    k = Kernel()
    k.smem_bytes = 2048 * len(goals)                   # hb_keyset_insert_cached: 256 filter entries per goal
    k.source = _emit_kernel(prog, ev, k.smem_bytes, [], threads, 2)
    k.cols = list(prog.cols)
    k.threads = threads
    k.tile = threads * VEC
    k.group_goals = [(key, info) for (key, tree), info in zip(goals, infos)]
    k.key = hashlib.sha256(k.source.encode()).hexdigest()
    return k
