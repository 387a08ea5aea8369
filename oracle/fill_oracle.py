"""CPU ORACLE for the histbook fill hot path -- TEST INFRASTRUCTURE ONLY.

This file is a plain-numpy restatement of the reference algorithm
(scikit-hep/histbook 1.2.5) for ``Hist.fill`` / ``Book.fill``:

  * expression evaluation      histbook/fill.py:85-162, histbook/calc/__init__.py:36-147,327-357
  * bin / intbin / split / cut histbook/calc/__init__.py:203-325
  * groupby / groupbin         histbook/calc/__init__.py:149-201
  * linearise + scatter-add    histbook/hist.py:392-504

It works on the neutral fill spec of ``histbook_b200/spec.py`` (goal trees that
histbook's own front end produced), so it needs neither histbook nor a GPU.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it -- as the checker, never as the product:
the product path (``histbook_b200.engine``) raises if its CUDA library is missing
and has no CPU fallback.

Parity status: PINNED.  ``tests/golden/*.npz`` hold inputs and outputs produced by
the unmodified reference (imported from /root/reference by
``tests/golden/make_golden.py``) for every fill-path case of the reference's own
``tests/test_hist.py`` / ``tests/test_book.py`` plus the five BASELINE configs at
reduced size; ``tests/test_oracle.py`` checks this oracle against all of them
bit-for-bit.

The arithmetic itself lives in numpy (third-party, ``requirements.txt:1``
``numpy>=1.8.0``, unpinned; 2.3.5 in this image): ufunc loops with NEP-50 dtype
promotion, ``numpy.digitize``, ``numpy.unique``, ``numpy.add.at`` (sequential,
in-input-order float64 adds) and the x86 float->int32 cast (out of range ->
INT_MIN).
"""
import collections
import math

import numpy

INDEXTYPE = numpy.int32          # calc/__init__.py:34
COUNTTYPE = numpy.float64        # hist.py:49

# fill.py:96,133 / expr.py:328-331: names that become constants when not supplied
MAYBECONSTANTS = {"pi": numpy.pi, "Pi": numpy.pi, "e": numpy.e, "E": numpy.e,
                  "inf": numpy.inf, "Inf": numpy.inf, "infinity": numpy.inf, "Infinity": numpy.inf,
                  "nan": numpy.nan, "NaN": numpy.nan, "Nan": numpy.nan}


# ----------------------------------------------------------------------------- elementwise library
# calc/__init__.py:38-103: the name -> numpy ufunc table

LIBRARY = {
    "numpy.add": numpy.add, "numpy.subtract": numpy.subtract, "numpy.multiply": numpy.multiply,
    "numpy.true_divide": numpy.true_divide, "numpy.equal": numpy.equal,
    "numpy.not_equal": numpy.not_equal, "numpy.less": numpy.less,
    "numpy.less_equal": numpy.less_equal, "numpy.isin": numpy.isin,
    "numpy.logical_or": numpy.logical_or, "numpy.logical_and": numpy.logical_and,
    "numpy.logical_not": numpy.logical_not,
    "abs": numpy.absolute, "fabs": numpy.absolute, "arccos": numpy.arccos, "arccosh": numpy.arccosh,
    "arcsin": numpy.arcsin, "arcsinh": numpy.arcsinh, "arctan2": numpy.arctan2,
    "arctan": numpy.arctan, "arctanh": numpy.arctanh, "ceil": numpy.ceil, "conj": numpy.conjugate,
    "copysign": numpy.copysign, "cos": numpy.cos, "cosh": numpy.cosh, "deg2rad": numpy.deg2rad,
    "exp2": numpy.exp2, "exp": numpy.exp, "expm1": numpy.expm1, "floor": numpy.floor,
    "hypot": numpy.hypot, "isfinite": numpy.isfinite, "isinf": numpy.isinf, "isnan": numpy.isnan,
    "log10": numpy.log10, "log1p": numpy.log1p, "log2": numpy.log2,
    "logaddexp2": numpy.logaddexp2, "logaddexp": numpy.logaddexp, "log": numpy.log,
    "max": numpy.maximum, "fmax": numpy.maximum, "min": numpy.minimum, "fmin": numpy.minimum,
    "pow": numpy.power, "rad2deg": numpy.rad2deg,
    # calc/__init__.py:72,95: "fmod" is first numpy.fmod, then overwritten by remainder
    "mod": numpy.remainder, "fmod": numpy.remainder,
    "rint": numpy.rint, "sign": numpy.sign, "sinh": numpy.sinh, "sin": numpy.sin,
    "sqrt": numpy.sqrt, "tanh": numpy.tanh, "tan": numpy.tan, "trunc": numpy.trunc,
    "heaviside": lambda x, middle=0.5: numpy.heaviside(x, middle),        # calc:73-74
    "where": lambda c, yes, no: numpy.where(c, yes, no),                   # calc:147
}

# calc/__init__.py:105-124: Abramowitz & Stegun 7.1.26 erf / erfc
_ERF_A = (0.254829592, -0.284496736, 1.421413741, -1.453152027, 1.061405429)
_ERF_P = 0.3275911


def _erf_common(values):
    a1, a2, a3, a4, a5 = _ERF_A
    sign = numpy.where(values < 0, -1.0, 1.0)
    v = numpy.absolute(values)
    t = 1.0 / (v * _ERF_P + 1)
    y = 1.0 - ((((a5*t + a4)*t + a3)*t + a2)*t + a1)*t * numpy.exp(numpy.negative(numpy.square(v)))
    return sign, y


def _erf(values):
    sign, y = _erf_common(values)
    return sign * y


def _erfc(values):
    sign, y = _erf_common(values)
    return 1.0 - sign * y


# calc/__init__.py:126-145: six-term Lanczos lgamma / gamma / factorial
_LANCZOS = (76.18009173, -86.50532033, 24.01409822, -1.231739516e0, 0.120858003e-2, -0.536382e-5)
_LANCZOS_STP = 2.50662827465


def _lgamma(values):
    x = values - 1.0
    tmp = x + 5.5
    tmp = (x + 0.5) * numpy.log(tmp) - tmp
    ser = numpy.ones(len(values), dtype=numpy.float64)
    for cof in _LANCZOS:
        x = x + 1.0
        ser = ser + cof / x
    return tmp + numpy.log(_LANCZOS_STP * ser)


LIBRARY["erf"] = _erf
LIBRARY["erfc"] = _erfc
LIBRARY["lgamma"] = _lgamma
LIBRARY["gamma"] = lambda values: numpy.exp(_lgamma(values))
LIBRARY["factorial"] = lambda values: numpy.round(numpy.exp(_lgamma(numpy.round(values) + 1)))


# ----------------------------------------------------------------------------- index ops

def op_bin(values, numbins, low, high, underflow, overflow, nanflow, closedlow):
    """calc/__init__.py:203-239.  Returns a numpy.ma int32 array."""
    shift = 1 if underflow else 0
    t = values - float(low)                                   # dtype: f32 stays f32, ints -> f64
    numpy.multiply(t, float(numbins) / float(high - low), t)  # host-computed constant, in place
    if closedlow:
        numpy.floor(t, t)
        if shift != 0:
            numpy.add(t, shift, t)
    else:
        numpy.ceil(t, t)
        numpy.add(t, shift - 1, t)
    with numpy.errstate(invalid="ignore"):
        out = numpy.ma.array(t, dtype=INDEXTYPE)           # x86 cast: NaN/inf/out-of-range -> INT_MIN
        if underflow:
            numpy.maximum(out, 0, out)
        else:
            out[out < 0] = numpy.ma.masked
        if overflow:
            numpy.minimum(out, shift + numbins, out)
        else:
            out[out >= (numbins + shift)] = numpy.ma.masked
        if nanflow:
            out[numpy.isnan(t)] = numbins + (1 if underflow else 0) + (1 if overflow else 0)
        else:
            out[numpy.isnan(t)] = numpy.ma.masked
    return out


def op_intbin(values, lo, hi, underflow, overflow):
    """calc/__init__.py:258-279."""
    shift = 1 if underflow else 0
    with numpy.errstate(invalid="ignore"):
        out = numpy.ma.array(values + (shift - lo), dtype=INDEXTYPE)
    if underflow:
        numpy.maximum(out, 0, out)
    else:
        out[out < 0] = numpy.ma.masked
    if overflow:
        numpy.minimum(out, shift + 1 + hi - lo, out)
    else:
        out[out > (shift + hi - lo)] = numpy.ma.masked
    return out


def op_split(values, edges, underflow, overflow, nanflow, closedlow):
    """calc/__init__.py:286-306."""
    out = numpy.ma.array(numpy.digitize(values, edges), dtype=INDEXTYPE)
    if not closedlow:
        out[numpy.isin(values, edges)] -= 1
    if not overflow:
        out[out == len(edges)] = numpy.ma.masked
    if nanflow:
        out[numpy.isnan(values)] = len(edges) + (1 if overflow else 0)
    else:
        out[numpy.isnan(values)] = numpy.ma.masked
    if not underflow:
        out[out == 0] = numpy.ma.masked
        numpy.subtract(out, 1, out)
    return out


def op_cut(values):
    """calc/__init__.py:325."""
    return numpy.ma.array(values, dtype=INDEXTYPE)


def op_groupby(values):
    """calc/__init__.py:149-154."""
    uniques, inverse = numpy.unique(values, return_inverse=True)
    return uniques, inverse.astype(INDEXTYPE)


def op_groupbin(values, binwidth, origin, nanflow, closedlow):
    """calc/__init__.py:156-196."""
    if origin == 0:
        k = numpy.multiply(values, 1.0 / float(binwidth))
    else:
        k = values - float(origin)
        numpy.multiply(k, 1.0 / float(binwidth), k)
    if closedlow:
        numpy.floor(k, k)
    else:
        numpy.ceil(k, k)
        numpy.subtract(k, 1, k)
    numpy.multiply(k, float(binwidth), k)
    if origin != 0:
        numpy.add(k, float(origin), k)

    ok = numpy.logical_not(numpy.isnan(k))
    if ok.all():
        uniques, inverse = numpy.unique(k, return_inverse=True)
        inverse = inverse.astype(INDEXTYPE)
    else:
        uniques, okinverse = numpy.unique(k[ok], return_inverse=True)
        if nanflow:
            inverse = numpy.full(k.shape, len(uniques), dtype=INDEXTYPE)
            inverse[ok] = okinverse
            uniques = list(uniques) + ["NaN"]
        else:
            inverse = numpy.full(k.shape, -1, dtype=INDEXTYPE)
            inverse[ok] = okinverse
    return uniques, inverse


def _call(fcn, args):
    if fcn.startswith("histbook.bin"):
        f = fcn[len("histbook.bin"):]
        return op_bin(args[0], args[1], args[2], args[3], f[0] == "U", f[1] == "O", f[2] == "N", f[3] == "L")
    if fcn.startswith("histbook.intbin"):
        f = fcn[len("histbook.intbin"):]
        return op_intbin(args[0], args[1], args[2], f[0] == "U", f[1] == "O")
    if fcn.startswith("histbook.split"):
        f = fcn[len("histbook.split"):]
        return op_split(args[0], args[1], f[0] == "U", f[1] == "O", f[2] == "N", f[3] == "L")
    if fcn == "histbook.cut":
        return op_cut(args[0])
    if fcn == "histbook.groupby":
        return op_groupby(args[0])
    if fcn.startswith("histbook.groupbin"):
        f = fcn[len("histbook.groupbin"):]
        return op_groupbin(args[0], args[1], args[2], f[0] == "N", f[1] == "L")
    # note: "x in {1, 2}" reaches numpy.isin with a Python *set* (expr.py:185-193), which numpy
    # turns into a 0-d object array -> every element compares False (reference behaviour, kept).
    if fcn in LIBRARY:
        return LIBRARY[fcn](*args)
    # calc/__init__.py:353-357: anything else (e.g. "round", "//") is an AssertionError at fill time
    raise AssertionError("no library entry for {0!r}".format(fcn))


# ----------------------------------------------------------------------------- interpreter

def _decode(value):
    if isinstance(value, dict):
        if "tuple" in value:
            return tuple(_decode(x) for x in value["tuple"])
        if "set" in value:
            return set(_decode(x) for x in value["set"])
    return value


def _key(tree):
    if isinstance(tree, list):
        return tuple(_key(x) for x in tree)
    if isinstance(tree, dict):
        return tuple(sorted((k, _key(v)) for k, v in tree.items()))
    if isinstance(tree, float):
        return ("f", tree.hex() if not math.isnan(tree) else "nan")
    if isinstance(tree, bool):
        return ("b", tree)
    if isinstance(tree, int):
        return ("i", tree)
    return tree


def _trees(spec):
    for h in spec["hists"]:
        for a in h["group"]:
            yield a["goal"]
        for a in h["fixed"]:
            if a["goal"] is not None:
                yield a["goal"]
        for a in h["profile"]:
            yield a["sumx"]
            yield a["sumx2"]
        if h["weight"] is not None and "w" in h["weight"]:
            yield h["weight"]["w"]
            yield h["weight"]["w2"]


def _leaves(tree, out):
    if tree[0] in ("name", "pred"):
        if tree[1] not in out:
            out.append(tree[1])
    elif tree[0] == "call":
        for x in tree[2:]:
            _leaves(x, out)


def fill_length(spec, arrays):
    """fill.py:88-110: the length of the first non-scalar field; 1 if all are scalars."""
    names = []
    for t in _trees(spec):
        _leaves(t, names)
    for name in sorted(names):
        try:
            array = arrays[name]
        except KeyError:
            if name in MAYBECONSTANTS:
                continue
            raise ValueError("required field {0} not found in fill arguments".format(repr(name)))
        array = numpy.asarray(array)
        if array.shape != ():
            return array.shape[0]
    return 1


class _Interp(object):
    def __init__(self, arrays, length):
        self.arrays = arrays
        self.length = length
        self.memo = {}

    def param(self, name):
        # fill.py:124-146
        try:
            array = self.arrays[name]
        except KeyError:
            if name in MAYBECONSTANTS:
                array = numpy.full(self.length, MAYBECONSTANTS[name])
            else:
                raise ValueError("required field {0} not found in fill arguments".format(repr(name)))
        if not isinstance(array, numpy.ndarray):
            array = numpy.array(array)
        if array.shape == ():
            array = numpy.full(self.length, array)
        if array.shape[0] != self.length:
            raise ValueError("array {0} has len {1} but other arrays have len {2}".format(repr(name), len(array), self.length))
        return array

    def __call__(self, tree):
        kind = tree[0]
        if kind == "const":
            return _decode(tree[1])
        key = _key(tree)
        if key in self.memo:
            return self.memo[key]
        if kind in ("name", "pred"):
            out = self.param(tree[1])
        elif kind == "bconst":
            out = numpy.full(self.length, _decode(tree[2]))
        elif kind == "call":
            with numpy.errstate(all="ignore"):
                out = _call(tree[1], [self(x) for x in tree[2:]])
        else:
            raise AssertionError(tree)
        self.memo[key] = out
        return out


# ----------------------------------------------------------------------------- accumulate (hist.py:392-504)

def _new_content(h, level):
    if level == len(h["group"]):
        return numpy.zeros(tuple(h["shape"]), dtype=COUNTTYPE)
    if h["group"][level]["kind"] == "groupby" and h["group"][level]["keeporder"]:
        return collections.OrderedDict()
    return {}


_EXACT = [False]


def _add_at(target, idx, values):
    """numpy.add.at (hist.py:439-456): sequential float64 adds in input order -- or, with ``exact`` set by
    fill(..., exact=True) (tests of the deterministic mode only), the correctly rounded exact sum of each bin's
    terms (math.fsum) added to the bin once."""
    if not _EXACT[0]:
        numpy.add.at(target, idx, values)
        return
    idx = numpy.asarray(idx)
    values = numpy.broadcast_to(numpy.asarray(values, dtype=numpy.float64), idx.shape)
    order = numpy.argsort(idx, kind="stable")
    sidx, svals = idx[order], values[order]
    starts = numpy.flatnonzero(numpy.concatenate([[True], sidx[1:] != sidx[:-1]])) if len(sidx) else []
    ends = list(starts[1:]) + [len(sidx)] if len(sidx) else []
    for a, b in zip(starts, ends):
        chunk = svals[a:b]
        with numpy.errstate(all="ignore"):
            total = math.fsum(chunk) if numpy.all(numpy.isfinite(chunk)) else float(numpy.sum(chunk))
        target[sidx[a]] += total


def _fill_block(h, content, indexes, sumxs, sumx2s, weight, weight2, length, absolute=False):
    """hist.py:429-456 (fillblock).  ``absolute`` (tests only) accumulates |term| instead of term: the
    per-bin sum of magnitudes that the float tolerance is stated against."""
    A = numpy.absolute if absolute else (lambda v: v)
    ncol = h["shape"][-1]
    flat = content.reshape((-1, ncol))
    for p, sumx, sumx2 in zip(h["profile"], sumxs, sumx2s):
        if indexes is None:
            indexes = numpy.ma.zeros(len(sumx), dtype=INDEXTYPE)
        if weight2 is not None:
            mask = numpy.ma.getmask(indexes)
            if mask is not numpy.ma.nomask:
                weight = weight[~mask]
                weight2 = weight2[~mask]
        _add_at(flat[:, p["sumwx"]], indexes.compressed(), A(sumx * weight))
        _add_at(flat[:, p["sumwx2"]], indexes.compressed(), A(sumx2 * weight))

    if weight2 is None:
        if indexes is None:
            flat[:, h["sumw"]] += A((1 if length is None else length) * weight)
        else:
            _add_at(flat[:, h["sumw"]], indexes.compressed(), A(weight))
    else:
        if indexes is None:
            indexes = numpy.ma.zeros(len(weight), dtype=INDEXTYPE)
        mask = numpy.ma.getmask(indexes)
        if mask is not numpy.ma.nomask:
            weight = weight[~mask]
            weight2 = weight2[~mask]
        _add_at(flat[:, h["sumw"]], indexes.compressed(), A(weight))
        _add_at(flat[:, h["sumw2"]], indexes.compressed(), A(weight2))


def _fill_dict(h, groups, level, content, indexes, sumxs, sumx2s, weight, weight2, allsel, length, absolute=False):
    """hist.py:458-499 (filldict): one boolean selection per unique key."""
    if level == len(groups):
        _fill_block(h, content, indexes, sumxs, sumx2s, weight, weight2, length, absolute)
        return
    uniques, inverse = groups[level]
    for idx, unique in enumerate(uniques):
        if allsel is None:
            sel = (inverse == idx)
        else:
            sel = (inverse[allsel] == idx)
        if unique not in content:
            content[unique] = _new_content(h, level + 1)
        if indexes is None:
            subindexes = numpy.ma.zeros(numpy.count_nonzero(sel), dtype=INDEXTYPE)
        else:
            subindexes = indexes[sel]
        subsumxs = [x[sel] for x in sumxs]
        subsumx2s = [x[sel] for x in sumx2s]
        if weight2 is None:
            subweight, subweight2 = weight, weight2
        else:
            subweight, subweight2 = weight[sel], weight2[sel]
        if allsel is None:
            suball = sel
        else:
            suball = allsel.copy()
            suball[inverse != idx] = False
        _fill_dict(h, groups, level + 1, content[unique], subindexes, subsumxs, subsumx2s,
                   subweight, subweight2, suball, length, absolute)


def _postfill(h, interp, content, length, absolute=False):
    # hist.py:393-405: row-major linearisation in (masked) int32
    indexes = None
    step = 0
    for axis, nbins in zip(h["fixed"], h["shape"]):
        if axis["goal"] is None:
            continue       # _nullaxis contributes no goal (axis.py:1413); only reached through proj.py results
        idx = interp(axis["goal"])
        if step == 0:
            indexes = idx
        else:
            if step == 1:
                indexes = indexes.copy()
            numpy.multiply(indexes, nbins, indexes)
            numpy.add(indexes, idx, indexes)
        step += 1

    sumxs = [interp(p["sumx"]) for p in h["profile"]]
    sumx2s = [interp(p["sumx2"]) for p in h["profile"]]

    # hist.py:413-427
    w = h["weight"]
    if w is None:
        weight, weight2 = 1, None
    elif "const" in w:
        weight = numpy.ones(length) * w["const"]
        weight2 = numpy.ones(length) * w["const2"]
    else:
        weight = interp(w["w"])
        weight2 = interp(w["w2"])
        sel = numpy.isnan(weight)
        if sel.any():
            weight = weight.copy()
            weight2 = weight2.copy()
            weight[sel] = 0.0
            weight2[sel] = 0.0

    groups = [interp(a["goal"]) for a in h["group"]]
    _fill_dict(h, groups, 0, content, indexes, sumxs, sumx2s, weight, weight2, None, length, absolute)


def fill(spec, arrays, contents=None, absolute=False, exact=False):
    """One ``Book.fill`` (book.py:540-586) / ``Hist.fill`` (hist.py:337-379) over ``arrays``.

    ``contents`` is the list of previous per-hist contents (``None`` entries = never
    filled); it is updated in place where possible and returned.  ``exact`` (not the reference's behaviour; the
    yardstick of the device's deterministic mode): every bin receives the correctly rounded exact sum of this
    fill's terms instead of their sequential float64 sum.
    """
    _EXACT[0] = bool(exact)
    try:
        return _fill(spec, arrays, contents, absolute)
    finally:
        _EXACT[0] = False


def _fill(spec, arrays, contents, absolute):
    hists = spec["hists"]
    if contents is None:
        contents = [None] * len(hists)
    contents = list(contents)
    for i, h in enumerate(hists):           # _prefill, hist.py:381-390
        if contents[i] is None:
            contents[i] = _new_content(h, 0)
    length = fill_length(spec, arrays)
    interp = _Interp(arrays, length)
    for i, h in enumerate(hists):
        _postfill(h, interp, contents[i], length, absolute)
    return contents
