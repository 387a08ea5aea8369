#!/usr/bin/env python
"""Generate golden fill fixtures by running the UNMODIFIED reference histbook.

Run in the build container (needs /root/reference or baseline/_ref)::

    python tests/golden/make_golden.py

Writes ``tests/golden/cases.json`` (per case: the histbook constructor source, the
neutral fill spec extracted by ``histbook_b200.spec`` and the layout of the
reference's outputs) and ``tests/golden/data.npz`` (inputs and reference outputs).
The reference is imported with histbook_b200's engine NOT installed, so every
output here is produced by the reference's own numpy path
(histbook/fill.py:85-162, histbook/hist.py:392-504).

Case sources: the reference's tests/test_hist.py (line numbers in each case's
``ref`` field), tests/test_book.py, and the five BASELINE.json configs at reduced
N, plus dtype / NaN / flag sweeps that the reference tests do not pin (SURVEY.md
section 8c) -- for those the reference's actual output is the pin.
"""
import json
import os
import sys
import warnings

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
warnings.filterwarnings("ignore")

import histbook_b200                               # noqa: E402
from histbook_b200 import _ref, spec as hbspec    # noqa: E402

histbook_b200.uninstall()                          # the reference's own numpy path produces every output below
hb = _ref.load()
ENV = {n: getattr(hb, n) for n in ("Hist", "Book", "bin", "intbin", "split", "cut", "profile", "groupby", "groupbin")}
ENV["numpy"] = numpy

CASES = []


def case(name, hists, fills, ref=""):
    """hists: list of constructor source strings; fills: list of dicts name -> array-like."""
    CASES.append({"name": name, "hists": hists, "fills": fills, "ref": ref})


nan = numpy.nan
inf = numpy.inf

# ---------------------------------------------------------------- tests/test_hist.py
case("calc", ['Hist(bin("x + 0.1", 10, 0, 1))'], [{"x": [0.4, 0.3, 0.3, 0.5, 0.4, 0.8]}], "tests/test_hist.py:43")
case("defs", ['Hist(bin("y", 10, 0, 1), defs={"y": "x + 0.1"})'], [{"x": [0.4, 0.3, 0.3, 0.5, 0.4, 0.8]}], "tests/test_hist.py:48")
case("nofixed_profile", ['Hist(profile("x"))'], [{"x": [3, 3, 5, 5]}], "tests/test_hist.py:64")
case("nofixed_weight", ['Hist(weight="x")'], [{"x": [3, 3, 5, 5]}], "tests/test_hist.py:68")
case("nofixed_profile_weight", ['Hist(profile("y"), weight="x")'], [{"x": [2], "y": [3]}], "tests/test_hist.py:72")
case("nofixed_plain", ['Hist()'], [{"x": [1, 2, 3]}, {}], "hist.py:444")
case("bin_profile_weight_noflow", ['Hist(bin("y", 1, -100, 100, underflow=False, overflow=False, nanflow=False), profile("y"), weight="x")'],
     [{"x": [2], "y": [3]}], "tests/test_hist.py:76")
case("bin_10_11", ['Hist(bin("x", 10, 10, 11))'], [{"x": [10.4, 10.3, 10.3, 10.5, 10.4, 10.8]}], "tests/test_hist.py:81")
_X = [0.4, 0.3, 123, 99, 0.3, nan, nan, nan, 0.5, -99, 0.4, 0.8]
for u in (True, False):
    for o in (True, False):
        for n in (True, False):
            for l in (True, False):
                if u or o or n or True:
                    case("bin_flags_%d%d%d%d" % (u, o, n, l),
                         ['Hist(bin("x", 10, 0, 1, underflow=%r, overflow=%r, nanflow=%r, closedlow=%r))' % (u, o, n, l)],
                         [{"x": _X}], "tests/test_hist.py:85-103")
_E = [0.0, 0.0001, 0.0001, 0.5, 0.5, 0.5, 0.9999, 0.9999, 0.9999, 0.9999, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0001, 1.0001, 1.0001, 1.0001, 1.0001, 1.0001, 1.5, 1.5, 1.5, 1.5, 1.5, 1.5, 1.5, 1.9999, 1.9999, 1.9999, 1.9999, 1.9999, 1.9999, 1.9999, 1.9999, 2.0, 2.0, 2.0, 2.0, 2.0, 2.0, 2.0, 2.0, 2.0, 2.0001, 2.0001, 2.0001, 2.0001, 2.0001, 2.0001, 2.0001, 2.0001, 2.0001, 2.0001]
case("bin_edges_closedlow", ['Hist(bin("x", 2, 0, 2))'], [{"x": _E}], "tests/test_hist.py:105")
case("bin_edges_closedhigh", ['Hist(bin("x", 2, 0, 2, closedlow=False))'], [{"x": _E}], "tests/test_hist.py:109")
case("bin_bin", ['Hist(bin("x", 3, 0, 3, underflow=False, overflow=False, nanflow=False), bin("y", 5, 0, 5, underflow=False, overflow=False, nanflow=False))'],
     [{"x": [1], "y": [3]}], "tests/test_hist.py:113")
case("bin_weight", ['Hist(bin("x", 10, 10, 11), weight="y")'],
     [{"x": [10.4, 10.3, 10.3, 10.5, 10.4, 10.8], "y": [0.1, 0.1, 0.1, 0.1, 0.1, 1.0]}], "tests/test_hist.py:118")
case("bin_filter", ['Hist(bin("x", 10, 10, 11), filter="y > 0")'],
     [{"x": [10.4, 10.3, 10.3, 10.5, 10.4, 10.8], "y": [-1, -1, -1, 1, 1, 1]}], "tests/test_hist.py:127")
case("bin_weight_filter1", ['Hist(bin("x", 10, 10, 11), weight=2, filter="y > 0")'],
     [{"x": [10.4, 10.3, 10.3, 10.5, 10.4, 10.8], "y": [-1, -1, -1, 1, 1, 1]}], "tests/test_hist.py:136")
case("bin_weight_filter2", ['Hist(bin("x", 10, 10, 11), weight="y", filter="y > 0")'],
     [{"x": [10.4, 10.3, 10.3, 10.5, 10.4, 10.8], "y": [-0.1, -0.1, -0.1, 0.1, 0.1, 1.0]}], "tests/test_hist.py:145")
case("bin_constweight", ['Hist(bin("x", 10, 10, 11), weight=2.5)'],
     [{"x": [10.4, 10.3, 10.3, 10.5, 10.4, 10.8]}], "hist.py:416-418")

# test_bin_big (tests/test_hist.py:154-303): seed 12345, 10 000 events, NaN in x and w
numpy.random.seed(12345)
xdata = numpy.round(numpy.random.normal(0, 1, 10000), 2)
weights = numpy.random.uniform(-2, 2, 10000)
both = numpy.random.randint(0, 9999, 10)
xdata[numpy.random.randint(0, 9999, 10)] = numpy.nan
weights[numpy.random.randint(0, 9999, 10)] = numpy.nan
xdata[both] = numpy.nan
weights[both] = numpy.nan
for u in (True, False):
    for o in (True, False):
        for n in (True, False):
            for l in (True, False):
                flags = "underflow=%r, overflow=%r, nanflow=%r, closedlow=%r" % (u, o, n, l)
                case("bin_big_%d%d%d%d" % (u, o, n, l),
                     ['Hist(bin("x", 10, 0, 1, %s))' % flags, 'Hist(bin("x", 10, 0, 1, %s), weight="w")' % flags],
                     [{"x": xdata, "w": weights}], "tests/test_hist.py:154-303")

for u in (True, False):
    for o in (True, False):
        case("intbin_%d%d" % (u, o), ['Hist(intbin("x", 0, 10, underflow=%r, overflow=%r))' % (u, o)],
             [{"x": list(range(-5, 15))}], "tests/test_hist.py:305-320")
case("intbin_float_nan", ['Hist(intbin("x", -2, 3))'], [{"x": [-3.7, -2.5, -0.5, 0.5, 2.9, 3.0, 3.99, 4.0, nan, inf, -inf, 1e300]}], "calc/__init__.py:258-279")
case("intbin_int64_wrap", ['Hist(intbin("x", 0, 5))'], [{"x": numpy.array([2**32 + 3, -1, 7, 2**31, 5, 0], dtype=numpy.int64)}], "SURVEY A.2")

case("split_single", ['Hist(split("x", 3))'], [{"x": [0, 1, 2, 3, 4, 5, 6]}, {"x": [0, 1, 2, 3, 4, 5, nan]}], "tests/test_hist.py:322-329")
case("split_two", ['Hist(split("x", (2.5, 3.5)))'], [{"x": [0, 1, 2, 3, 4, 5, 6]}], "tests/test_hist.py:331")
case("split_three", ['Hist(split("x", (2.5, 3.5, 5.0)))'], [{"x": [0, 1, 2, 3, 4, 5, 6]}], "tests/test_hist.py:335")
case("split_three_closedhigh", ['Hist(split("x", (2.5, 3.5, 5.0), closedlow=False))'], [{"x": [0, 1, 2, 3, 4, 5, 6]}], "tests/test_hist.py:339")
for u in (True, False):
    for o in (True, False):
        for n in (True, False):
            if u or o or n:
                case("split1_flags_%d%d%d" % (u, o, n), ['Hist(split("x", 3, underflow=%r, overflow=%r, nanflow=%r))' % (u, o, n)],
                     [{"x": [nan, 1, 2, 3, 4, 5, 6]}], "tests/test_hist.py:343-351")
            for l in (True, False):
                case("split3_flags_%d%d%d%d" % (u, o, n, l),
                     ['Hist(split("x", (2.5, 3.5, 5.0), underflow=%r, overflow=%r, nanflow=%r, closedlow=%r))' % (u, o, n, l)],
                     [{"x": [nan, 1, 2, 2.5, 3, 3.5, 4, 5, 6, inf, -inf]}], "tests/test_hist.py:353-362")

case("cut_pred", ['Hist(cut("p"))'], [{"p": [False, True, True]}], "tests/test_hist.py:365")
case("cut_expr", ['Hist(cut("x > 3"))'], [{"x": [1, 2, 3, 4, 5]}], "tests/test_hist.py:369")
case("cut_and_pred", ['Hist(cut("x > 3 and p"))'], [{"x": [1, 2, 3, 4, 5], "p": [True, False, True, False, True]}], "tests/test_hist.py:373")
case("cut_mod", ['Hist(cut("x > 3 and y % 2 == 0"))'], [{"x": [1, 2, 3, 4, 5], "y": [0, 1, 2, 3, 4]}], "tests/test_hist.py:377")
case("cut_split", ['Hist(cut("x > 3 and y % 2 == 0"), split("x", 3.5))'], [{"x": [1, 2, 3, 4, 5], "y": [0, 1, 2, 3, 4]}], "tests/test_hist.py:381")
case("profile1", ['Hist(bin("x", 10, 10, 11), profile("y"))'],
     [{"x": [10.4, 10.3, 10.3, 10.5, 10.4, 10.8], "y": [0.1, 0.1, 0.1, 0.1, 0.1, 1.0]}], "tests/test_hist.py:385")
case("profile2", ['Hist(bin("x", 10, 10, 11), profile("y"), profile("2*y"))'],
     [{"x": [10.4, 10.3, 10.3, 10.5, 10.4, 10.8], "y": [0.1, 0.1, 0.1, 0.1, 0.1, 1.0]}], "tests/test_hist.py:389")

case("groupbin", ['Hist(groupbin("x", 10.0), bin("y", 4, 1.0, 5.0, underflow=False, overflow=False, nanflow=False))'],
     [{"x": [0, 10, 15, 20], "y": [1, 2, 3, 4]}], "tests/test_hist.py:409")
case("groupbin_origin", ['Hist(groupbin("x", 10.0, origin=1.0), bin("y", 4, 1.0, 5.0, underflow=False, overflow=False, nanflow=False))'],
     [{"x": [0, 10, 15, 20], "y": [1, 2, 3, 4]}], "tests/test_hist.py:415")
case("groupbin_nan_refill", ['Hist(groupbin("x", 2.5), bin("y", 4, 1.0, 5.0))', 'Hist(groupbin("x", 2.5, nanflow=False, closedlow=False), bin("y", 4, 1.0, 5.0), weight="y")'],
     [{"x": [0.0, 10, nan, 2.5, -7.5, 5.0], "y": [1.0, 2, 3, 4, 9, nan]}, {"x": [100.0, 10, nan, 2.4], "y": [1.5, 2.5, 3.5, 4.5]}], "calc/__init__.py:156-196")
case("groupby_num", ['Hist(groupby("c"), bin("x", 3, 1.0, 4.0, underflow=False, overflow=False, nanflow=False))'],
     [{"c": [1, 2, 3, 2, 1, 1, 1], "x": [1, 2, 3, 2, 1, 1, 3]}], "tests/test_hist.py:394 (numeric keys)")
case("groupby_groupbin_num", ['Hist(bin("y", 4, 1.0, 5.0, underflow=False, overflow=False, nanflow=False), groupby("c"), groupbin("x", 10.0))'],
     [{"c": [1, 1, 2, 2], "x": [0, 10, 15, 20], "y": [1, 2, 3, 4]}], "tests/test_hist.py:422 (numeric keys)")
case("groupby_str", ['Hist(groupby("c"), bin("x", 3, 1.0, 4.0, underflow=False, overflow=False, nanflow=False))'],
     [{"c": ["one", "two", "three", "two", "one", "one", "one"], "x": [1, 2, 3, 2, 1, 1, 3]}, {"c": ["two", "zero"], "x": [3.5, 1.5]}], "tests/test_hist.py:394 (string keys, plus a refill)")
case("groupby_groupby_str", ['Hist(groupby("c1"), groupby("c2"), bin("x", 1, 1.0, 2.0, underflow=False, overflow=False, nanflow=False))'],
     [{"c1": ["one", "two", "one", "two"], "c2": ["uno", "uno", "dos", "dos"], "x": [1, 1, 1, 1]}], "tests/test_hist.py:401")
case("groupby_groupbin_str", ['Hist(bin("y", 4, 1.0, 5.0, underflow=False, overflow=False, nanflow=False), groupby("c"), groupbin("x", 10.0))',
                              'Hist(groupbin("x", 10.0), bin("y", 4, 1.0, 5.0, underflow=False, overflow=False, nanflow=False), groupby("c"))'],
     [{"c": ["one", "one", "two", "two"], "x": [0, 10, 15, 20], "y": [1, 2, 3, 4]}], "tests/test_hist.py:422-435")
case("groupbin_masked_key", ['Hist(groupbin("x", 1.0), bin("y", 2, 0.0, 2.0, underflow=False, overflow=False, nanflow=False))'],
     [{"x": [0.5, 1.5, 2.5], "y": [0.5, 5.0, 1.5]}], "hist.py:470-478 (key exists although every row is masked)")

# ---------------------------------------------------------------- tests/test_book.py
case("book_fill", ['Hist(bin("x", 2, 0, 2, underflow=False, overflow=False, nanflow=False))',
                   'Hist(bin("x", 2, 0, 2, underflow=False, overflow=False, nanflow=False), bin("y", 3, 0, 3, underflow=False, overflow=False, nanflow=False))',
                   'Hist(split("x", (1, 2, 3)))'],
     [{"x": [1, 1, 1, 0, 0], "y": [0, 1, 2, 2, 1]}, {"x": [1.5, 0.5], "y": [2.5, 0.5]}], "tests/test_book.py:43-56")

# ---------------------------------------------------------------- dtype sweeps
rs = numpy.random.RandomState(7)
_xf = rs.normal(0, 1, 4000)
_xf[::97] = nan
_xf[5] = inf
_xf[6] = -inf
_xf[7] = 1e300
_xf[8] = 2147483648.0 * 0.1
for dt in ("float64", "float32", "int64", "int32", "int16", "uint8", "bool"):
    if dt.startswith("float"):
        col = _xf.astype(dt)
    elif dt == "bool":
        col = (_xf > 0)
    else:
        col = rs.randint(-20 if not dt.startswith("u") else 0, 20, 4000).astype(dt)
    case("dtype_" + dt, ['Hist(bin("x", 10, -3, 3))', 'Hist(bin("x*0.7 + 1", 7, -2, 2, closedlow=False), weight="x")',
                         'Hist(intbin("x", -3, 3), profile("x"))', 'Hist(split("x", (-1.5, 0, 0.7, 1)), cut("x > 0"))'],
         [{"x": col}], "fill.py:101-102,138-139")
case("f32_edge", ['Hist(bin("x", 10, 0, 1))'], [{"x": numpy.array([0.7, 0.1, 0.3, 0.6, 0.9], dtype=numpy.float32)}], "SURVEY A.7")
case("scalar_broadcast", ['Hist(bin("x", 4, 0, 4), weight="w")', 'Hist(bin("x + c", 4, 0, 4), profile("c*x"))'],
     [{"x": [0.5, 1.5, 2.5, 3.5, 3.6], "w": 2.0, "c": 0.25}], "fill.py:112-146")
case("maybeconstants", ['Hist(bin("x * pi", 4, 0, 4), weight="e")'], [{"x": [0.1, 0.5, 0.9, 1.2, 2.0]}], "fill.py:96,133")

# ---------------------------------------------------------------- elementwise library
_a = rs.uniform(-3, 3, 3000)
_b = rs.uniform(0.1, 4, 3000)
_a[::131] = nan
case("library_exact", ['Hist(bin("abs(a)", 10, 0, 3), weight="b")', 'Hist(bin("floor(a) + ceil(b)", 10, -3, 8))',
                       'Hist(bin("trunc(a*1.5)", 9, -4.5, 4.5), bin("rint(b)", 5, 0, 5))', 'Hist(bin("max(a, b) - min(a, b)", 20, 0, 7))',
                       'Hist(bin("a / b", 20, -5, 5), profile("sqrt(b)"))', 'Hist(bin("a % b", 20, 0, 4))', 'Hist(bin("sign(a) * b", 16, -4, 4))',
                       'Hist(bin("where(a > 0, a, -b)", 14, -4, 3))', 'Hist(cut("a > 1 or b < 1"), cut("not (a > 0) and b >= 2"))',
                       'Hist(cut("a == b or a != a"), cut("isnan(a)"))', 'Hist(bin("copysign(b, a)", 8, -4, 4))', 'Hist(bin("heaviside(a, 0.5) + b", 10, 0, 5))',
                       'Hist(bin("a**3 - 2*a**2 + b/3", 25, -60, 20))', 'Hist(bin("1/a + a**-2", 30, -5, 5))', 'Hist(bin("hypot(a, b)", 10, 0, 5))',
                       'Hist(bin("(a + b) * (a - b) * 2", 12, -10, 10))', 'Hist(cut("b in {1, 2, 3}"), cut("isinf(1/a)"))'],
     [{"a": _a, "b": _b}], "calc/__init__.py:38-103; expr.py:633-866")
case("library_transcendental", ['Hist(bin("sin(a) + cos(b)", 10, -2, 2))', 'Hist(bin("exp(a) * log(b)", 20, -10, 30))',
                                'Hist(bin("arctan2(a, b)", 12, -3.2, 3.2))', 'Hist(bin("tanh(a) + log1p(b) + expm1(a/4)", 15, -2, 4))',
                                'Hist(bin("erf(a) + erfc(b)", 10, -1, 2))', 'Hist(bin("lgamma(b) + gamma(b)", 10, -1, 10))',
                                'Hist(bin("b ** a", 10, 0, 10))', 'Hist(bin("log10(b) + log2(b) + exp2(a)", 10, -5, 10))',
                                'Hist(bin("arcsin(a/3) + arccos(a/3) + arctan(a)", 10, -1, 4))', 'Hist(bin("sinh(a) + cosh(a) + tan(a)", 10, -20, 20))',
                                'Hist(bin("arcsinh(a) + arccosh(b+1) + arctanh(a/4)", 10, -4, 6))', 'Hist(bin("deg2rad(a) + rad2deg(b)", 10, 0, 240))',
                                'Hist(bin("logaddexp(a, b) + logaddexp2(a, b)", 10, 0, 10))', 'Hist(bin("factorial(b)", 10, 0, 30))'],
     [{"a": _a, "b": _b}], "calc/__init__.py:38-145 (not bit-identical between libm and CUDA: compared with a flip budget)")

# ---------------------------------------------------------------- BASELINE.json configs at reduced N
N = 20000
rs = numpy.random.RandomState(12345)
x, y, z = rs.normal(0, 1, N), rs.normal(0, 1, N), rs.normal(0, 1, N)
w = rs.uniform(0.5, 1.5, N)
n = rs.poisson(3, N).astype(numpy.int64)
C1 = ['Hist(bin("x", 100, -5, 5))']
C2 = ['Hist(bin("x+y", 200, -5, 5), bin("sqrt(x**2+y**2)", 200, 0, 5), weight="w")']
C3 = ['Hist(bin("x", 1000, -5, 5), profile("y"), profile("y**2"), cut("x>0"))']
C3F = ['Hist(bin("x", 1000, -5, 5), profile("y"), profile("y**2"), filter="x>0")']
C4 = []
for e in ["x", "y", "z", "x+y", "x-y", "x*y", "sqrt(x**2+y**2)", "sqrt(x**2+y**2+z**2)"]:
    C4 += ['Hist(bin(%r, 100, -5, 5))' % e, 'Hist(bin(%r, 100, -5, 5), weight="w")' % e,
           'Hist(split(%r, (-2, -1, -0.5, 0, 0.5, 1, 2)))' % e, 'Hist(bin(%r, 50, -5, 5), intbin("n", 0, 9))' % e,
           'Hist(bin(%r, 50, -5, 5), cut("z > 0"))' % e, 'Hist(bin(%r, 100, -5, 5), profile("z"))' % e,
           'Hist(groupbin("n", 2), bin(%r, 50, -5, 5))' % e, 'Hist(bin(%r, 100, -5, 5), filter="n >= 3", weight="w")' % e]
C5 = ['Hist(bin("x", 200, -5, 5), bin("y", 200, -5, 5), bin("z", 200, -5, 5), weight="w")']
case("c1", C1, [{"x": x}], "BASELINE.json configs[0]")
case("c2", C2, [{"x": x, "y": y, "w": w}], "BASELINE.json configs[1]")
case("c3", C3, [{"x": x, "y": y}], "BASELINE.json configs[2]")
case("c3_filter", C3F, [{"x": x, "y": y}], "BASELINE.json configs[2] (filter reading)")
case("c4", C4, [{"x": x[:5000], "y": y[:5000], "z": z[:5000], "w": w[:5000], "n": n[:5000]},
                {"x": x[5000:8000], "y": y[5000:8000], "z": z[5000:8000], "w": w[5000:8000], "n": n[5000:8000]}], "BASELINE.json configs[3]; SURVEY App. D")
case("c5", C5, [{"x": x, "y": y, "z": z, "w": w}], "BASELINE.json configs[4]")
xn, wn = x.copy(), w.copy()
xn[::1000] = nan
xn[1::1000] = inf
wn[2::1000] = nan
case("c2_nanlaced", C2, [{"x": xn, "y": y, "w": wn}], "SURVEY 8d NaN/inf-laced variant")
case("c5_nanlaced", C5, [{"x": xn, "y": y, "z": z, "w": wn}], "SURVEY 8d NaN/inf-laced variant")


# ---------------------------------------------------------------- run the reference

def flatten(content, path=()):
    """nested dict -> [(keypath, ndarray)] in the reference's insertion order"""
    if content is None:
        return None
    if isinstance(content, numpy.ndarray):
        return [(list(path), content)]
    out = []
    for k, v in content.items():
        kk = k if isinstance(k, str) else (int(k) if isinstance(k, (int, numpy.integer)) else float(k))
        out.extend(flatten(v, path + (kk,)))
    return out


def main():
    meta = []
    data = {}
    for ci, c in enumerate(CASES):
        hists = [eval(src, dict(ENV)) for src in c["hists"]]
        spec = hbspec.book_spec(hists)
        book = hb.Book(dict(("h%03d" % i, h) for i, h in enumerate(hists))) if len(hists) > 1 else None
        fills_meta = []
        for fi, f in enumerate(c["fills"]):
            fm = {}
            for name, value in f.items():
                arr = numpy.asarray(value) if not isinstance(value, (int, float)) else value
                if isinstance(arr, numpy.ndarray):
                    key = "c%d_f%d_%s" % (ci, fi, name)
                    data[key] = arr
                    fm[name] = {"npz": key}
                else:
                    fm[name] = {"scalar": arr}
            fills_meta.append(fm)
            args = dict((name, (numpy.asarray(v) if not isinstance(v, (int, float)) else v)) for name, v in f.items())
            if book is not None:
                book.fill(**args)
            else:
                hists[0].fill(**args)
        outs = []
        for hi, h in enumerate(hists):
            flat = flatten(h._content)
            entry = []
            for ki, (keypath, arr) in enumerate(flat):
                key = "c%d_o%d_%d" % (ci, hi, ki)
                if arr.size > 100000:
                    nz = numpy.flatnonzero(arr)
                    data[key + "_idx"] = nz.astype(numpy.int64)
                    data[key + "_val"] = arr.reshape(-1)[nz]
                    entry.append({"keys": keypath, "sparse": key, "shape": list(arr.shape)})
                else:
                    data[key] = arr
                    entry.append({"keys": keypath, "npz": key})
            outs.append(entry)
        meta.append({"name": c["name"], "ref": c["ref"], "hists": c["hists"], "spec": spec,
                     "fills": fills_meta, "outputs": outs})
        print("%-28s %2d hists, %d fills" % (c["name"], len(hists), len(c["fills"])))
    with open(os.path.join(HERE, "cases.json"), "w") as f:
        json.dump({"numpy": numpy.__version__, "histbook": hb.__version__, "cases": meta}, f, allow_nan=True, sort_keys=True)
    numpy.savez_compressed(os.path.join(HERE, "data.npz"), **data)
    print("wrote %d cases" % len(meta))


if __name__ == "__main__":
    main()
