"""The CPU oracle against the UNMODIFIED reference, live, on seeded random histogram definitions.

tests/test_oracle.py pins the oracle on the committed golden cases; this widens the pin: for a few hundred randomly
drawn ``Hist(...)`` definitions (axis kinds, flags, expressions, weights, filters, profiles, group axes; data laced with
NaN / +-inf / values far outside the binning) the reference's own numpy fill path (histbook/fill.py:85-162,
histbook/hist.py:392-504, run from ``baseline/_ref`` with the device backend uninstalled) and ``oracle.fill_oracle`` must
agree bit for bit -- same dict keys in the same order, same float64 bits.  Definitions the reference itself refuses
(e.g. profile x weight x masked rows raise inside numpy.ma) are skipped, not counted.

CPU only; needs the reference package (``baseline/_ref``), which also ships to the GPU box.
"""
import warnings

import numpy
import pytest

histbook_b200 = pytest.importorskip("histbook_b200")
if not histbook_b200._ref.available():
    pytest.skip("the histbook package (baseline/_ref) is not available", allow_module_level=True)

import golden_util                                   # noqa: E402
from histbook_b200 import spec as hbspec             # noqa: E402
from oracle import fill_oracle                       # noqa: E402

EXPRS = ["x", "y", "x + y", "x - y", "x * y", "sqrt(x**2 + y**2)", "abs(x)", "x**2", "2*x + 1", "x/(1 + y**2)", "exp(x/4)",
         "x + y + z", "z", "min(x, y)", "where(x > 0, x, y)", "hypot(x, y)", "floor(x) + 0.5", "arctan2(y, x)", "x % 2", "sign(y)*x**2"]
CUTS = ["x > 0", "y <= 0.5", "x + y < 1", "z != 0", "x >= -1 and y < 1"]
WEIGHTS = [None, None, "w", "w*2", "w*w", 2.5]
FILTERS = [None, None, None, "y > -1", "x < 1.5"]


def _flags(rs):
    return "underflow=%r, overflow=%r, nanflow=%r" % tuple(bool(b) for b in rs.randint(0, 2, 3))


def _fixed_axis(rs):
    kind = rs.choice(["bin", "bin", "bin", "intbin", "split", "cut"])
    if kind == "bin":
        lo = float(rs.choice([-5.0, -2.5, 0.0, -1.0]))
        hi = lo + float(rs.choice([1.0, 2.5, 5.0, 10.0]))
        return 'bin(%r, %d, %r, %r, %s, closedlow=%r)' % (str(rs.choice(EXPRS)), int(rs.choice([1, 2, 7, 10, 33, 100])), lo, hi,
                                                          _flags(rs), bool(rs.randint(0, 2)))
    if kind == "intbin":
        lo = int(rs.choice([-1, 0, 1, 2]))
        return 'intbin(%r, %d, %d, underflow=%r, overflow=%r)' % (str(rs.choice(["n", "k", "n + k"])), lo, lo + int(rs.choice([0, 3, 6])),
                                                                  bool(rs.randint(0, 2)), bool(rs.randint(0, 2)))
    if kind == "split":
        edges = tuple(sorted(set(float(e) for e in numpy.round(rs.uniform(-3, 3, int(rs.choice([1, 2, 5]))), 1))))
        return 'split(%r, %r, %s, closedlow=%r)' % (str(rs.choice(EXPRS)), edges, _flags(rs), bool(rs.randint(0, 2)))
    return 'cut(%r)' % str(rs.choice(CUTS))


def _group_axis(rs):
    kind = rs.choice(["groupby", "groupbin", "groupbin"])
    if kind == "groupby":
        return 'groupby(%r)' % str(rs.choice(["n", "k"]))
    return 'groupbin(%r, %r, origin=%r, nanflow=%r, closedlow=%r)' % (str(rs.choice(["x", "y", "x + y", "n"])), float(rs.choice([0.5, 1.0, 2.0])),
                                                                       float(rs.choice([0.0, 0.25])), bool(rs.randint(0, 2)), bool(rs.randint(0, 2)))


def _definition(rs):
    parts = []
    if rs.rand() < 0.3:
        parts.append(_group_axis(rs))
    for _ in range(int(rs.choice([0, 1, 1, 1, 2, 2, 3]))):
        parts.append(_fixed_axis(rs))
    for _ in range(int(rs.choice([0, 0, 0, 1, 2]))):
        parts.append('profile(%r)' % str(rs.choice(EXPRS)))
    if len(set(parts)) != len(parts):
        parts = list(dict.fromkeys(parts))
    weight, filt = WEIGHTS[rs.randint(len(WEIGHTS))], FILTERS[rs.randint(len(FILTERS))]
    if weight is not None:
        parts.append("weight=%r" % (weight,))
    if filt is not None:
        parts.append("filter=%r" % (filt,))
    return "Hist(%s)" % ", ".join(parts)


def _columns(rs, n):
    x, y, z = rs.normal(0, 1.5, n), rs.normal(0.3, 1, n), numpy.round(rs.normal(0, 1, n), 1)
    w = rs.uniform(-2, 2, n)
    for arr, every in ((x, 97), (y, 131), (w, 151)):
        arr[rs.randint(0, n, max(1, n // every))] = numpy.nan
    x[rs.randint(0, n, 3)] = numpy.inf
    y[rs.randint(0, n, 3)] = -numpy.inf
    x[rs.randint(0, n, 3)] = 1e300
    return {"x": x, "y": y, "z": z, "w": w, "n": rs.poisson(3.0, n).astype(numpy.int64), "k": rs.randint(-2, 5, n).astype(numpy.int32)}


def _flat(content):
    if content is None:
        return None
    return golden_util.flatten(content)


@pytest.mark.parametrize("seed", range(8))
def test_oracle_equals_reference_on_random_definitions(seed):
    rs = numpy.random.RandomState(1000 + seed)
    compared = 0
    with histbook_b200.reference_path() as hb, warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = dict((name, getattr(hb, name)) for name in ("Hist", "bin", "intbin", "split", "cut", "profile", "groupby", "groupbin"))
        for _ in range(40):
            source = _definition(rs)
            fills = [_columns(rs, int(rs.choice([1, 17, 500, 2000]))) for _ in range(int(rs.choice([1, 2])))]
            try:
                hist = eval(source, dict(env))
                for cols in fills:
                    hist.fill(**cols)
            except Exception:
                continue                                    # the reference refuses this definition / these rows
            spec = hbspec.book_spec([hist])
            contents = None
            for cols in fills:
                contents = fill_oracle.fill(spec, cols, contents)
            exp, got = _flat(hist._content), _flat(contents[0])
            assert (exp is None) == (got is None), source
            if exp is not None:
                assert [k for k, _ in got] == [k for k, _ in exp], source
                for (k, a), (_, b) in zip(got, exp):
                    assert a.shape == b.shape and a.dtype == numpy.float64, source
                    assert numpy.array_equal(a, b, equal_nan=True), (source, k)
            compared += 1
    assert compared >= 22, compared


@pytest.mark.parametrize("seed", range(4))
def test_oracle_equals_reference_on_random_books(seed):
    """The same through ``Book.fill`` (histbook/book.py:540-586): one pass of shared instructions, every member's
    ``_postfill``; the oracle fills the book's spec in one call."""
    rs = numpy.random.RandomState(3000 + seed)
    compared = 0
    with histbook_b200.reference_path() as hb, warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = dict((name, getattr(hb, name)) for name in ("Hist", "bin", "intbin", "split", "cut", "profile", "groupby", "groupbin"))
        for _ in range(25):
            sources = [_definition(rs) for _ in range(int(rs.choice([2, 3, 4])))]
            fills = [_columns(rs, int(rs.choice([17, 500, 2000]))) for _ in range(int(rs.choice([1, 2])))]
            try:
                hists = [eval(source, dict(env)) for source in sources]
                book = hb.Book(**dict(("h%d" % i, h) for i, h in enumerate(hists)))
                for cols in fills:
                    book.fill(**cols)
            except Exception:
                continue
            members = [book["h%d" % i] for i in range(len(hists))]
            spec = hbspec.book_spec(members)
            contents = None
            for cols in fills:
                contents = fill_oracle.fill(spec, cols, contents)
            assert len(contents) == len(members)
            for source, member, got in zip(sources, members, contents):
                exp, got = _flat(member._content), _flat(got)
                assert (exp is None) == (got is None), source
                if exp is not None:
                    assert [k for k, _ in got] == [k for k, _ in exp], source
                    for (k, a), (_, b) in zip(got, exp):
                        assert a.shape == b.shape and numpy.array_equal(a, b, equal_nan=True), (source, k)
            compared += 1
    assert compared >= 5, compared
