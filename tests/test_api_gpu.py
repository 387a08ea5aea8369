"""The unchanged histbook API on the device backend (run on the B200 with -m gpu).

These read like the reference's own tests/test_hist.py and tests/test_book.py (line numbers cited) because the
contract is that nothing above ``fill`` changes: same constructors, same ``_content`` layout, same merge
algebra, same read side (``table``, ``project``, pickle, JSON).  histbook itself comes from ``baseline/_ref``
(the unmodified reference, pip-installed with --target; git-ignored but shipped to the GPU box).
"""
import pickle

import numpy
import pytest

pytestmark = pytest.mark.gpu

histbook_b200 = pytest.importorskip("histbook_b200")
if not histbook_b200._ref.available():
    pytest.skip("the histbook package (baseline/_ref) is not available", allow_module_level=True)

from histbook import *       # noqa: E402,F401,F403
from histbook import Book, ChannelsBook, Hist, SamplesBook, SystematicsBook, bin, cut, groupbin, groupby, intbin, profile, split  # noqa: E402


def test_backend_is_installed():
    assert histbook_b200.installed()
    assert isinstance(Hist.__dict__["_content"], property)


def test_calc_and_defs():                                    # tests/test_hist.py:43-51
    h = Hist(bin("x + 0.1", 10, 0, 1))
    h.fill(x=[0.4, 0.3, 0.3, 0.5, 0.4, 0.8])
    assert h._content.tolist() == [[0], [0], [0], [0], [0], [2], [2], [1], [0], [0], [1], [0], [0]]
    h = Hist(bin("y", 10, 0, 1), defs={"y": "x + 0.1"})
    h.fill(x=[0.4, 0.3, 0.3, 0.5, 0.4, 0.8])
    assert h._content.tolist() == [[0], [0], [0], [0], [0], [2], [2], [1], [0], [0], [1], [0], [0]]


def test_nofixed():                                          # tests/test_hist.py:64-78
    h = Hist(profile("x"))
    h.fill(x=[3, 3, 5, 5])
    assert h._content.tolist() == [16, 68, 4]
    h = Hist(weight="x")
    h.fill(x=[3, 3, 5, 5])
    assert h._content.tolist() == [16, 68]
    h = Hist(profile("y"), weight="x")
    h.fill(x=[2], y=[3])
    assert h._content.tolist() == [6, 18, 2, 4]


def test_bin_flags():                                        # tests/test_hist.py:80-111
    x = [0.4, 0.3, 123, 99, 0.3, numpy.nan, numpy.nan, numpy.nan, 0.5, -99, 0.4, 0.8]
    h = Hist(bin("x", 10, 0, 1))
    h.fill(x=x)
    assert h._content.tolist() == [[1], [0], [0], [0], [2], [2], [1], [0], [0], [1], [0], [2], [3]]
    h = Hist(bin("x", 10, 0, 1, underflow=False, overflow=False, nanflow=False))
    h.fill(x=x)
    assert h._content.tolist() == [[0], [0], [0], [2], [2], [1], [0], [0], [1], [0]]


def test_weight_filter():                                    # tests/test_hist.py:118-152
    h = Hist(bin("x", 10, 10, 11), weight="y")
    h.fill(x=[10.4, 10.3, 10.3, 10.5, 10.4, 10.8], y=[0.1, 0.1, 0.1, 0.1, 0.1, 1.0])
    assert h._content.tolist() == [[0.0, 0.0], [0.0, 0.0], [0.0, 0.0], [0.0, 0.0], [0.2, 0.020000000000000004], [0.2, 0.020000000000000004], [0.1, 0.010000000000000002], [0.0, 0.0], [0.0, 0.0], [1.0, 1.0], [0.0, 0.0], [0.0, 0.0], [0.0, 0.0]]
    h = Hist(bin("x", 10, 10, 11)).weight(2).filter("y > 0")
    h.fill(x=[10.4, 10.3, 10.3, 10.5, 10.4, 10.8], y=[-1, -1, -1, 1, 1, 1])
    assert h._content.tolist() == [[0, 0], [0, 0], [0, 0], [0, 0], [0, 0], [2, 4], [2, 4], [0, 0], [0, 0], [2, 4], [0, 0], [0, 0], [0, 0]]


def test_intbin_split_cut_profile():                          # tests/test_hist.py:305-392
    h = Hist(intbin("x", 0, 10, underflow=True, overflow=True))
    h.fill(x=range(-5, 15))
    assert h._content.tolist() == [[5], [1], [1], [1], [1], [1], [1], [1], [1], [1], [1], [1], [4]]
    h = Hist(split("x", (2.5, 3.5, 5.0), closedlow=False))
    h.fill(x=[0, 1, 2, 3, 4, 5, 6])
    assert h._content.tolist() == [[3], [1], [2], [1], [0]]
    h = Hist(cut("x > 3 and y % 2 == 0"), split("x", 3.5))
    h.fill(x=[1, 2, 3, 4, 5], y=[0, 1, 2, 3, 4])
    assert h._content.tolist() == [[[3], [1], [0]], [[0], [1], [0]]]
    h = Hist(bin("x", 10, 10, 11), profile("y"), profile("2*y"))
    h.fill(x=[10.4, 10.3, 10.3, 10.5, 10.4, 10.8], y=[0.1, 0.1, 0.1, 0.1, 0.1, 1.0])
    assert h._content[9].tolist() == [1.0, 1.0, 2.0, 4.0, 1.0]
    assert h._content[4].tolist() == [0.2, 0.020000000000000004, 0.4, 0.08000000000000002, 2.0]


def test_groupbin_and_numeric_groupby():                      # tests/test_hist.py:409-435 (numeric keys)
    h = Hist(groupbin("x", 10.0), bin("y", 4, 1.0, 5.0, underflow=False, overflow=False, nanflow=False))
    h.fill(x=[0, 10, 15, 20], y=[1, 2, 3, 4])
    assert h._content[0.0].tolist() == [[1], [0], [0], [0]]
    assert h._content[10.0].tolist() == [[0], [1], [1], [0]]
    assert h._content[20.0].tolist() == [[0], [0], [0], [1]]
    h.fill(x=[30.0, 0.5], y=[1.5, 4.5])                       # refill: new key appears, old ones accumulate
    assert h._content[30.0].tolist() == [[1], [0], [0], [0]]
    assert h._content[0.0].tolist() == [[1], [0], [0], [1]]
    h = Hist(bin("y", 4, 1.0, 5.0, underflow=False, overflow=False, nanflow=False), groupby("c"), groupbin("x", 10.0))
    h.fill(c=[1, 1, 2, 2], x=[0, 10, 15, 20], y=[1, 2, 3, 4])
    assert h._content[1][0.0].tolist() == [[1], [0], [0], [0]]
    assert h._content[2][20.0].tolist() == [[0], [0], [0], [1]]
    assert h.groupkeys(0) == set([1, 2])


def test_groupby_strings():                                   # tests/test_hist.py:394-407,422-435 (string keys, verbatim)
    h = Hist(groupby("c"), bin("x", 3, 1.0, 4.0, underflow=False, overflow=False, nanflow=False))
    h.fill(c=["one", "two", "three", "two", "one", "one", "one"], x=[1, 2, 3, 2, 1, 1, 3])
    assert h._content["one"].tolist() == [[3], [0], [1]]
    assert h._content["two"].tolist() == [[0], [2], [0]]
    assert h._content["three"].tolist() == [[0], [0], [1]]
    assert list(h._content.keys()) == ["one", "three", "two"]                    # numpy.unique order
    h.fill(c=numpy.array(["two", "zero"]), x=[3.5, 1.5])                          # refill: codes stay, a new key appears
    assert h._content["two"].tolist() == [[0], [2], [1]]
    assert h._content["zero"].tolist() == [[1], [0], [0]]
    assert h._content["one"].tolist() == [[3], [0], [1]]

    h = Hist(groupby("c1"), groupby("c2"), bin("x", 1, 1.0, 2.0, underflow=False, overflow=False, nanflow=False))
    h.fill(c1=["one", "two", "one", "two"], c2=["uno", "uno", "dos", "dos"], x=[1, 1, 1, 1])
    assert h._content["one"]["uno"].tolist() == [[1]]
    assert h._content["two"]["uno"].tolist() == [[1]]
    assert h._content["one"]["dos"].tolist() == [[1]]
    assert h._content["two"]["dos"].tolist() == [[1]]

    h = Hist(bin("y", 4, 1.0, 5.0, underflow=False, overflow=False, nanflow=False), groupby("c"), groupbin("x", 10.0))
    h.fill(c=["one", "one", "two", "two"], x=[0, 10, 15, 20], y=[1, 2, 3, 4])
    assert h._content["one"][0.0].tolist() == [[1], [0], [0], [0]]
    assert h._content["one"][10.0].tolist() == [[0], [1], [0], [0]]
    assert h._content["two"][10.0].tolist() == [[0], [0], [1], [0]]
    assert h._content["two"][20.0].tolist() == [[0], [0], [0], [1]]

    h = Hist(groupbin("x", 10.0), bin("y", 4, 1.0, 5.0, underflow=False, overflow=False, nanflow=False), groupby("c"))
    h.fill(c=["one", "one", "two", "two"], x=[0, 10, 15, 20], y=[1, 2, 3, 4])
    assert h._content[0.0]["one"].tolist() == [[1], [0], [0], [0]]
    assert h._content[10.0]["one"].tolist() == [[0], [1], [0], [0]]
    assert h._content[10.0]["two"].tolist() == [[0], [0], [1], [0]]
    assert h._content[20.0]["two"].tolist() == [[0], [0], [0], [1]]

    # against the reference path on random data, and through + (host dicts with string keys go back to the device)
    rs = numpy.random.RandomState(7)
    names = numpy.array(["ttbar", "wjets", "zjets", "qcd", "data"])
    c = names[rs.randint(0, 5, 20000)]
    x = rs.normal(0, 1, 20000)
    w = rs.uniform(0.5, 1.5, 20000)
    h = Hist(groupby("c"), bin("x", 20, -3, 3), weight="w")
    h.fill(c=c, x=x, w=w)
    with histbook_b200.reference_path() as hb:
        r = hb.Hist(hb.groupby("c"), hb.bin("x", 20, -3, 3), weight="w")
        r.fill(c=c, x=x, w=w)
        expected = dict((str(k), v.copy()) for k, v in r._content.items())
    assert sorted(str(k) for k in h._content.keys()) == sorted(expected)
    for k, v in h._content.items():
        assert numpy.allclose(v, expected[str(k)], rtol=1e-12, atol=0)
    both = h + h
    both.fill(c=c[:100], x=x[:100], w=w[:100])
    assert numpy.all(both._content["qcd"][:, 0] >= 2 * h._content["qcd"][:, 0])


def test_project_on_the_device():                             # histbook/proj.py:227-296
    rs = numpy.random.RandomState(11)
    n = 200000
    x, y, z, w = rs.normal(0, 1, n), rs.normal(0, 1, n), rs.normal(0, 1, n), rs.uniform(0.5, 1.5, n)
    h = Hist(bin("x", 20, -3, 3), bin("y", 15, -3, 3), split("z", (-1, 0, 1)), weight="w")
    h.fill(x=x, y=y, z=z, w=w)
    full = histbook_b200.peek(h).copy()                        # (23, 18, 6, 2)
    for keep, axes in ((("x",), (1, 2)), (("y", "z"), (0,)), ((0, 2), (1,)), ((), (0, 1, 2)), (("x", "y", "z"), ())):
        p = h.project(*keep)
        st = p.__dict__["_hb"]
        assert st.dev_valid and not st.host_valid              # computed and kept on the device
        exp = numpy.sum(full, axes)
        bound = numpy.sum(numpy.abs(full), axes)
        got = p._content
        assert got.shape == exp.shape == tuple(p._shape)
        assert numpy.all(numpy.abs(got - exp) <= 1e-12 * bound)
    # the projection is an ordinary histogram: it can be filled further and tabulated
    p = h.project("x")
    p.fill(x=x[:1000], w=w[:1000])
    assert abs(p._content[:, 0].sum() - (w.sum() + w[:1000].sum())) <= 1e-9
    # unweighted counts stay exact; profile columns are carried along
    c = Hist(bin("x", 50, -3, 3), cut("y > 0"), profile("z"))
    c.fill(x=x, y=y, z=z)
    q = c.project("x")
    assert q.__dict__["_hb"].dev_valid and not q.__dict__["_hb"].host_valid
    ref = histbook_b200.peek(c).sum(axis=1)
    assert numpy.array_equal(q._content[:, 2], ref[:, 2]) and q._content[:, 2].sum() == n
    assert numpy.allclose(q._content[:, :2], ref[:, :2], rtol=1e-12, atol=1e-12)
    assert len(q.table()) == 53
    # with group axes the reference's own (host) path is taken
    g = Hist(groupby("c"), bin("x", 5, -3, 3), bin("y", 4, -3, 3))
    g.fill(c=rs.randint(0, 3, 1000), x=x[:1000], y=y[:1000])
    assert sum(v.sum() for v in g.project("c", "x")._content.values()) == 1000


def _host_twin(h):
    """A copy of ``h`` whose bins are host-resident (a plain numpy ``_content``): every read-side method takes the
    reference's own code on it."""
    t = h.copy()
    t._content = numpy.array(histbook_b200.peek(h))
    st = t.__dict__["_hb"]
    assert st.host_valid
    return t


def test_table_select_and_fraction_on_the_device():           # histbook/proj.py:298-496 (select), 498-640 (table), 642-806 (fraction)
    """The read side of a device-resident histogram: ``table`` computes its columns on the device (hb_arena_table), ``select``
    slices there (hb_arena_slice) and the result stays on the device, ``fraction`` is built from both.  Every result is compared
    bit for bit with the reference's own numpy code run on a host copy of the same bins."""
    rs = numpy.random.RandomState(5)
    n = 300000
    x, y, z, w = rs.normal(0, 1, n), rs.normal(0, 1, n), rs.normal(0, 1, n), rs.uniform(-0.5, 1.5, n)     # (some negative weights)
    hists = [Hist(bin("x", 40, -3, 3), cut("y > 0"), profile("z"), profile("z**2"), weight="w"),
             Hist(bin("x", 40, -3, 3), split("y", (-1, 0, 1)), profile("z")),
             Hist(bin("x", 25, -2, 2, underflow=False, nanflow=False), intbin("k", 0, 4), weight="w"),
             Hist(bin("x", 30, -3, 3), filter="y > 0")]
    k = rs.randint(-1, 7, n)
    for h in hists:
        h.fill(x=x, y=y, z=z, w=w, k=k)
        twin = _host_twin(h)
        st = h.__dict__["_hb"]
        profs = [str(p.expr) for p in h._profile]
        with numpy.errstate(all="ignore"):
            for opts in ({}, {"error": False}, {"effcount": True}, {"count": False, "effcount": True}, {"recarray": False}, {"columns": True},
                         {"count": False, "error": False, "effcount": True}):
                for use in ([], profs[:1], profs):
                    if not use and not opts.get("count", True) and not opts.get("effcount", False):
                        continue
                    got, exp = h.table(*use, **opts), twin.table(*use, **opts)
                    if opts.get("columns"):
                        assert got[1] == exp[1]
                        got, exp = got[0], exp[0]
                    assert got.dtype == exp.dtype and got.shape == exp.shape, (opts, use)
                    assert got.tobytes() == exp.tobytes(), (opts, use)        # bit for bit, NaN included
            assert st.dev_valid and not st.host_valid                          # table() did not pull the bins to the host
            # normalized tables take the reference's path (host) and still agree
            a, b = h.copy().table(normalized=True), twin.table(normalized=True)
            assert a.tobytes() == b.tobytes()
    with pytest.raises(TypeError):
        hists[0].table(bogus=True)
    with pytest.raises(IndexError):
        hists[0].table("nosuch")

    # select: slices along one fixed axis, on the device
    h = hists[0]
    twin = _host_twin(h)
    for expr in ("x < 0", "x >= -1.5", "x >= -1.5 and x < 1.5", "y > 0", "x < 0 and y > 0"):
        got, exp = h.select(expr), twin.select(expr)
        gst = got.__dict__["_hb"]
        assert gst.dev_valid and not gst.host_valid, expr                      # sliced and kept on the device
        assert got == exp, expr                                                # axes and contents (Hist.__eq__)
        assert numpy.array_equal(got._content, exp._content) and got._content.shape == tuple(got._shape)
    h2, twin2 = hists[1], _host_twin(hists[1])
    for expr in ("y >= 0", "y < -1", "x >= 0 and y >= 0 and y < 1"):
        got, exp = h2.select(expr), twin2.select(expr)
        assert got == exp and numpy.array_equal(got._content, exp._content), expr
    with pytest.raises(ValueError):
        h.select("x < 0.01")                                                   # not a bin edge: same error as the reference
    # a selected histogram is an ordinary one: fill it further, tabulate it
    sel = h.select("x >= 0")
    before = sel._content.copy()
    sel.fill(x=x[:1000], y=y[:1000], z=z[:1000], w=w[:1000])
    assert not numpy.array_equal(sel._content, before)

    # drop: every profile dropped = the trailing [sumw, sumw2] columns, on the device
    d, dt = hists[0].drop(*hists[0]._profile), _host_twin(hists[0]).drop(*hists[0]._profile)
    assert d.__dict__["_hb"].dev_valid and not d.__dict__["_hb"].host_valid
    assert d == dt and numpy.array_equal(d._content, dt._content)
    d, dt = hists[1].drop("z"), _host_twin(hists[1]).drop("z")                 # unweighted: [sumw]
    assert d == dt and numpy.array_equal(d._content, dt._content)
    with pytest.raises(IndexError):
        hists[0].drop("nosuch")

    # fraction (efficiency of a cut with its interval): drop + project + select underneath, all on the device
    with numpy.errstate(all="ignore"):
        for kw in ({}, {"error": "normal"}, {"error": "wilson"}, {"level": 2}):
            try:
                exp = _host_twin(hists[0]).fraction("y > 0", **kw)
            except Exception as e:                                             # (whatever the reference raises, e.g. for a missing scipy)
                with pytest.raises(type(e)):
                    hists[0].fraction("y > 0", **kw)
                continue
            got = hists[0].fraction("y > 0", **kw)
            assert hists[0].__dict__["_hb"].dev_valid and not hists[0].__dict__["_hb"].host_valid      # the bins never left the device
            assert got.dtype == exp.dtype and got.shape == exp.shape
            for name in got.dtype.names:
                assert numpy.allclose(got[name], exp[name], rtol=1e-12, atol=0, equal_nan=True), (kw, name)


def test_pickle_and_json_roundtrip():                         # tests/test_hist.py:437-449
    h = Hist(split("x", (1, 2, 3)), bin("y", 10, 0, 1), defs={"y": "x + 0.1"}, weight="sqrt(x)", filter="x > 2")
    assert h == pickle.loads(pickle.dumps(h))
    h.fill(x=[1, 2, 3])
    assert h == pickle.loads(pickle.dumps(h))
    assert h == Hist.fromjson(h.tojson())
    h2 = pickle.loads(pickle.dumps(h))
    h2.fill(x=[3, 3])                                          # host content goes back to the device and accumulates
    assert h2._content.sum() > h._content.sum()


def test_book_fill_and_algebra():                             # tests/test_book.py:43-142
    b = Book()
    b["one"] = Hist(bin("x", 2, 0, 3, underflow=False, overflow=False, nanflow=False))
    b["two"] = Hist(split("y", 1.5, nanflow=False))
    b.fill(x=[1, 1, 1, 2, 2], y=[1, 1, 1, 2, 2])
    assert b["one"]._content.tolist() == [[3], [2]]
    assert b["two"]._content.tolist() == [[3], [2]]

    book1, book2 = Book(), Book()
    book1["a"] = Hist(bin("x", 4, 0, 4), fill=[0, 0, 0])
    book1["b"] = Hist(bin("x", 4, 0, 4), fill=[0, 1, 1])
    book1["d"] = Hist(bin("x", 4, 0, 4), fill=[])
    book2["a"] = Hist(bin("x", 4, 0, 4), fill=[2, 2])
    book2["b"] = Hist(bin("x", 4, 0, 4), fill=[0, 3, 3, 3, 3])
    book2["c"] = Hist(bin("x", 4, 0, 4), fill=[0, 1, 2, 3])
    book = book1 + book2
    assert book["a"][:].tolist() == [[0], [3], [0], [2], [0], [0], [0]]
    assert book["b"][:].tolist() == [[0], [2], [2], [0], [4], [0], [0]]
    assert book["d"][:].tolist() == [[0], [0], [0], [0], [0], [0], [0]]
    assert book["c"][:].tolist() == [[0], [1], [1], [1], [1], [0], [0]]
    assert book1["a"][:].tolist() == [[0], [3], [0], [0], [0], [0], [0]]      # operands untouched
    book3 = book1 * 1.5
    assert book3["b"][:].tolist() == [[0], [1.5], [3], [0], [0], [0], [0]]
    book1 += book2
    assert book1["b"][:].tolist() == [[0], [2], [2], [0], [4], [0], [0]]
    book1 *= 2
    assert book1["a"][:].tolist() == [[0], [6], [0], [4], [0], [0], [0]]
    book1.fill(x=[3.5])                                       # host-modified contents are uploaded again
    assert book1["a"][:].tolist() == [[0], [6], [0], [4], [1], [0], [0]]


def test_merge_algebra_stays_on_the_device():                 # Hist.__add__/__iadd__/__mul__/__imul__ histbook/hist.py:506-607
    rs = numpy.random.RandomState(7)
    cols1 = {"x": rs.normal(0, 1, 20000), "y": rs.normal(0, 1, 20000), "w": rs.uniform(0.5, 1.5, 20000)}
    cols2 = {"x": rs.normal(0.5, 1, 30000), "y": rs.normal(0, 2, 30000), "w": rs.uniform(0.5, 1.5, 30000)}

    def make():
        return Hist(bin("x", 40, -4, 4), bin("y", 30, -4, 4), weight="w")
    a, b = make(), make()
    a.fill(**cols1)
    b.fill(**cols2)
    ca, cb = histbook_b200.peek(a).copy(), histbook_b200.peek(b).copy()

    def on_device(h):
        st = h.__dict__["_hb"]
        return st.dev_valid and not st.host_valid
    c = a + b
    assert on_device(a) and on_device(b) and on_device(c)
    assert numpy.array_equal(histbook_b200.peek(c), ca + cb)          # one IEEE add per bin, like numpy
    d = c * 0.3
    e = 3 * c
    assert on_device(c) and on_device(d) and on_device(e)
    assert numpy.array_equal(histbook_b200.peek(d), (ca + cb) * 0.3)
    assert numpy.array_equal(histbook_b200.peek(e), (ca + cb) * 3)
    a += b
    assert on_device(a) and numpy.array_equal(histbook_b200.peek(a), ca + cb)
    assert numpy.array_equal(histbook_b200.peek(b), cb)                # the right operand is untouched
    a *= 2
    assert on_device(a) and numpy.array_equal(histbook_b200.peek(a), (ca + cb) * 2)
    a.fill(**cols1)                                                   # and keeps filling
    assert abs(a._content[..., 0].sum() - (2 * (ca + cb)[..., 0].sum() + ca[..., 0].sum())) < 1e-6
    # results do not alias their operands
    c.fill(**cols1)
    assert numpy.array_equal(histbook_b200.peek(b), cb) and numpy.array_equal(histbook_b200.peek(d), (ca + cb) * 0.3)
    # the reference's errors
    with pytest.raises(TypeError):
        a + Hist(bin("x", 41, -4, 4), bin("y", 30, -4, 4), weight="w")
    with pytest.raises(TypeError):
        a + 1
    with pytest.raises(TypeError):
        a * "2"
    # one side read on the host (host copy authoritative) or never filled: the reference's own path, same numbers
    f = make()
    f.fill(**cols2)
    cf = f._content.copy()
    g = b + f
    assert numpy.array_equal(g._content, cb + cf)
    assert numpy.array_equal((make() + b)._content, cb)


def test_group():                                             # tests/test_book.py:144-158, Hist.group histbook/hist.py:609-659
    book1, book2 = Book(), Book()
    book1["a"] = Hist(bin("x", 4, 0, 4), fill=[0, 0, 0])
    book1["b"] = Hist(bin("x", 4, 0, 4), fill=[0, 1, 1])
    book1["d"] = Hist(bin("x", 4, 0, 4), fill=[])
    book2["a"] = Hist(bin("x", 4, 0, 4), fill=[2, 2])
    book2["b"] = Hist(bin("x", 4, 0, 4), fill=[0, 3, 3, 3, 3])
    book2["c"] = Hist(bin("x", 4, 0, 4), fill=[0, 1, 2, 3])
    book = Book.group(x=book1, y=book2)
    assert book["a"].groupkeys("source") == set(["x", "y"])
    assert book["b"].groupkeys("source") == set(["x", "y"])
    assert book["c"].groupkeys("source") == set(["y"])
    assert book["d"].groupkeys("source") == set(["x"])
    assert book["a"]._content["x"].tolist() == [[0], [3], [0], [0], [0], [0], [0]]
    assert book["a"]._content["y"].tolist() == [[0], [0], [0], [2], [0], [0], [0]]
    assert book1["a"]._content.tolist() == [[0], [3], [0], [0], [0], [0], [0]]          # the sources keep their own contents

    # the grouped histogram is an ordinary Hist(groupby("source"), bin(...)): it keeps filling on the device, old and new keys
    g = Hist.group(x=book1["b"], y=book2["b"])
    g.fill(x=[0.5, 1.5, 2.5, 2.5], source=numpy.array(["x", "y", "z", "z"]))
    assert g.groupkeys("source") == set(["x", "y", "z"])
    assert g._content["x"].tolist() == [[0], [2], [2], [0], [0], [0], [0]]
    assert g._content["y"].tolist() == [[0], [1], [1], [0], [4], [0], [0]]
    assert g._content["z"].tolist() == [[0], [0], [0], [2], [0], [0], [0]]
    assert book1["b"]._content.tolist() == [[0], [1], [2], [0], [0], [0], [0]]
    with pytest.raises(ValueError):
        Hist.group(by="source", one=g)                         # groupby("source") already exists


def test_views_and_copy_on_fill():                            # tests/test_book.py:160-173
    everything = ChannelsBook(
        mass=SamplesBook(["data", "signal", "background"],
                         SystematicsBook(Hist(bin("x", 5, 0, 5), systematic=[0]),
                                         Hist(bin("x + epsilon", 5, 0, 5), systematic=[1]),
                                         Hist(bin("x - epsilon", 5, 0, 5), systematic=[-1]))),
        truth=SamplesBook(["signal", "background"],
                          Book(par1=Hist(bin("par1", 5, 0, 5)), par2=Hist(bin("par2", 5, 0, 5)))))
    rs = numpy.random.RandomState(1)
    everything.view("*/data/*").fill(x=rs.uniform(0, 5, 100000), epsilon=rs.normal(0, 0.01, 100000))
    assert everything["mass/data/0/0"].table(recarray=False)[:, 0].sum() == 100000
    assert everything["mass/signal/0/0"].table(recarray=False).sum() == 0

    h = Hist(bin("x", 4, 0, 4), fill=[0.5, 1.5])
    c = h.copyonfill()
    assert c._content.tolist() == h._content.tolist()
    c.fill(x=[2.5])
    assert h._content.tolist() == [[0], [1], [1], [0], [0], [0], [0]]
    assert c._content.tolist() == [[0], [1], [1], [1], [0], [0], [0]]
    d = c.copy()
    d.fill(x=[3.5])
    assert c._content.sum() == 3 and d._content.sum() == 4
    c.clear()
    assert c._content is None


def test_device_columns_pandas_and_read_side():
    import torch
    rs = numpy.random.RandomState(3)
    x, y, w = rs.normal(0, 1, 200000), rs.normal(0, 1, 200000), rs.uniform(0.5, 1.5, 200000)
    h = Hist(bin("x + y", 20, -5, 5), bin("sqrt(x**2 + y**2)", 10, 0, 5), weight="w")
    h.fill(x=torch.from_numpy(x).cuda(), y=torch.from_numpy(y).cuda(), w=torch.from_numpy(w).cuda())
    with histbook_b200.reference_path() as hb:
        r = hb.Hist(hb.bin("x + y", 20, -5, 5), hb.bin("sqrt(x**2 + y**2)", 10, 0, 5), weight="w")
        r.fill(x=x, y=y, w=w)
        expect = r._content
        expect_table = r.project("x + y").table(recarray=False)
    assert numpy.allclose(h._content, expect, rtol=1e-12, atol=0)
    assert numpy.allclose(h.project("x + y").table(recarray=False), expect_table, rtol=1e-12)
    import pandas
    h2 = Hist(bin("x", 10, -3, 3))
    h2.fill(pandas.DataFrame({"x": x}))
    h3 = Hist(bin("x", 10, -3, 3))
    h3.fill(x=x)
    assert numpy.array_equal(h2._content, h3._content) and h2._content.sum() == len(x)
    assert h2.pandas().shape[0] == 13


def test_randomised_book_against_the_reference_path():
    rs = numpy.random.RandomState(99)
    n = 300000
    cols = {"x": rs.normal(0, 1, n), "y": rs.normal(0, 1, n), "z": rs.normal(0, 1, n), "w": rs.uniform(0.5, 1.5, n),
            "n": rs.poisson(3, n).astype(numpy.int64)}
    cols["x"][::1009] = numpy.nan

    def make(hb):
        b = hb.Book()
        b["a"] = hb.Hist(hb.bin("x", 100, -5, 5))
        b["b"] = hb.Hist(hb.bin("x*y", 50, -5, 5), hb.intbin("n", 0, 9), weight="w")
        b["c"] = hb.Hist(hb.split("z", (-2, -1, 0, 1, 2)), hb.cut("x > y"), hb.profile("z"))
        b["d"] = hb.Hist(hb.groupbin("n", 2), hb.bin("sqrt(x**2+y**2+z**2)", 50, 0, 5), filter="n >= 3", weight="w")
        return b
    ours = make(__import__("histbook"))
    ours.fill(**cols)
    ours.fill(**cols)
    with histbook_b200.reference_path() as hb:
        ref = make(hb)
        ref.fill(**cols)
        ref.fill(**cols)
        expect = dict((k, ref[k]._content) for k in "abcd")
    for k in "abc":
        got = ours[k]._content
        both_nan = numpy.isnan(got) & numpy.isnan(expect[k])
        assert numpy.allclose(numpy.where(both_nan, 0, got), numpy.where(both_nan, 0, expect[k]), rtol=1e-11, atol=1e-9), k
    assert numpy.array_equal(ours["a"]._content, expect["a"])
    assert set(ours["d"]._content.keys()) == set(expect["d"].keys())
    for key in expect["d"]:
        assert numpy.allclose(ours["d"]._content[key], expect["d"][key], rtol=1e-11, atol=1e-9)


def test_columns_produced_on_a_side_stream():
    """Device columns written by kernels on a non-blocking torch stream: the fill (engine stream 0) must wait for them,
    and the side stream must wait for the fill before it overwrites the columns again."""
    import torch
    side = torch.cuda.Stream()
    n = 20_000_000
    h = Hist(bin("x", 10, 0, 10))
    total = 0
    with torch.cuda.stream(side):
        x = torch.empty(n, device="cuda", dtype=torch.float64)
        for value in (0.5, 3.5, 7.5):
            big = torch.full((n,), value, device="cuda", dtype=torch.float64)
            for _ in range(6):
                big = big * 1.0 + 0.0                      # keep the side stream busy ahead of the copy below
            x.copy_(big)
            h.fill(x=x)                                    # reads x; the next x.copy_ must not overtake it
            total += n
    c = h._content[:, 0]
    assert c.sum() == total
    assert c[1] == n and c[4] == n and c[8] == n


def test_column_on_another_gpu_is_refused():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    h = Hist(bin("x", 10, 0, 10))
    with pytest.raises(NotImplementedError, match="lives on cuda:1"):
        h.fill(x=torch.ones(100, device="cuda:1", dtype=torch.float64))
    h.fill(x=torch.ones(100, device="cuda:0", dtype=torch.float64))
    assert h._content[:, 0].sum() == 100


def test_cuda_array_interface_stream_key_is_honoured():
    """__cuda_array_interface__ v3: an exporter that names its stream is waited for, whatever torch's current stream is."""
    import torch

    class Exported(object):
        def __init__(self, tensor, stream):
            self.tensor, self.stream = tensor, stream

        @property
        def __cuda_array_interface__(self):
            out = dict(self.tensor.__cuda_array_interface__)
            out["stream"], out["version"] = self.stream, 3
            return out

    side = torch.cuda.Stream()
    n = 20_000_000
    with torch.cuda.stream(side):
        x = torch.full((n,), 2.5, device="cuda", dtype=torch.float64)
        for _ in range(8):
            x = x * 1.0 + 0.0
    h = Hist(bin("x", 10, 0, 10))
    h.fill(x=Exported(x, side.cuda_stream))                # torch's current stream is the default one here
    assert h._content[3, 0] == n
    side.synchronize()


def test_group_axis_histograms_add_on_the_device():           # Hist.__add__/__iadd__ on dict contents, histbook/hist.py:513-537
    """Two device-resident histograms with group axes are added slab slot by slab slot after uniting their keys; the
    result -- values, keys and key order -- is what the reference's recursion gives on host copies."""
    rs = numpy.random.RandomState(3)
    n1 = rs.poisson(3, 30000).astype(numpy.int64)
    n2 = rs.poisson(6, 20000).astype(numpy.int64) + 3             # other keys, some shared
    cols1 = {"x": rs.normal(0, 1, 30000), "n": n1, "w": rs.uniform(0.5, 1.5, 30000)}
    cols2 = {"x": rs.normal(0, 1, 20000), "n": n2, "w": rs.uniform(0.5, 1.5, 20000)}

    def make():
        return Hist(groupbin("n", 2), groupby("n"), bin("x", 20, -3, 3), weight="w")
    a, b = make(), make()
    a.fill(**cols1)
    b.fill(**cols2)
    with histbook_b200.reference_path() as hb:
        ra = hb.Hist(hb.groupbin("n", 2), hb.groupby("n"), hb.bin("x", 20, -3, 3), weight="w")
        rb = hb.Hist(hb.groupbin("n", 2), hb.groupby("n"), hb.bin("x", 20, -3, 3), weight="w")
        ra.fill(**cols1)
        rb.fill(**cols2)
        rc = ra + rb
        expected = rc._content

    def check(got, exp):
        assert list(got) == list(exp)
        for k in exp:
            assert list(got[k]) == list(exp[k]), k
            for k2 in exp[k]:
                assert numpy.all(numpy.abs(got[k][k2] - exp[k][k2]) <= 1e-12 * numpy.maximum(numpy.abs(exp[k][k2]), 1.0))
    c = a + b
    for h in (a, b, c):
        st = h.__dict__["_hb"]
        assert st.dev_valid and not st.host_valid                 # nothing went through the host
    check(histbook_b200.peek(c), expected)
    nkeys_b = len(b.__dict__["_hb"].tuples)
    a += b
    assert a.__dict__["_hb"].dev_valid and not a.__dict__["_hb"].host_valid
    check(histbook_b200.peek(a), expected)
    assert len(b.__dict__["_hb"].tuples) == nkeys_b                # the right operand is untouched
    a.fill(**cols2)                                                # and keeps filling: the united keys are known to the next kernel
    assert abs(sum(v[..., 0].sum() for d in a._content.values() for v in d.values()) - (cols1["w"].sum() + 2 * cols2["w"].sum())) < 1e-6


def test_serialisation_keeps_the_device_copy():               # Hist.tojson / __getstate__, histbook/hist.py:700-741
    """Pickle and JSON read the bins without taking the device copy out of service (a plain ``_content`` read would make
    the host copy the authoritative one and the next fill would upload it again); large contents cross in chunks
    through page-locked staging."""
    rs = numpy.random.RandomState(5)
    h = Hist(bin("x", 300, -4, 4), bin("y", 300, -4, 4), bin("z", 40, -4, 4), weight="w")       # 302 * 302 * 42 * 2 doubles = 61 MB
    cols = {"x": rs.normal(0, 1, 200000), "y": rs.normal(0, 1, 200000), "z": rs.normal(0, 1, 200000), "w": rs.uniform(0.5, 1.5, 200000)}
    h.fill(**cols)
    st = h.__dict__["_hb"]
    blob = pickle.dumps(h)
    assert h.__dict__["_hb"] is st and st.dev_valid and not st.host_valid
    h2 = pickle.loads(blob)
    assert numpy.array_equal(h2._content, histbook_b200.peek(h))
    g = Hist(groupby("c"), bin("x", 10, -3, 3))
    g.fill(c=["a", "b", "a", "c"], x=[0.1, 0.2, 0.3, 0.4])
    j = g.tojson()
    assert g.__dict__["_hb"].dev_valid and not g.__dict__["_hb"].host_valid
    assert sorted(j["content"]) == ["a", "b", "c"] and sum(sum(r[0] for r in v) for v in j["content"].values()) == 4.0
    h.fill(**cols)                                                # no upload in between: still the same arena, twice the weight
    assert h.__dict__["_hb"] is st and abs(h._content[..., 0].sum() - 2 * cols["w"].sum()) < 1e-6


# ------------------------------------------------------------------------------------------------
# the behaviours DESIGN.md section 8 lists as chosen where the reference has none (or a different one): pinned here

def test_chosen_behaviours_where_the_reference_differs():
    # profile axis + masked rows: the reference raises (weight compressed, sumx not; hist.py:429-440); here masked rows are skipped
    h = Hist(bin("x", 2, 0, 1, underflow=False, overflow=False, nanflow=False), profile("y"))
    cols = {"x": numpy.array([-1.0, 0.2, 0.7, 5.0, numpy.nan]), "y": numpy.array([1.0, 2.0, 3.0, 4.0, 5.0])}
    h.fill(**cols)
    assert h._content.tolist() == [[2.0, 4.0, 1.0], [3.0, 9.0, 1.0]]
    with histbook_b200.reference_path() as hb:
        r = hb.Hist(hb.bin("x", 2, 0, 1, underflow=False, overflow=False, nanflow=False), hb.profile("y"))
        with pytest.raises(ValueError):
            r.fill(**cols)
    # groupby on a float key that holds NaN: ONE NaN key however many fills (the reference adds a key per fill: nan != nan)
    g = Hist(groupby("k"), bin("x", 2, 0, 1))
    for _ in range(3):
        g.fill(k=numpy.array([1.5, numpy.nan, numpy.nan]), x=numpy.array([0.1, 0.6, 0.7]))
    keys = list(g._content)
    assert len(keys) == 2 and sum(1 for k in keys if k != k) == 1
    nan_key = [k for k in keys if k != k][0]
    assert g._content[nan_key][:, 0].tolist() == [0.0, 0.0, 6.0, 0.0, 0.0]
    # cut on a non-boolean column: values outside {0, 1} are masked (the reference indexes outside the axis)
    c = Hist(cut("p"))
    c.fill(p=numpy.array([0, 1, 1, 2, -1, 7], dtype=numpy.int64))
    assert c._content.tolist() == [[1.0], [2.0]]
    # Book.fill honours copy-on-fill (the reference's Book.fill writes into the shared content, book.py:582-586)
    a = Hist(bin("x", 2, 0, 1))
    a.fill(x=[0.1])
    b = a.copyonfill()
    book = Book(b=b)
    book.fill(x=[0.7, 0.8])
    assert a._content[:, 0].tolist() == [0.0, 1.0, 0.0, 0.0, 0.0] and b._content[:, 0].tolist() == [0.0, 1.0, 2.0, 0.0, 0.0]
    # one Hist under two names of a Book is filled once
    one = Hist(bin("x", 2, 0, 1))
    twice = Book(first=one, second=one)
    twice.fill(x=[0.2, 0.3])
    assert one._content[:, 0].sum() == 2.0


@pytest.mark.parametrize("expr,lo,hi", [("sin(x)", -1, 1), ("cos(x)", -1, 1), ("tan(x)", -5, 5), ("exp(x)", 0, 20), ("log(abs(x) + 0.001)", -7, 2),
                                        ("log10(abs(x) + 0.001)", -3, 1), ("tanh(x)", -1, 1), ("arctan2(x, y)", -3.2, 3.2), ("hypot(x, y)", 0, 5),
                                        ("pow(abs(x), 2.5)", 0, 10), ("erf(x)", -1, 1), ("sinh(x)", -5, 5), ("arcsin(x / 5)", -1.6, 1.6)])
def test_libm_class_functions_at_ten_million_events(expr, lo, hi):
    """sin, exp, log, pow, hypot, arctan2 ... are CUDA's libm on the device and numpy's SIMD loops in the reference:
    both are accurate to 1-2 ulp but not identical, so an event whose value lies within ~2 ulp of a bin edge may land
    in the neighbouring bin.  With 100 bins that is ~1e-13 of the events: at 1e7 events the histograms must have equal
    totals and differ by at most 2 moved events (4 units of |difference|) -- stated per function, measured here."""
    rs = numpy.random.RandomState(77)
    n = 10**7
    cols = {"x": rs.normal(0, 1.5, n), "y": rs.normal(0, 1.5, n)}
    h = Hist(bin(expr, 100, lo, hi))
    h.fill(**cols)
    with histbook_b200.reference_path() as hb:
        r = hb.Hist(hb.bin(expr, 100, lo, hi))
        r.fill(**cols)
        exp = r._content
    got = h._content
    assert got.sum() == exp.sum() == n
    assert numpy.abs(got - exp).sum() <= 4, (expr, float(numpy.abs(got - exp).sum()))
