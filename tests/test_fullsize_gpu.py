"""The five BASELINE.json configurations at their FULL sizes (1e7 / 1e8 / 1e8 / 1e9 / 1e9 events) on the B200.

The CPU oracle cannot fill 1e9 events in a test, so at these sizes the device fill is checked against an independent
plain-PyTorch float64 formulation of the same operations (index arithmetic with separate, uncontracted torch kernels and
``torch.bincount`` as the scatter), plus size-independent properties of a histogram fill:

  * unweighted counts and bin indices: bit-exact (bincount of the same indices; every event lands in exactly one bin
    when all flow bins are on, so the total is the number of events),
  * weighted / profile sums: |device - torch| <= 1e-12 * (per-bin sum of |term|)  (north star tolerance),
  * linearity: filling two halves one after the other equals filling everything at once (counts exactly).

Columns are generated on the device with a seeded torch generator.
"""
import numpy
import pytest

import golden_util

pytestmark = pytest.mark.gpu

RTOL = 1e-12


def _torch():
    import torch
    return torch


def _columns(names, n, seed):
    torch = _torch()
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    cols = {}
    for name in names:
        if name == "w":
            cols[name] = torch.rand(n, generator=g, device="cuda", dtype=torch.float64) + 0.5
        elif name == "n":
            cols[name] = torch.poisson(torch.full((n,), 3.0, device="cuda", dtype=torch.float32), generator=g).to(torch.int64)
        else:
            cols[name] = torch.randn(n, generator=g, device="cuda", dtype=torch.float64)
    return cols


def _bin(v, numbins, low, high):
    """histbook.binUONL (calc/__init__.py:203-239) for finite input, as separate float64 torch kernels."""
    torch = _torch()
    t = torch.floor((v - float(low)) * (float(numbins) / float(high - low))) + 1.0
    return t.clamp_(0.0, float(numbins + 1)).to(torch.int64)


def _fill(spec, cols, chunks=1):
    from histbook_b200 import fillable
    sf = fillable.SpecFill(spec)
    n = len(next(iter(cols.values())))
    edges = [n * k // chunks // 4096 * 4096 if 0 < k < chunks else (0 if k == 0 else n) for k in range(chunks + 1)]
    for a, b in zip(edges[:-1], edges[1:]):
        sf.fill(dict((k, v[a:b]) for k, v in cols.items()))
    return sf.contents()


def _check_sum(got, idx, term, nbins):
    """got[bin] against bincount(idx, term) with the tolerance relative to bincount(idx, |term|)."""
    torch = _torch()
    exp = torch.bincount(idx, weights=term, minlength=nbins).cpu().numpy()
    bound = torch.bincount(idx, weights=term.abs(), minlength=nbins).cpu().numpy()
    assert got.shape == exp.shape
    assert numpy.all(numpy.abs(got - exp) <= RTOL * bound), float(numpy.max(numpy.abs(got - exp) / numpy.maximum(bound, 1e-300)))


def teardown_function(function):
    _torch().cuda.empty_cache()


def test_c1_full_size_counts_exact():
    torch = _torch()
    n = 10**7
    spec = golden_util.load()["c1"]["spec"]
    cols = _columns(["x"], n, 12345)
    got = _fill(spec, cols)[0]
    exp = torch.bincount(_bin(cols["x"], 100, -5, 5), minlength=103).cpu().numpy()
    assert got.shape == (103, 1) and got[:, 0].sum() == n
    assert numpy.array_equal(got[:, 0], exp.astype(numpy.float64))
    # linearity: three unequal chunks accumulate to the same counts
    assert numpy.array_equal(_fill(spec, cols, chunks=3)[0], got)


def test_c2_full_size():
    torch = _torch()
    n = 10**8
    spec = golden_util.load()["c2"]["spec"]
    cols = _columns(["w", "x", "y"], n, 12345)
    got = _fill(spec, cols)[0]
    x, y, w = cols["x"], cols["y"], cols["w"]
    idx = _bin(x + y, 200, -5, 5) * 203 + _bin(torch.sqrt(x * x + y * y), 200, 0, 5)
    flat = got.reshape(-1, 2)
    _check_sum(flat[:, 0], idx, w, 203 * 203)
    _check_sum(flat[:, 1], idx, w * w, 203 * 203)
    assert abs(flat[:, 0].sum() - float(w.sum())) <= 1e-12 * float(w.sum()) * 10
    # which bins are populated at all is index arithmetic: exact
    populated = torch.bincount(idx, minlength=203 * 203).cpu().numpy() > 0
    assert numpy.array_equal(flat[:, 0] != 0.0, populated)
    del idx
    two = _fill(spec, cols, chunks=2)[0].reshape(-1, 2)
    assert numpy.all(numpy.abs(two - flat) <= RTOL * numpy.abs(flat))


def test_c3_full_size():
    torch = _torch()
    n = 10**8
    spec = golden_util.load()["c3"]["spec"]
    assert spec["hists"][0]["shape"] == [1003, 2, 5]
    cols = _columns(["x", "y"], n, 12345)
    got = _fill(spec, cols)[0].reshape(-1, 5)
    x, y = cols["x"], cols["y"]
    idx = _bin(x, 1000, -5, 5) * 2 + (x > 0).to(torch.int64)
    y2 = y * y
    counts = torch.bincount(idx, minlength=2006).cpu().numpy().astype(numpy.float64)
    assert numpy.array_equal(got[:, 4], counts) and got[:, 4].sum() == n
    _check_sum(got[:, 0], idx, y, 2006)
    _check_sum(got[:, 1], idx, y2, 2006)
    _check_sum(got[:, 2], idx, y2, 2006)
    _check_sum(got[:, 3], idx, y2 * y2, 2006)
    # one accumulated term feeds both columns; the block partials reach them in different orders (fast mode)
    assert numpy.all(numpy.abs(got[:, 1] - got[:, 2]) <= RTOL * numpy.abs(got[:, 1]))


def test_c5_full_size():
    torch = _torch()
    n = 10**9
    spec = golden_util.load()["c5"]["spec"]
    cols = _columns(["w", "x", "y", "z"], n, 12345)
    got = _fill(spec, cols)[0].reshape(-1, 2)
    idx = _bin(cols["x"], 200, -5, 5)
    idx.mul_(203).add_(_bin(cols["y"], 200, -5, 5))
    idx.mul_(203).add_(_bin(cols["z"], 200, -5, 5))
    w = cols["w"]
    _check_sum(got[:, 0], idx, w, 203 ** 3)
    populated = torch.bincount(idx, minlength=203 ** 3).cpu().numpy() > 0
    assert numpy.array_equal(got[:, 0] != 0.0, populated)
    w2 = w * w
    _check_sum(got[:, 1], idx, w2, 203 ** 3)
    total = float(w.sum())
    assert abs(got[:, 0].sum() - total) <= 1e-11 * total


def test_c4_full_size_book():
    torch = _torch()
    n = 10**9
    spec = golden_util.load()["c4"]["spec"]
    assert len(spec["hists"]) == 64
    cols = _columns(["n", "w", "x", "y", "z"], n, 12345)
    got = _fill(spec, cols)
    checked = 0
    for h, content in zip(spec["hists"], got):
        if h["weight"] is not None or h["profile"]:
            continue
        if h["group"]:
            # groupbin("n", 2) x bin: every event has exactly one key and one bin
            assert isinstance(content, dict) and len(content) >= 5
            assert sum(float(v.sum()) for v in content.values()) == n
            checked += 1
            continue
        # unweighted, all flow bins on: the counts add up to the number of events, exactly
        assert float(content.sum()) == n
        checked += 1
        fixed = h["fixed"]
        if len(fixed) == 1 and fixed[0]["kind"] == "bin" and fixed[0]["goal"][2][0] == "name":
            name = fixed[0]["goal"][2][1]
            numbins, low, high = (fixed[0]["goal"][k][1] for k in (3, 4, 5))
            exp = torch.bincount(_bin(cols[name], numbins, low, high), minlength=numbins + 3).cpu().numpy()
            assert numpy.array_equal(content[:, 0], exp.astype(numpy.float64))
    assert checked >= 30
    # a weighted member of the book against torch: Hist(bin("x",100,-5,5), weight="w")
    h1 = spec["hists"][1]
    assert h1["weight"] is not None and h1["fixed"][0]["goal"][2] == ["name", "x"]
    idx = _bin(cols["x"], 100, -5, 5)
    _check_sum(got[1][:, 0], idx, cols["w"], 103)
    _check_sum(got[1][:, 1], idx, cols["w"] * cols["w"], 103)
