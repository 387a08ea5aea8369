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
    """All 64 histograms of the north-star book at 1e9 events against plain torch: every count bit-exact (bin, split,
    bin x intbin, bin x cut, groupbin x bin), every weighted, filtered and profile column within 1e-12 of the per-bin sum
    of |term|.  The eight axis expressions are evaluated as histbook's normal form orders them."""
    torch = _torch()
    n = 10**9
    spec = golden_util.load()["c4"]["spec"]
    assert len(spec["hists"]) == 64
    cols = _columns(["n", "w", "x", "y", "z"], n, 12345)
    got = _fill(spec, cols)
    x, y, z, w, nn = cols["x"], cols["y"], cols["z"], cols["w"], cols["n"]
    edges = torch.tensor([-2, -1, -0.5, 0, 0.5, 1, 2], dtype=torch.float64, device="cuda")
    nbin = (nn + 1).clamp_(0, 11)                            # intbin("n", 0, 9): [under][0..9][over]
    zpos = (z > 0).to(torch.int64)
    gkey = torch.floor(nn.to(torch.float64) * 0.5) * 2.0     # groupbin("n", 2)
    keys = sorted(float(k) for k in torch.unique(gkey).cpu().numpy())
    fw = torch.where(nn >= 3, w, torch.zeros_like(w))        # where(n >= 3, 1, 0) * w
    exprs = [lambda: x, lambda: y, lambda: z, lambda: x + y, lambda: x - y, lambda: x * y,
             lambda: torch.sqrt(x * x + y * y), lambda: torch.sqrt(x * x + (y * y + z * z))]
    for e, make in enumerate(exprs):
        v = make()
        i100, i50 = _bin(v, 100, -5, 5), _bin(v, 50, -5, 5)
        h = got[8 * e: 8 * e + 8]
        # 0: Hist(bin(e, 100, -5, 5))
        c100 = torch.bincount(i100, minlength=103).cpu().numpy().astype(numpy.float64)
        assert numpy.array_equal(h[0][:, 0], c100) and h[0].sum() == n, e
        # 1: weight="w"
        _check_sum(h[1][:, 0], i100, w, 103)
        _check_sum(h[1][:, 1], i100, w * w, 103)
        # 2: split(e, (-2, -1, -0.5, 0, 0.5, 1, 2)): [under][6 finite][over][nan]
        isplit = torch.bucketize(v, edges, right=True)
        assert numpy.array_equal(h[2][:, 0], torch.bincount(isplit, minlength=9).cpu().numpy().astype(numpy.float64)), e
        del isplit
        # 3: bin(e, 50) x intbin("n", 0, 9)
        assert numpy.array_equal(h[3].reshape(-1), torch.bincount(i50 * 12 + nbin, minlength=53 * 12).cpu().numpy().astype(numpy.float64)), e
        # 4: bin(e, 50) x cut("z > 0")
        assert numpy.array_equal(h[4].reshape(-1), torch.bincount(i50 * 2 + zpos, minlength=106).cpu().numpy().astype(numpy.float64)), e
        # 5: bin(e, 100), profile("z"): [sum z, sum z^2, count]
        _check_sum(h[5][:, 0], i100, z, 103)
        _check_sum(h[5][:, 1], i100, z * z, 103)
        assert numpy.array_equal(h[5][:, 2], c100), e
        # 6: groupbin("n", 2) x bin(e, 50): one block per key, keys as the reference spells them
        assert sorted(float(k) for k in h[6]) == keys, e
        for k, block in h[6].items():
            sel = gkey == float(k)
            assert numpy.array_equal(block[:, 0], torch.bincount(i50[sel], minlength=53).cpu().numpy().astype(numpy.float64)), (e, k)
            del sel
        # 7: bin(e, 100), filter="n >= 3", weight="w"
        _check_sum(h[7][:, 0], i100, fw, 103)
        _check_sum(h[7][:, 1], i100, fw * fw, 103)
        del v, i100, i50
        torch.cuda.empty_cache()
