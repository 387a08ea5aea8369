"""The GENERATED kernels of big books, executed on CPU threads (tests/emul_util.py, tests/emul/cuda_emul.h) and compared
with the oracle: counts bit-exact, float sums within 1e-12 * sum|term| -- the same bar as the GPU parity tests.  This is
how a change of the code generator (tickets, scatter, chunked sums, tails, flag bits, group-slot variants ...) is checked
where there is no GPU; the `-m gpu` tests remain the parity tests proper (they run the real thing through the C ABI)."""
import numpy
import pytest

import emul_util
import golden_util
from oracle import fill_oracle


def _columns(n, seed, laced=False):
    rs = numpy.random.RandomState(seed)
    arrays = {"x": rs.normal(0, 1, n), "y": rs.normal(0, 1, n), "z": rs.normal(0, 1, n), "w": rs.uniform(0.5, 1.5, n),
              "n": rs.poisson(3, n).astype(numpy.int64)}
    if laced and n:
        # NaN / inf / huge values in every float column (the int32 cast quirk, NaN weights zeroed, NaN bins)
        for k, name in enumerate(("x", "y", "z", "w")):
            idx = rs.randint(0, n, size=max(1, n // 50))
            arrays[name][idx] = rs.choice([numpy.nan, numpy.inf, -numpy.inf, 1e300, -1e300, 0.0, -0.0], size=len(idx))
        arrays["n"][rs.randint(0, n, size=max(1, n // 100))] = rs.choice([-5, 40, 2**32 + 3], size=max(1, n // 100))
    return arrays


def _check(spec, arrays, options, grid=2, null_follower_maps=True):
    mine = dict((k, arrays[k]) for k in emul_util.hbspec.spec_fields(spec))
    exp = fill_oracle.fill(spec, mine)
    bound = fill_oracle.fill(spec, mine, absolute=True)
    kern, got = emul_util.run(spec, mine, exp, options, grid=grid, null_follower_maps=null_follower_maps)
    for hid, (g, e, b) in enumerate(zip(got, exp, bound)):
        blocks = [(g, e, b)] if isinstance(e, numpy.ndarray) else [(g[k], e[k], b[k]) for k in e]
        for gg, ee, bb in blocks:
            assert gg.shape == ee.shape
            with numpy.errstate(invalid="ignore"):
                close = numpy.abs(gg - ee) <= 1e-12 * bb
            same = (gg == ee) | (numpy.isnan(gg) & numpy.isnan(ee))
            assert numpy.all(close | same), "histogram %d differs from the oracle (options %r)" % (hid, options)
            if spec["hists"][hid]["weight"] is None:
                s = spec["hists"][hid]["sumw"]
                assert numpy.array_equal(gg[..., s], ee[..., s]), "counts of histogram %d are not bit-exact (options %r)" % (hid, options)
    return kern


C4_VARIANTS = [
    {},                                             # the defaults: full-tile copy of the event code, flag bits for filtered weights
    {"group_fast": True},                           # the variant fillable picks when the histograms share their slot maps
    {"sort_gate": False},
    {"sort_full": False, "group_fast": True},
    {"sort_ept": 2, "group_fast": True},            # tiles of 2048 events
    {"sort_lanes": 128},
    {"sort_alias_tails": False},
    {"sort_sum": "bins"},
    {"sort_branchfree": False},
]


@pytest.mark.parametrize("options", C4_VARIANTS, ids=lambda o: ",".join("%s=%s" % kv for kv in sorted(o.items())) or "default")
def test_c4_book_on_the_emulator(options):
    """the north-star book: 64 histograms, 8 sort groups, group axes; two blocks, full tiles and a ragged last one"""
    spec = golden_util.load()["c4"]["spec"]
    kern = _check(spec, _columns(3 * 3072 + 1234, 11), options)
    assert kern.sorted is not None and len(kern.sorted["groups"]) == 8


def test_c4_book_with_nan_inf_and_wrapping_integers():
    spec = golden_util.load()["c4"]["spec"]
    for options in ({"group_fast": True}, {"sort_gate": False, "sort_full": False}):
        _check(spec, _columns(7001, 5, laced=True), options)


def test_c4_book_small_and_degenerate_fills():
    spec = golden_util.load()["c4"]["spec"]
    for n in (1, 31, 3072, 3073):
        _check(spec, _columns(n, 100 + n), {"group_fast": True}, grid=1)
    # every event in one bin of every axis; all weights rejected by the filter; every weight NaN
    arrays = _columns(5000, 3)
    for name in ("x", "y", "z"):
        arrays[name][:] = 0.25
    _check(spec, arrays, {"group_fast": True})
    arrays = _columns(5000, 4)
    arrays["n"][:] = 1
    _check(spec, arrays, {"group_fast": True})
    arrays = _columns(5000, 6)
    arrays["w"][:] = numpy.nan
    _check(spec, arrays, {"group_fast": True})


def test_followers_with_their_own_slot_maps():
    """the generic kernel with every histogram's own map (what fillable passes when the maps differ)"""
    spec = golden_util.load()["c4"]["spec"]
    _check(spec, _columns(4000, 8), {}, null_follower_maps=False)
