"""Deterministic mode (``set_options(deterministic=True)``) on the B200.

Every float64 term is accumulated as an exact 128-bit fixed-point integer (shared memory per block, then global),
and rounded to float64 once per fill.  So a fill must

  * equal, bit for bit, the correctly rounded exact sum of its terms (``fill_oracle.fill(..., exact=True)``,
    math.fsum per bin), which is also within the north star's 1e-12 of the reference's sequential sum;
  * not depend on block shape, grid size, window placement or the order atomics arrive in.
"""
import numpy
import pytest

import golden_util
from oracle import fill_oracle

pytestmark = pytest.mark.gpu


def _columns(names, n, seed, negative_weights=False):
    rs = numpy.random.RandomState(seed)
    out = {}
    for name in names:
        if name == "w":
            out[name] = rs.uniform(-1.0, 1.5, n) if negative_weights else rs.uniform(0.5, 1.5, n)
        elif name == "n":
            out[name] = rs.poisson(3.0, n).astype(numpy.int64)
        else:
            out[name] = rs.normal(0.0, 1.0, n)
    return out


def _run(spec, fills, **options):
    from histbook_b200 import fillable
    saved = dict(fillable._options)
    try:
        fillable.set_options(deterministic=True, fx_stats=True, **options)
        sf = fillable.SpecFill(spec)
        missed = 0
        for arrays in fills:
            sf.fill(arrays)
            missed = fillable.last_stats.get("fx_missed_total", 0)
        assert fillable.last_stats["deterministic"] is True
        return sf.contents(), missed
    finally:
        fillable._options.clear()
        fillable._options.update(saved)
        fillable.set_options()


def _flat(contents):
    out = []
    for c in contents:
        out.extend(a for _, a in golden_util.flatten(c))
    return out


def _assert_identical(a, b):
    fa, fb = _flat(a), _flat(b)
    assert len(fa) == len(fb)
    for x, y in zip(fa, fb):
        assert x.shape == y.shape
        assert numpy.array_equal(x.view(numpy.uint64), y.view(numpy.uint64))


@pytest.mark.parametrize("name,n,neg", [("c2", 300000, False), ("c2", 300000, True), ("c3", 300000, False), ("c5", 200000, False),
                                        ("c4", 20000, False), ("c1", 100000, False)])
def test_equals_the_exactly_rounded_sum(name, n, neg):
    from histbook_b200 import spec as hbspec
    spec = golden_util.load()[name]["spec"]
    fills = [_columns(hbspec.spec_fields(spec), n, 11, neg), _columns(hbspec.spec_fields(spec), n // 3 + 7, 12, neg)]
    got, missed = _run(spec, fills)
    exp = None
    for arrays in fills:
        exp = fill_oracle.fill(spec, arrays, exp, exact=True)
    assert missed == 0
    _assert_identical(got, exp)
    # ... and the exact sum is inside the stated tolerance of the reference's own (sequential) sum
    ref = bound = None
    for arrays in fills:
        ref = fill_oracle.fill(spec, arrays, ref)
        bound = fill_oracle.fill(spec, arrays, bound, absolute=True)
    for g, r, b in zip(_flat(got), _flat(ref), _flat(bound)):
        assert numpy.all(numpy.abs(g - r) <= 1e-12 * b)


def test_independent_of_launch_shape_and_strategy():
    from histbook_b200 import spec as hbspec
    for name, n in (("c2", 400000), ("c3", 250000)):
        spec = golden_util.load()[name]["spec"]
        fills = [_columns(hbspec.spec_fields(spec), n, 21, True)]
        base, _ = _run(spec, fills)
        for options in ({"threads": 256, "min_blocks": 1, "win_threads": 256, "win_min_blocks": 1, "win_smem": 64 * 1024},
                        {"window": False, "strategy": {"*": "global"}},
                        {"smem_budget": 0, "window": True, "win_smem": 100 * 1024, "win_threads": 512, "win_min_blocks": 2}):
            other, _ = _run(spec, fills, **options)
            _assert_identical(base, other)


def test_values_outside_the_exact_window_are_still_added():
    """Weights spanning more than the 128-bit field holds exactly (100 binades below the pilot's maximum): the small ones
    are added as float64 REDs (counted in fx_missed_total), the result stays inside the tolerance."""
    spec = golden_util.load()["c2"]["spec"]
    rs = numpy.random.RandomState(3)
    n = 50000
    arrays = {"x": rs.normal(size=n), "y": rs.normal(size=n), "w": numpy.where(rs.uniform(size=n) < 0.01, 1e-40, 1.0) * rs.uniform(0.5, 1.5, n)}
    got, missed = _run(spec, [arrays])
    assert missed > 0
    ref = fill_oracle.fill(spec, arrays)
    bound = fill_oracle.fill(spec, arrays, absolute=True)
    for g, r, b in zip(_flat(got), _flat(ref), _flat(bound)):
        assert numpy.all(numpy.abs(g - r) <= 1e-12 * b)


def test_fast_fixed_point_option_matches_the_oracle():
    """``fx=True`` without the deterministic merge: exact per block, float64 REDs between blocks."""
    from histbook_b200 import fillable, spec as hbspec
    spec = golden_util.load()["c3"]["spec"]
    arrays = _columns(hbspec.spec_fields(spec), 200000, 31)
    saved = dict(fillable._options)
    try:
        fillable.set_options(fx=True)
        sf = fillable.SpecFill(spec)
        sf.fill(arrays)
        got = sf.contents()
    finally:
        fillable._options.clear()
        fillable._options.update(saved)
        fillable.set_options()
    ref = fill_oracle.fill(spec, arrays)
    bound = fill_oracle.fill(spec, arrays, absolute=True)
    for g, r, b in zip(_flat(got), _flat(ref), _flat(bound)):
        assert numpy.all(numpy.abs(g - r) <= 1e-12 * b)


def test_fixed_point_pilot_follows_the_data():
    """Default options on a small weighted table (all terms fixed point).  The pilot of the first, tiny fill sees weights
    near 1; later fills bring weights a million times larger (outside the field's headroom: they are added as float64
    REDs and counted) -- every fill must still match the oracle, and the pilot is repeated once it has seen enough."""
    from histbook_b200 import fillable
    spec = golden_util.load()["bin_weight"]["spec"] if "bin_weight" in golden_util.load() else None
    if spec is None:
        spec = [c for n, c in golden_util.load().items() if c["spec"]["hists"][0]["weight"] and len(c["spec"]["hists"]) == 1
                and not c["spec"]["hists"][0]["group"] and "w" in c["spec"]["hists"][0]["weight"]][0]["spec"]
    from histbook_b200 import spec as hbspec
    fields = hbspec.spec_fields(spec)
    rs = numpy.random.RandomState(5)

    def cols(n, scale):
        out = {}
        for f in fields:
            out[f] = rs.uniform(0.0, 4.0, n) * (scale if f in ("w", "weight", "y") else 1.0)
        return out
    fills = [cols(3000, 1.0), cols(5000, 1.0e6), cols(300000, 1.0e6), cols(300000, 1.0e6)]
    saved = dict(fillable._options)
    try:
        fillable.set_options(fx_stats=True)
        sf = fillable.SpecFill(spec)
        exp = bound = None
        missed = []
        for arrays in fills:
            sf.fill(arrays)
            missed.append(fillable.last_stats.get("fx_missed_total", 0))
            exp = fill_oracle.fill(spec, arrays, exp)
            bound = fill_oracle.fill(spec, arrays, bound, absolute=True)
            for g, r, b in zip(_flat(sf.contents()), _flat(exp), _flat(bound)):
                assert numpy.all(numpy.abs(g - r) <= 1e-12 * b)
    finally:
        fillable._options.clear()
        fillable._options.update(saved)
        fillable.set_options()
    assert any(v[0].fx_terms for v in sf.plan.kernels.values())
    # fill 2 still runs on the first pilot's scales: its large weights are REDs, counted; fill 3 brings 100x the data the
    # pilot had seen, the pilot is repeated, and from then on everything fits
    assert missed[0] == 0 and missed[1] > 0 and missed[2] == missed[1] and missed[3] == missed[1]
