"""GPU parity (run on the B200 with -m gpu): the CUDA fill path, called through the C ABI, against

  * the golden outputs of the unmodified reference (tests/golden, every case), and
  * the CPU oracle on fresh seeded inputs.

Bars (BASELINE.json north_star): bin indices and unweighted counts bit-exact; weighted / profile sums
within 1e-12 * (per-bin sum of |term|), which the oracle computes with ``absolute=True``.
"""
import numpy
import pytest

import golden_util
from oracle import fill_oracle

pytestmark = pytest.mark.gpu

RTOL = 1e-12
# functions evaluated by CUDA's libm instead of numpy's: not bit-identical, an event within an ulp of a bin
# edge may land in the neighbouring bin ("hard parts")
LIBM_CASES = {"library_transcendental"}
CASES = golden_util.names()


def run_spec(spec, fills, device=False):
    import torch
    from histbook_b200 import fillable
    sf = fillable.SpecFill(spec)
    for arrays in fills:
        if device:
            arrays = dict((k, torch.from_numpy(numpy.ascontiguousarray(v)).cuda() if isinstance(v, numpy.ndarray) and v.dtype != numpy.bool_ and v.ndim == 1 else v)
                          for k, v in arrays.items())
        sf.fill(arrays)
    return sf.contents()


def compare(name, spec, got, exp, bound):
    assert len(got) == len(exp)
    for hid, (g, e, b) in enumerate(zip(got, exp, bound)):
        fg, fe, fb = golden_util.flatten(g), e, golden_util.flatten(b)
        assert [k for k, _ in fg] == [k for k, _ in fe], (name, hid)
        h = spec["hists"][hid]
        for (k, ga), (_, ea), (_, ba) in zip(fg, fe, fb):
            assert ga.shape == ea.shape and ga.dtype == numpy.float64
            nan_g, nan_e = numpy.isnan(ga), numpy.isnan(ea)
            assert numpy.array_equal(nan_g, nan_e), (name, hid, k)
            ok = ~nan_e
            with numpy.errstate(invalid="ignore"):
                diff = numpy.where(ga[ok] == ea[ok], 0.0, numpy.abs(ga[ok] - ea[ok]))     # equal infinities are equal
            tol = RTOL * numpy.where(numpy.isfinite(ba[ok]), ba[ok], numpy.inf)
            assert numpy.all(diff <= tol), (name, hid, k, float(diff.max()))
            if h["weight"] is None:       # unweighted counts: bit-exact
                assert numpy.array_equal(ga[..., h["sumw"]], ea[..., h["sumw"]]), (name, hid, k)


@pytest.mark.parametrize("name", [n for n in CASES if n not in LIBM_CASES])
def test_golden_host_columns(name):
    case = golden_util.load()[name]
    fills = golden_util.fills(case)
    got = run_spec(case["spec"], fills)
    bound = None
    for arrays in fills:
        bound = fill_oracle.fill(case["spec"], arrays, bound, absolute=True)
    compare(name, case["spec"], got, golden_util.outputs(case), bound)


@pytest.mark.parametrize("name", ["c1", "c2", "c3", "c4", "c5", "c2_nanlaced", "dtype_float32", "dtype_int64", "dtype_int16", "bin_big_1111"])
def test_golden_device_columns(name):
    case = golden_util.load()[name]
    fills = golden_util.fills(case)
    got = run_spec(case["spec"], fills, device=True)
    bound = None
    for arrays in fills:
        bound = fill_oracle.fill(case["spec"], arrays, bound, absolute=True)
    compare(name, case["spec"], got, golden_util.outputs(case), bound)


def test_libm_class_functions_flip_budget():
    """sin/exp/log/... come from CUDA's libm: totals must agree exactly, and at most a handful of
    events may sit in a neighbouring bin."""
    case = golden_util.load()["library_transcendental"]
    got = run_spec(case["spec"], golden_util.fills(case))
    for hid, (g, e) in enumerate(zip(got, golden_util.outputs(case))):
        ea = e[0][1]
        assert g.sum() == ea.sum(), hid
        assert numpy.abs(g - ea).sum() <= 2 * 4, (hid, case["hists"][hid], float(numpy.abs(g - ea).sum()))


@pytest.mark.parametrize("name,n", [("c1", 3000017), ("c2", 2000003), ("c3", 2000003), ("c5", 1500001), ("c4", 300007)])
def test_against_oracle_fresh_inputs(name, n):
    """Sizes the oracle finishes in seconds; unaligned N exercises the tail path; two fills accumulate."""
    case = golden_util.load()[name]
    rs = numpy.random.RandomState(2024)
    cols = {"x": rs.normal(0, 1, n), "y": rs.normal(0, 1, n), "z": rs.normal(0, 1, n),
            "w": rs.uniform(0.5, 1.5, n), "n": rs.poisson(3, n).astype(numpy.int64)}
    cols["x"][::4099] = numpy.nan
    cols["w"][1::5003] = numpy.nan
    cols["y"][2::6007] = numpy.inf
    from histbook_b200 import spec as hbspec
    fields = hbspec.spec_fields(case["spec"])
    half = n // 2 + 1
    fills = [dict((f, cols[f][:half]) for f in fields), dict((f, cols[f][half:]) for f in fields)]
    got = run_spec(case["spec"], fills, device=(name != "c4"))
    exp = bound = None
    for arrays in fills:
        exp = fill_oracle.fill(case["spec"], arrays, exp)
        bound = fill_oracle.fill(case["spec"], arrays, bound, absolute=True)
    compare(name, case["spec"], got, [golden_util.flatten(e) for e in exp], bound)


def test_unaligned_device_pointer_takes_scalar_path():
    import torch
    case = golden_util.load()["c2"]
    rs = numpy.random.RandomState(5)
    n = 100001
    x, y, w = rs.normal(0, 1, n + 1), rs.normal(0, 1, n + 1), rs.uniform(0.5, 1.5, n + 1)
    dx, dy, dw = (torch.from_numpy(a).cuda()[1:] for a in (x, y, w))      # 8-byte, not 32-byte aligned
    from histbook_b200 import fillable
    sf = fillable.SpecFill(case["spec"])
    sf.fill(x=dx, y=dy, w=dw)
    arrays = {"x": x[1:], "y": y[1:], "w": w[1:]}
    exp = fill_oracle.fill(case["spec"], arrays)
    bound = fill_oracle.fill(case["spec"], arrays, absolute=True)
    compare("c2-unaligned", case["spec"], sf.contents(), [golden_util.flatten(e) for e in exp], bound)


def test_empty_and_tiny_fills():
    case = golden_util.load()["c2"]
    from histbook_b200 import fillable
    sf = fillable.SpecFill(case["spec"])
    sf.fill(x=numpy.zeros(0), y=numpy.zeros(0), w=numpy.zeros(0))
    assert sf.contents()[0].shape == (203, 203, 2) and sf.contents()[0].sum() == 0
    sf.fill(x=numpy.array([0.25]), y=numpy.array([0.5]), w=numpy.array([2.0]))
    c = sf.contents()[0]
    assert c.sum() == 6.0 and numpy.count_nonzero(c) == 2


def test_errors_match_reference_messages():
    case = golden_util.load()["c2"]
    from histbook_b200 import fillable
    sf = fillable.SpecFill(case["spec"])
    with pytest.raises(ValueError, match="required field 'w' not found in fill arguments"):
        sf.fill(x=numpy.zeros(3), y=numpy.zeros(3))
    with pytest.raises(ValueError, match="has len . but other arrays have len ."):
        sf.fill(x=numpy.zeros(3), y=numpy.zeros(3), w=numpy.zeros(2))


def test_strategies_agree():
    """Shared-memory privatised and global-RED accumulation give the same histogram."""
    from histbook_b200 import fillable
    case = golden_util.load()["c3"]
    rs = numpy.random.RandomState(11)
    n = 400000
    arrays = {"x": rs.normal(0, 1, n), "y": rs.normal(0, 1, n)}
    out = {}
    try:
        for mode in ("smem", "global"):
            fillable.set_options(strategy={"*": mode})
            sf = fillable.SpecFill(case["spec"])
            sf.fill(arrays)
            out[mode] = sf.contents()[0]
    finally:
        fillable.set_options(strategy=None)
    bound = fill_oracle.fill(case["spec"], arrays, absolute=True)[0]
    assert numpy.all(numpy.abs(out["smem"] - out["global"]) <= 2e-12 * bound)
    assert numpy.array_equal(out["smem"][..., 4], out["global"][..., 4])


def test_dynamic_tile_handout_agrees():
    """Option ``dynamic`` (tiles handed out from a global counter instead of the static grid-stride split): every
    event is still counted exactly once -- counts bit-exact, for whole launches, ragged tails and tiny fills."""
    from histbook_b200 import fillable
    c1, c3 = golden_util.load()["c1"], golden_util.load()["c3"]
    rs = numpy.random.RandomState(5)
    try:
        for mode in ("tail", True):
            fillable.set_options(dynamic=mode)
            for n in (0, 1, 4095, 4097, 1000003, 6000000):
                x, y = rs.normal(0, 1, n), rs.normal(0, 1, n)
                sf = fillable.SpecFill(c1["spec"])
                sf.fill(x=x)
                sf.fill(x=x[: n // 3])
                got = sf.contents()[0]
                if n == 0:
                    assert got is None or not numpy.any(got)
                    continue
                exp = fill_oracle.fill(c1["spec"], {"x": numpy.concatenate([x, x[: n // 3]])})[0]
                assert numpy.array_equal(got, exp), (mode, n)
                sf = fillable.SpecFill(c3["spec"])
                sf.fill(x=x, y=y)
                got = sf.contents()[0]
                if n:
                    exp = fill_oracle.fill(c3["spec"], {"x": x, "y": y})[0]
                    bound = fill_oracle.fill(c3["spec"], {"x": x, "y": y}, absolute=True)[0]
                    assert numpy.array_equal(got[..., 4], exp[..., 4]), (mode, n)
                    assert numpy.all(numpy.abs(got - exp) <= 1e-12 * bound), (mode, n)
    finally:
        fillable.set_options(dynamic=None)


OPTION_SETS = [
    {"fx": False}, {"fx": True}, {"fx": "mix"}, {"window": False}, {"win_dense": True}, {"win_pair": False}, {"merge_pair": False},
    {"win_table32": False}, {"pair_red": False}, {"prefetch": True}, {"replicas": 2}, {"f64_atomic": "exch"}, {"unroll": 1},
    {"unroll": 0}, {"evict_first": False}, {"red_evict_last": True}, {"group_slots": 2}, {"threads": 256, "min_blocks": 4},
    {"smem_budget": 16384}, {"block_times": True}, {"sort": True}, {"sort": True, "sort_lanes": 128}, {"sort": True, "sort_lanes": 32},
    {"sort": True, "sort_ept": 1}, {"sort": True, "sort_sum": "bins", "sort_split": 1}, {"sort": True, "sort_sum": "bins", "sort_split": 16}, {"sort": True, "sort_sum": "bins"}, {"sort": True, "sort_sum": "chunks", "sort_lanes": 32}, {"sort": True, "sort_sum": "bins", "sort_lanes": 256}, {"sort": False},
    {"sort": True, "sort_pipe": True}, {"sort": True, "sort_pipe": False}, {"sort": True, "sort_pipe": True, "sort_ept": 1},
    {"sort": True, "sort_pipe": True, "sort_ept": 4}, {"sort": True, "sort_pipe": True, "sort_threads": 512},
    {"sort": True, "sort_ept": 3}, {"sort": True, "sort_alias_tails": False}, {"sort": True, "sort_ept": 3, "sort_lanes": 128},
]


@pytest.mark.parametrize("options", OPTION_SETS, ids=lambda o: ",".join("%s=%s" % kv for kv in sorted(o.items())))
def test_tuning_options_do_not_change_the_histogram(options):
    """Every kernel-generation option of the sweeps (profiles/r01_sweeps.txt) is a different way to accumulate the same
    sums: counts bit-exact and float sums within the bound against the oracle, for a hot-window histogram (c2), a
    privatised profile (c3), a large global one (c5) and a group-axis Book member (c4's first 8 histograms)."""
    from histbook_b200 import fillable
    rs = numpy.random.RandomState(99)
    n = 1200007                                                           # > the 1 M-event window pilot, ragged tail
    cols = {"x": rs.normal(0, 1, n), "y": rs.normal(0, 1, n), "z": rs.normal(0, 1, n),
            "w": rs.uniform(0.5, 1.5, n), "n": rs.poisson(3, n).astype(numpy.int64)}
    cases = golden_util.load()
    specs = [("c2", cases["c2"]["spec"]), ("c3", cases["c3"]["spec"]), ("c5", cases["c5"]["spec"]),
             ("c4[:8]", dict(cases["c4"]["spec"], hists=cases["c4"]["spec"]["hists"][:8]))]
    from histbook_b200 import spec as hbspec
    try:
        fillable.set_options(**options)
        for name, spec in specs:
            fields = hbspec.spec_fields(spec)
            fills = [dict((f, cols[f]) for f in fields), dict((f, cols[f][: n // 3]) for f in fields)]
            got = run_spec(spec, fills, device=True)
            exp = bound = None
            for arrays in fills:
                exp = fill_oracle.fill(spec, arrays, exp)
                bound = fill_oracle.fill(spec, arrays, bound, absolute=True)
            compare("%s %r" % (name, options), spec, got, [golden_util.flatten(e) for e in exp], bound)
    finally:
        fillable.set_options(**dict((k, None) for k in options))


def _sortable(name):
    from histbook_b200 import codegen, spec as hbspec
    case = golden_util.load()[name]
    fills = golden_util.fills(case)
    try:
        coltypes = dict((f, (numpy.asarray(fills[0][f]).dtype, numpy.ndim(fills[0][f]) == 0)) for f in hbspec.spec_fields(case["spec"]) if f in fills[0])
        return codegen.generate(case["spec"], coltypes, {"sort": True}).sorted is not None
    except Exception:
        return False


@pytest.mark.parametrize("pipe", [False, True], ids=["phases", "pipe"])
@pytest.mark.parametrize("name", [n for n in CASES if n not in LIBM_CASES])
def test_golden_sorted_strategy(name, pipe):
    """Every golden case that has float64 terms on fixed axes, with the sorted strategy of big books forced on
    (per-tile counting sort by bin index + chunked sums instead of per-event atomics): same bars as everywhere.
    Both kernels: phases behind block barriers, and the producer / consumer pipeline (``sort_pipe``)."""
    from histbook_b200 import fillable
    if not _sortable(name):
        pytest.skip("no float64 term on a fixed-axis histogram: nothing to sort")
    case = golden_util.load()[name]
    fills = golden_util.fills(case)
    try:
        fillable.set_options(sort=True, sort_pipe=pipe)
        got = run_spec(case["spec"], fills)
        assert fillable.last_stats.get("sorted"), name
        assert pipe or not fillable.last_stats["sorted"].get("pipe"), name      # (a pipeline that does not fit falls back to phases)
    finally:
        fillable.set_options(sort=None, sort_pipe=None)
    bound = None
    for arrays in fills:
        bound = fill_oracle.fill(case["spec"], arrays, bound, absolute=True)
    compare(name, case["spec"], got, golden_util.outputs(case), bound)


@pytest.mark.parametrize("lanes,summing", [(32, "bins"), (128, "bins"), (64, "chunks"), (128, "chunks"), (32, "pipe")])
def test_sorted_strategy_on_degenerate_distributions(lanes, summing):
    """Every event in ONE bin, two alternating bins, masked rows, NaN weights, tiles with a single live event, more
    bins than lanes: the counting sort and the bin ownership of codegen.SortGroup."""
    from histbook_b200 import fillable
    cases = golden_util.load()
    spec = dict(cases["c4"]["spec"], hists=cases["c4"]["spec"]["hists"][:8])
    rs = numpy.random.RandomState(3)
    try:
        if summing == "pipe":
            fillable.set_options(sort=True, sort_pipe=True)
        else:
            fillable.set_options(sort=True, sort_lanes=lanes, sort_sum=summing, sort_pipe=False)
        for n in (1, 31, 1536, 1537, 2048, 2049, 3072, 3073, 4096 + 64, 303104 + 17, 454656 + 1, 700001):
            variants = []
            const = {"x": numpy.full(n, 0.25), "y": numpy.full(n, -0.75), "z": numpy.full(n, 1.5), "w": numpy.full(n, 1.25), "n": numpy.full(n, 4, dtype=numpy.int64)}
            variants.append(const)
            two = dict(const, x=numpy.where(numpy.arange(n) % 64 < 32, 0.25, 3.05))            # chunk-aligned alternation
            variants.append(two)
            nanw = dict(const, w=numpy.where(numpy.arange(n) % 3 == 0, numpy.nan, 0.5), x=rs.normal(0, 3, n))
            nanw["x"][::7] = numpy.nan
            variants.append(nanw)
            for arrays in variants:
                sf = fillable.SpecFill(spec)
                sf.fill(arrays)
                sf.fill(dict((k, v[: n // 2]) for k, v in arrays.items()))
                got = sf.contents()
                exp = bound = None
                for a in (arrays, dict((k, v[: n // 2]) for k, v in arrays.items())):
                    exp = fill_oracle.fill(spec, a, exp)
                    bound = fill_oracle.fill(spec, a, bound, absolute=True)
                compare("degenerate n=%d" % n, spec, got, [golden_util.flatten(e) for e in exp], bound)
    finally:
        fillable.set_options(sort=None, sort_lanes=None, sort_sum=None, sort_pipe=None)


def test_book_of_320_histograms_is_split_into_parts():
    """More histograms than one kernel's parameter block holds (the reference has no limit, book.py:540-586): the book
    is filled as several kernels in one pass over the columns (Plan.parts_for / hb_fill_multi).  Against the oracle."""
    from histbook_b200 import fillable
    base = golden_util.load()["c4"]["spec"]
    spec = dict(base, hists=base["hists"] * 5)
    rs = numpy.random.RandomState(17)
    n = 200003
    cols = {"x": rs.normal(0, 1, n), "y": rs.normal(0, 1, n), "z": rs.normal(0, 1, n),
            "w": rs.uniform(0.5, 1.5, n), "n": rs.poisson(3, n).astype(numpy.int64)}
    sf = fillable.SpecFill(spec)
    sf.fill(cols)
    assert fillable.last_stats["parts"] == 4
    got = sf.contents()
    exp = fill_oracle.fill(base, cols)
    bound = fill_oracle.fill(base, cols, absolute=True)
    compare("320 hists", spec, got, [golden_util.flatten(e) for e in exp] * 5, bound * 5)


def test_half_books_agree_with_the_single_kernel():
    """parts=2 fills c4 as two half-books (512-thread blocks, two per SM); the default is the single 1024-thread kernel: same bins."""
    from histbook_b200 import fillable
    spec = golden_util.load()["c4"]["spec"]
    rs = numpy.random.RandomState(23)
    n = 700001
    cols = {"x": rs.normal(0, 1, n), "y": rs.normal(0, 1, n), "z": rs.normal(0, 1, n),
            "w": rs.uniform(0.5, 1.5, n), "n": rs.poisson(3, n).astype(numpy.int64)}
    out = {}
    try:
        for parts in (2, 1):
            fillable.set_options(parts=parts)
            sf = fillable.SpecFill(spec)
            sf.fill(cols)
            sf.fill(dict((k, v[:1000]) for k, v in cols.items()))
            assert fillable.last_stats.get("parts", 1) == parts
            out[parts] = sf.contents()
    finally:
        fillable.set_options(parts=None)
    bound = fill_oracle.fill(spec, cols, absolute=True)
    for hid, (a, b, bd) in enumerate(zip(out[2], out[1], bound)):
        fa, fb, fbd = golden_util.flatten(a), golden_util.flatten(b), golden_util.flatten(bd)
        assert [k for k, _ in fa] == [k for k, _ in fb]
        for (k, x), (_, y), (_, z) in zip(fa, fb, fbd):
            assert numpy.all(numpy.abs(x - y) <= 2e-12 * z), (hid, k)
            if spec["hists"][hid]["weight"] is None:
                s = spec["hists"][hid]["sumw"]
                assert numpy.array_equal(x[..., s], y[..., s])


def test_box_window_on_a_large_histogram():
    """c5's 134 MB histogram with data narrow enough for a box window to pay (sigma = 6 bins: a 24^3 box holds ~85 %):
    the events inside go to shared memory, the rest to paired REDs; same bars against the oracle, window reported."""
    from histbook_b200 import fillable
    spec = golden_util.load()["c5"]["spec"]
    rs = numpy.random.RandomState(41)
    n = 1500001
    cols = {"x": rs.normal(0.3, 0.3, n), "y": rs.normal(-1.0, 0.3, n), "z": rs.normal(0, 0.3, n), "w": rs.uniform(0.5, 1.5, n)}
    cols["x"][::5003] = numpy.nan
    cols["y"][1::4001] = 7.5
    sf = fillable.SpecFill(spec)
    fills = [dict((k, torch_or_numpy(v[: n // 2])) for k, v in cols.items()), dict((k, torch_or_numpy(v[n // 2:])) for k, v in cols.items())]
    for f in fills:
        sf.fill(f)
    win = fillable.last_stats.get("window")
    assert win and win["bins"] > 8000 and fillable.last_stats["window_coverage"] > 0.6, fillable.last_stats
    exp = bound = None
    for a in (dict((k, v[: n // 2]) for k, v in cols.items()), dict((k, v[n // 2:]) for k, v in cols.items())):
        exp = fill_oracle.fill(spec, a, exp)
        bound = fill_oracle.fill(spec, a, bound, absolute=True)
    compare("c5 box", spec, sf.contents(), [golden_util.flatten(e) for e in exp], bound)


def torch_or_numpy(v):
    import torch
    return torch.from_numpy(numpy.ascontiguousarray(v)).cuda()


def test_file_backed_columns_stream_through_the_bounce_ring(tmp_path):
    """Columns that live in files (numpy.memmap: pageable, faulted in from disk as they are read) are filled like any
    other host array -- in chunks through the page-locked bounce ring, never materialised as a whole."""
    from histbook_b200 import engine, fillable
    spec = golden_util.load()["c2"]["spec"]
    rs = numpy.random.RandomState(31)
    n = 3000017
    data = {"x": rs.normal(0, 1, n), "y": rs.normal(0, 1, n), "w": rs.uniform(0.5, 1.5, n)}
    maps = {}
    for k, v in data.items():
        path = str(tmp_path / (k + ".f64"))
        v.tofile(path)
        maps[k] = numpy.memmap(path, dtype=numpy.float64, mode="r", shape=(n,))
    engine.lib().hb_set_chunk_events(1 << 19)                  # several chunks: 3 ring slots and 2 device buffers all get reused
    try:
        sf = fillable.SpecFill(spec)
        sf.fill(maps)
        sf.fill(dict((k, v[: n // 3]) for k, v in maps.items()))
    finally:
        engine.lib().hb_set_chunk_events(1 << 24)
    exp = bound = None
    for a in (data, dict((k, v[: n // 3]) for k, v in data.items())):
        exp = fill_oracle.fill(spec, a, exp)
        bound = fill_oracle.fill(spec, a, bound, absolute=True)
    compare("c2 memmap", spec, sf.contents(), [golden_util.flatten(e) for e in exp], bound)


def test_back_to_back_host_fills_do_not_race_on_the_staging_buffers():
    """Two fills from DIFFERENT host arrays with a slow book and no read in between (the advisor's round-1 finding: the
    second call's first copy could overwrite a staging buffer the first call's last kernel was still reading)."""
    from histbook_b200 import fillable
    spec = golden_util.load()["c4"]["spec"]
    rs = numpy.random.RandomState(37)
    n = 600011

    def cols(seed_shift):
        r = numpy.random.RandomState(37 + seed_shift)
        return {"x": r.normal(0, 1, n), "y": r.normal(0, 1, n), "z": r.normal(0, 1, n), "w": r.uniform(0.5, 1.5, n),
                "n": r.poisson(3, n).astype(numpy.int64)}
    a, b, c = cols(0), cols(1), cols(2)
    sf = fillable.SpecFill(spec)
    sf.fill(a)
    sf.fill(b)
    sf.fill(c)                                                # nothing is read until here
    exp = bound = None
    for arrays in (a, b, c):
        exp = fill_oracle.fill(spec, arrays, exp)
        bound = fill_oracle.fill(spec, arrays, bound, absolute=True)
    compare("c4 back to back", spec, sf.contents(), [golden_util.flatten(e) for e in exp], bound)
