"""The product's HOST path (fillable.SpecFill -> fill_hists: column binding, key discovery, goal tables, slot maps, choice of the
kernel variant, Store) run end to end on CPU: the engine's device entry points are swapped for the emulator
(tests/emul_engine.py), the generated kernels execute on CPU threads, the results are compared with the golden outputs of the
unmodified reference and with the oracle -- the same comparisons as tests/test_gpu_parity.py makes on the B200.

The inline PTX of the template has host forms in the shim (the carry chains of the fixed-point adds, the vector loads;
tests/emul_util.PTX_FORMS), so hot windows, fixed-point accumulation and the four-events-per-thread path run here too: every
golden case is run twice, with fx / window / box switched off and -- where that changes the kernel -- with the defaults.
The deterministic mode has its own file (tests/test_deterministic_emulated.py).  Not covered (GPU tests only): what lives in the
engine itself (staging, streams, arenas on the device, its utility kernels)."""
import shutil

import numpy
import pytest

import emul_engine
import golden_util
from oracle import fill_oracle
from test_gpu_parity import compare, LIBM_CASES

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="the kernel emulator is compiled with g++")

PLAIN = dict(fx=False, window=False, box=False)


def _run(spec, fills, **options):
    from histbook_b200 import fillable
    saved = dict((k, fillable._options.get(k)) for k in options)
    with emul_engine.installed():
        try:
            fillable.set_options(**options)
            sf = fillable.SpecFill(spec)
            for arrays in fills:
                sf.fill(arrays)
            return sf.contents(), list(emul_engine.LAUNCHES), dict(fillable.last_stats)
        finally:
            fillable.set_options(**saved)


def _emulable_case(name):
    case = golden_util.load()[name]
    if name in LIBM_CASES:
        return False
    for arrays in golden_util.fills(case):
        for v in arrays.values():
            if isinstance(v, numpy.ndarray) and v.dtype.kind not in "fiub":
                return False                            # string keys: dictionary-encoded on the host, fine, but keep this suite small
    return True


def _kernel_text(case, options):
    """the fill kernel these options generate for the case's first fill (None if its columns are not plain arrays)"""
    from histbook_b200 import codegen, columns, spec as hbspec
    try:
        fields = hbspec.spec_fields(case["spec"])
        cols, length = columns.bind_all(fields, golden_util.fills(case)[0])
        coltypes = dict((n, (cols[n].dtype, cols[n].is_scalar)) for n in fields)
        return codegen.generate(case["spec"], coltypes, options).source
    except Exception:
        return None


@pytest.mark.parametrize("mode", ["plain", "default"])
@pytest.mark.parametrize("name", [n for n in golden_util.names() if n not in ("c2", "c3", "c5", "c2_nanlaced")])
def test_golden_cases_through_the_host_path(name, mode):
    """every golden case of the reference (several fills, scalars, group axes ...) through SpecFill on the emulated engine"""
    if not _emulable_case(name):
        pytest.skip("not a case for the emulated engine")
    case = golden_util.load()[name]
    fills = golden_util.fills(case)
    if sum(len(next((v for v in arrays.values() if isinstance(v, numpy.ndarray) and v.ndim == 1), [])) for arrays in fills) > 400000:
        pytest.skip("too many events for CPU threads")
    if mode == "default" and _kernel_text(case, {}) == _kernel_text(case, PLAIN):
        pytest.skip("the defaults generate the same kernel as fx / window / box off")
    got, launches, stats = _run(case["spec"], fills, **(PLAIN if mode == "plain" else {}))
    bound = None
    for arrays in fills:
        bound = fill_oracle.fill(case["spec"], arrays, bound, absolute=True)
    compare(name, case["spec"], got, golden_util.outputs(case), bound)


def _c4_columns(n, seed, nmax=None):
    rs = numpy.random.RandomState(seed)
    cols = {"x": rs.normal(0, 1, n), "y": rs.normal(0, 1, n), "z": rs.normal(0, 1, n), "w": rs.uniform(0.5, 1.5, n),
            "n": rs.poisson(3, n).astype(numpy.int64)}
    if nmax is not None:
        cols["n"] = rs.randint(0, nmax, n).astype(numpy.int64)
    return cols


def test_c4_book_picks_the_group_fast_variant_and_leaves_it_when_it_must():
    """three fills of the north-star book: the first two hold few group keys -> the ``gfast`` kernel (one slot-map load per
    event, no global path for slots beyond the privatised ones); the third brings 30 keys -> more slab slots than are
    privatised -> the generic kernel.  Contents accumulate over the fills like the oracle's."""
    spec = golden_util.load()["c4"]["spec"]
    fills = [_c4_columns(5000, 1), _c4_columns(4000, 2), _c4_columns(4000, 3, nmax=60)]
    got, launches, stats = _run(spec, fills)
    fill_launches = [l for l in launches if l[1] == 1024]
    assert [l[3] for l in fill_launches] == ["group_fast", "group_fast", ""], launches
    assert all(l[0] for l in fill_launches)                                   # the full-tile copy of the event code is there
    assert stats["group_fast"] is False and stats["sorted"]["flags"] == 1     # (the last fill's statistics)
    exp = bound = None
    for arrays in fills:
        exp = fill_oracle.fill(spec, arrays, exp)
        bound = fill_oracle.fill(spec, arrays, bound, absolute=True)
    compare("c4", spec, got, [golden_util.flatten(e) for e in exp], bound)


def test_group_fast_can_be_switched_off():
    spec = golden_util.load()["c4"]["spec"]
    fills = [_c4_columns(3500, 9)]
    got, launches, stats = _run(spec, fills, group_fast=False)
    assert [l[3] for l in launches if l[1] == 1024] == [""]
    exp = fill_oracle.fill(spec, fills[0])
    bound = fill_oracle.fill(spec, fills[0], absolute=True)
    compare("c4", spec, got, [golden_util.flatten(e) for e in exp], bound)


def test_empty_fill_and_refill():
    spec = golden_util.load()["c4"]["spec"]
    empty = dict((k, v[:0]) for k, v in _c4_columns(10, 1).items())
    fills = [empty, _c4_columns(3000, 4), empty]
    got, launches, stats = _run(spec, fills)
    exp = bound = None
    for arrays in fills:
        exp = fill_oracle.fill(spec, arrays, exp)
        bound = fill_oracle.fill(spec, arrays, bound, absolute=True)
    compare("c4", spec, got, [golden_util.flatten(e) for e in exp], bound)


@pytest.mark.parametrize("name", ["c1", "c2", "c3", "c4", "c5"])
def test_baseline_configurations_with_their_default_kernels(name):
    """the five BASELINE configurations as the product runs them: c1 u32 counts, c2 pilot -> ragged hot window (dense variant,
    fixed point inside), c3 fixed-point pilot + fixed-point profile sums, c4 sorted big book, c5 box window that is not placed ->
    paired global REDs; two fills each, NaN / inf laced, through SpecFill on the emulated engine"""
    spec = golden_util.load()[name]["spec"]
    n = 30000
    rs = numpy.random.RandomState(3)
    cols = {"x": rs.normal(0, 1, n), "y": rs.normal(0, 1, n), "z": rs.normal(0, 1, n), "w": rs.uniform(0.5, 1.5, n),
            "n": rs.poisson(3, n).astype(numpy.int64)}
    cols["x"][::997] = numpy.nan
    cols["w"][1::1009] = numpy.nan
    cols["y"][2::1013] = numpy.inf
    from histbook_b200 import spec as hbspec
    fields = hbspec.spec_fields(spec)
    fills = [dict((k, cols[k][:n // 2 + 7]) for k in fields), dict((k, cols[k][n // 2 + 7:]) for k in fields)]
    got, launches, stats = _run(spec, fills)
    if name == "c2":
        assert stats["window"] is not None and stats["window"]["dense"] and stats["window_coverage"] > 0.95
        assert len(launches) == 3                               # pilot through the global path, fixed-point range pilot, windowed fill
    if name == "c3":
        assert len(launches) == 3                               # range pilot + two fills
    if name == "c5":
        assert stats["window"] is None and stats["window_coverage"] < 0.25
    exp = bound = None
    for arrays in fills:
        exp = fill_oracle.fill(spec, arrays, exp)
        bound = fill_oracle.fill(spec, arrays, bound, absolute=True)
    compare(name, spec, got, [golden_util.flatten(e) for e in exp], bound)


def test_hot_window_follows_data_that_drifts(monkeypatch):
    """c2's ragged hot window is placed from the first fills; when the data moves away, the periodic look at the histogram
    (fillable._watch_window) places a new window from the recent fills alone -- or, where a new one would not be better by
    WINDOW_DRIFT, at least reports the share of the recent fills the old one really caught (last_stats "window_coverage")."""
    from histbook_b200 import fillable
    monkeypatch.setattr(fillable, "WINDOW_CHECK_FILLS", 2)
    spec = golden_util.load()["c2"]["spec"]

    def run(drift):
        monkeypatch.setattr(fillable, "WINDOW_DRIFT", drift)
        rs = numpy.random.RandomState(12)

        def batch(n, shift):
            return {"x": rs.normal(shift, 0.3, n), "y": rs.normal(shift, 0.3, n), "w": rs.uniform(0.5, 1.5, n)}
        fills = [batch(12000, 0.0), batch(6000, 0.0)] + [batch(6000, 1.2) for _ in range(5)]
        seen = []
        with emul_engine.installed():
            sf = fillable.SpecFill(spec)
            exp = bound = None
            for arrays in fills:
                sf.fill(arrays)
                st = dict(fillable.last_stats)
                seen.append((st["window"]["bins"] if st["window"] else 0, round(st["window_coverage"], 3)))
                exp = fill_oracle.fill(spec, arrays, exp)
                bound = fill_oracle.fill(spec, arrays, bound, absolute=True)
            compare("c2", spec, sf.contents(), [golden_util.flatten(e) for e in exp], bound)
        return seen
    seen = run(0.05)
    assert seen[0][1] > 0.95 and seen[1] == seen[0]               # placed on the first fill
    assert seen[4][0] != seen[0][0] and seen[-1][1] > 0.95        # after two fills of moved data: a new window, placed from them alone
    seen = run(2.0)                                               # never replace: the look reports what the old window caught
    assert all(b == seen[0][0] for b, c in seen) and seen[-1][1] < 0.05


def test_a_book_beyond_one_kernels_parameter_block_is_filled_in_parts():
    """130 histograms: two kernels (fillable.Plan.parts_for, PART_MAX_HISTS = 96) in one pass over the columns, one key
    discovery for the whole book; the parts run the generic kernels (fillable._fill_parts)"""
    from histbook_b200 import fillable
    base = golden_util.load()["c4"]["spec"]
    hists = list(base["hists"]) * 2 + list(base["hists"][:2])
    spec = dict(base, hists=hists)
    assert len(hists) == 130 > fillable.PART_MAX_HISTS
    fills = [_c4_columns(4000, 31), _c4_columns(3500, 32)]
    got, launches, stats = _run(spec, fills)
    assert stats.get("parts") == 2
    exp = bound = None
    for arrays in fills:
        exp = fill_oracle.fill(spec, arrays, exp)
        bound = fill_oracle.fill(spec, arrays, bound, absolute=True)
    compare("c4 x 2", spec, got, [golden_util.flatten(e) for e in exp], bound)
