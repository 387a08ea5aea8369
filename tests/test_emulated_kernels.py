"""The GENERATED kernels of big books, executed on CPU threads (tests/emul_util.py, tests/emul/cuda_emul.h) and compared
with the oracle: counts bit-exact, float sums within 1e-12 * sum|term| -- the same bar as the GPU parity tests.  This is
how a change of the code generator (tickets, scatter, chunked sums, tails, flag bits, group-slot variants ...) is checked
where there is no GPU; the `-m gpu` tests remain the parity tests proper (they run the real thing through the C ABI)."""
import shutil

import numpy
import pytest

import emul_util
import golden_util
from oracle import fill_oracle

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="the kernel emulator is compiled with g++")


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
    {"sort_lanes": 64},
    {"split_table": True, "group_fast": True},
    {"sort_pair": True, "group_fast": True},        # two events per round of the summing loop
    {"sort_pair": True, "sort_prefetch": True, "sort_threads": 768, "sort_ept": 4},
    {"sort_chunk_odd": False},                      # chunks of tile / lanes event numbers with padding (and its division) instead of 50
    {"sort_chunk_odd": False, "sort_full": False, "split_table": False, "sort_gate": False},      # the kernel as it was before these options
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


def _bin_axis(field, numbins, low, high, flags="UONL"):
    return {"goal": ["call", "histbook.bin" + flags, ["name", field], ["const", numbins], ["const", float(low)], ["const", float(high)]],
            "kind": "bin", "totbins": numbins + sum(1 for c in flags[:3] if c != "_")}


def _split_axis(field, edges, flags="UONL"):
    return {"goal": ["call", "histbook.split" + flags, ["name", field], ["const", {"tuple": [float(e) for e in edges]}]],
            "kind": "split", "totbins": len(edges) - 1 + sum(1 for c in flags[:3] if c != "_")}


def _hist(fixed, weight=None):
    shape = [a["totbins"] for a in fixed] + [1 if weight is None else 2]
    w = None if weight is None else {"w": ["name", weight], "w2": ["call", "numpy.multiply", ["name", weight], ["name", weight]]}
    return {"fixed": fixed, "group": [], "profile": [], "shape": shape, "sumw": 0, "sumw2": None if weight is None else 1, "weight": w}


def test_split_index_from_the_bin_index():
    """codegen._apply_split_tables: a split axis on a value that also has a closed-low bin axis takes its digitize count from
    the bin index and one table look-up.  Values ON the edges, one ulp beside them, on and beside the bin edges, far outside,
    NaN and +-inf; edges that are bin edges, edges that are not, closed-high splits, splits without flow bins, a bin axis
    without underflow (no shift), and a combination whose bins hold two edges (must keep the compare chain)."""
    from histbook_b200 import codegen
    combos = [
        ("a", (100, -5.0, 5.0, "UONL"), ((-2.0, -1.0, -0.5, 0.0, 0.5, 1.0, 2.0), "UONL")),          # c4's own
        ("b", (37, -3.3, 4.1, "UONL"), ((-2.05, 0.1, 0.7, 3.3), "UONH")),                              # nothing aligned, closed high
        ("c", (50, 0.0, 1.0, "_ONL"), ((0.1, 0.3, 0.30000000000000004, 0.9), "___L")),                 # two edges one ulp apart: one bin holds both -> chain
        ("d", (64, -1.0, 1.0, "_O_L"), ((-0.5, 0.25, 0.75), "_ONL")),                                   # no underflow bin: index without shift
        ("e", (10, -1e6, 1e6, "UONL"), ((-7e5, -1.0, 1.0, 3e5), "UO_H")),
    ]
    hists, arrays = [], {}
    rs = numpy.random.RandomState(21)
    n = 6000
    arrays["w"] = rs.uniform(0.5, 1.5, n)
    for field, (numbins, low, high, bflags), (edges, sflags) in combos:
        hists.append(_hist([_bin_axis(field, numbins, low, high, bflags)], weight="w"))       # the sort group of this value
        hists.append(_hist([_split_axis(field, edges, sflags)]))
        hists.append(_hist([_bin_axis(field, numbins, low, high, bflags), _split_axis(field, edges, sflags)]))
        hists.append(_hist([_split_axis(field, edges, sflags)], weight="w"))
        width = (high - low) / numbins
        special = list(edges) + [low + k * width for k in range(0, numbins + 1, max(1, numbins // 7))] + [low, high]
        vals = []
        for v in special:
            v = numpy.float64(v)
            vals.extend([v, numpy.nextafter(v, numpy.inf), numpy.nextafter(v, -numpy.inf), numpy.nextafter(numpy.nextafter(v, numpy.inf), numpy.inf)])
        vals.extend([numpy.nan, numpy.inf, -numpy.inf, 1e300, -1e300, 3e9 * width + low, 0.0, -0.0])
        col = rs.uniform(low - 0.1 * (high - low), high + 0.1 * (high - low), n)
        pos = rs.permutation(n)[:len(vals) * 3]
        col[pos] = numpy.tile(numpy.array(vals, dtype=numpy.float64), 3)
        arrays[field] = col
    spec = {"version": 1, "hists": hists}
    assert len(hists) > 16
    coltypes = dict((k, (v.dtype, False)) for k, v in arrays.items())
    kern = codegen.generate(spec, coltypes, {"split_table": True})
    assert kern.sorted is not None
    assert kern.source.count("const int4 sq") == 4                # combos a, b, d, e (each split goal is lowered once); c keeps the chain
    assert "const int4 sq" not in codegen.generate(spec, coltypes, {}).source      # (an option: measured slower on c4)
    _check(spec, arrays, {"split_table": True})
    _check(spec, arrays, {})


def test_split_table_is_refused_where_it_is_not_valid():
    from histbook_b200 import codegen
    assert codegen.split_table([0.3, 0.30000000000000004], 0.0, 50.0, 50) is None            # both edges in bin 15
    assert codegen.split_table([1.0, 0.5], 0.0, 1.0, 10) is None and codegen.split_table([float("nan")], 0.0, 1.0, 10) is None
    tab = codegen.split_table([-2.0, 0.0, 0.5], -5.0, 10.0, 100)
    assert len(tab) == 100 and tab[29] == (float("inf"), 0) and tab[30] == (-2.0, 0) and tab[31] == (float("inf"), 1)
    assert tab[50] == (0.0, 1) and tab[55] == (0.5, 2) and tab[99] == (float("inf"), 3)
    # an edge outside the axis: counted from the first bin on (below) or never (above)
    tab = codegen.split_table([-9.0, 9.0], -5.0, 10.0, 100)
    assert tab[0] == (float("inf"), 1) and tab[99] == (float("inf"), 1)


def _emulable(name):
    """golden cases the emulator harness takes with the sorted strategy forced on: one fill of plain numeric arrays, a
    float64 term on a fixed-axis histogram (something to sort), at most one group axis per histogram, exact arithmetic"""
    from histbook_b200 import codegen
    case = golden_util.load()[name]
    fills = golden_util.fills(case)
    if len(fills) != 1:
        return False
    arrays = fills[0]
    fields = emul_util.hbspec.spec_fields(case["spec"])
    if not fields or any(n not in arrays or not isinstance(arrays[n], numpy.ndarray) or arrays[n].dtype.kind not in "fiub" or arrays[n].ndim != 1
                         for n in fields):
        return False
    if any(len(h["group"]) > 1 for h in case["spec"]["hists"]):
        return False
    try:
        kern = codegen.generate(case["spec"], dict((n, (arrays[n].dtype, False)) for n in fields), {"sort": True})
    except Exception:
        return False
    return kern.sorted is not None and kern.window is None and not kern.fx_terms and not kern.inexact


@pytest.mark.parametrize("name", golden_util.names())
def test_golden_cases_sorted_on_the_emulator(name):
    """every golden case the harness can take, through the sorted kernel (forced on), against the oracle"""
    if not _emulable(name):
        pytest.skip("not a case for the emulator harness")
    case = golden_util.load()[name]
    arrays = golden_util.fills(case)[0]
    _check(case["spec"], arrays, {"sort": True}, grid=2)


def test_a_barrier_nobody_completes_fails_the_launch_instead_of_hanging_it():
    """the stand-in's watchdog (tests/emul/cuda_emul.h): one thread leaves before the block barrier"""
    import ctypes
    source = ('#define HB_EMU_BARRIER_TIMEOUT 1\n#include "hb_template.cuh"\n'
              'extern "C" __global__ void hb_fill_kernel(const HbParams p)\n{\n'
              '    extern __shared__ __align__(16) unsigned char hb_smem[];\n'
              '    if (threadIdx.x == 3 && p.n == 1) return;\n    __syncthreads();\n'
              '    if (threadIdx.x == 0) reinterpret_cast<double*>(p.hist[0])[0] += 1.0;\n}\n')
    lib = emul_util.build(source)
    lib.hb_emu_run.argtypes = [ctypes.POINTER(emul_util.HbParams), ctypes.c_int, ctypes.c_int]
    out = numpy.zeros(1)
    p = emul_util.HbParams()
    p.hist[0] = out.ctypes.data
    p.n = 0
    assert lib.hb_emu_run(ctypes.byref(p), 2, 64) == 0 and out[0] == 2.0          # every thread arrives: two blocks, two adds
    p.n = 1
    assert lib.hb_emu_run(ctypes.byref(p), 1, 64) == 1                            # thread 3 never arrives: reported after the timeout
