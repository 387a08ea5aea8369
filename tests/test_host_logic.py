"""CPU-side checks (no GPU): the C ABI library loads and exports every declared symbol, the code
generator lowers every golden spec and NVRTC accepts the result for sm_100a, column binding follows
histbook/fill.py:88-146, and the device path refuses to run without a device."""
import ctypes
import os
import re

import numpy
import pytest

import golden_util
from histbook_b200 import codegen, columns, engine, spec as hbspec

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(REPO, "include", "hbfill.h")).read()
    declared = set(re.findall(r"\b(hb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    lib = engine.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), "libhbfill.so does not export " + name
    assert declared == set(engine.SYMBOLS), (declared ^ set(engine.SYMBOLS))
    assert lib.hb_version() == 1


def test_no_cpu_fallback():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    from histbook_b200 import fillable
    sf = fillable.SpecFill(golden_util.load()["c1"]["spec"])
    with pytest.raises(engine.EngineError):
        sf.fill(x=numpy.zeros(10))


def _coltypes(case):
    f = golden_util.fills(case)[0]
    out = {}
    for n, v in f.items():
        a = numpy.asarray(v)
        if a.dtype.kind in "US":                 # string groupby keys reach the kernel as int64 dictionary codes
            a = numpy.zeros(a.shape, dtype=numpy.int64)
        out[n] = (a.dtype, a.shape == ())
    for n in hbspec.spec_fields(case["spec"]):
        if n not in out and n in codegen.MAYBECONSTANTS:
            out[n] = (numpy.dtype("float64"), True)
    return out


@pytest.mark.parametrize("name", ["c1", "c2", "c3", "c4", "c5", "library_exact", "dtype_float32", "dtype_uint8",
                                  "groupby_groupbin_num", "split3_flags_0110", "scalar_broadcast", "maybeconstants"])
def test_generated_kernels_compile_for_sm100a(name):
    case = golden_util.load()[name]
    kern = codegen.generate(case["spec"], _coltypes(case))
    cubin = engine.compile_source(kern.source)
    assert cubin[:4] == b"\x7fELF"
    assert "--fmad" not in kern.source and "hb_fill_kernel" in kern.source
    keys = codegen.generate_keys(case["spec"], _coltypes(case))
    if any(h["group"] for h in case["spec"]["hists"]):
        assert engine.compile_source(keys.source)[:4] == b"\x7fELF"
    else:
        assert keys is None


def test_pipelined_sorted_kernel_compiles_and_is_laid_out_twice():
    """sort_pipe: producer warps evaluate tile k + 1 while one consumer warp per sort group sums tile k; tick counters,
    sorted event numbers and staged values exist twice and everything fits the SM's shared memory."""
    case = golden_util.load()["c4"]
    kern = codegen.generate(case["spec"], _coltypes(case), {"sort_pipe": True})
    pipe = kern.sorted["pipe"]
    assert pipe["consumer_warps"] == 8 and pipe["producers"] == 1024 - 8 * 32 and kern.tile == pipe["producers"] * kern.sorted["ept"]
    assert kern.smem_bytes <= 227 * 1024
    assert "bar.arrive" in kern.source and kern.source.count("hb_sort_find(") == 1          # ONE copy of the summing code for 8 like groups
    assert engine.compile_source(kern.source)[:4] == b"\x7fELF"
    # a forced single group (any histogram with a float64 term)
    case = golden_util.load()["c3"]
    kern = codegen.generate(case["spec"], _coltypes(case), {"sort": True, "sort_pipe": True})
    assert kern.sorted["pipe"]["consumer_warps"] == 1 and kern.tile == 992 * kern.sorted["ept"]
    assert engine.compile_source(kern.source)[:4] == b"\x7fELF"
    # and the kernel with phases behind block barriers is still there
    kern = codegen.generate(golden_util.load()["c4"]["spec"], _coltypes(golden_util.load()["c4"]), {"sort_pipe": False})
    assert "pipe" not in kern.sorted and "bar.arrive" not in kern.source


def test_every_golden_spec_lowers():
    for name in golden_util.names():
        case = golden_util.load()[name]
        kern = codegen.generate(case["spec"], _coltypes(case))
        assert len(kern.plans) == len(case["spec"]["hists"])


def test_common_subexpressions_are_shared_across_a_book():
    case = golden_util.load()["c4"]
    kern = codegen.generate(case["spec"], _coltypes(case))
    # x*x appears in sqrt(x**2+y**2) and sqrt(x**2+y**2+z**2), for 8 histograms each: one multiply
    assert kern.source.count("hb_mul<double>(c3, c3)") + kern.source.count("hb_mul<double>(c2, c2)") <= 2
    assert len([l for l in kern.source.splitlines() if "sqrt(" in l]) == 2


def test_identical_accumulations_are_merged():
    # c3: profile("y")'s sum-of-squares and profile("y**2")'s sum are the same term (y*y)
    case = golden_util.load()["c3"]
    kern = codegen.generate(case["spec"], _coltypes(case))
    hp = kern.plans[0]
    assert len(hp.terms) == 4 and sorted(len(d) for d in hp.dest) == [1, 1, 1, 2]


def test_dtype_follows_numpy_promotion():
    spec = golden_util.load()["dtype_float32"]["spec"]
    k32 = codegen.generate(spec, {"x": (numpy.dtype("float32"), False)})
    k64 = codegen.generate(spec, {"x": (numpy.dtype("float64"), False)})
    assert "hb_bin_axis<float," in k32.source and "hb_bin_axis<double," in k64.source


def test_unknown_function_is_an_assertion_error_like_the_reference():
    spec = {"version": 1, "hists": [{"group": [], "fixed": [{"kind": "bin", "totbins": 13, "goal":
            ["call", "histbook.binUONL", ["call", "round", ["name", "x"]], ["const", 10], ["const", 0.0], ["const", 1.0]]}],
            "profile": [], "weight": None, "shape": [13, 1], "sumw": 0, "sumw2": None}]}
    with pytest.raises(AssertionError):
        codegen.generate(spec, {"x": (numpy.dtype("float64"), False)})


def test_bind_columns_errors_and_broadcast():
    cols, n = columns.bind_all(["pi", "w", "x"], {"x": [1.0, 2.0, 3.0], "w": 2})
    assert n == 3 and cols["w"].is_scalar and cols["w"].dtype == numpy.int64 and cols["pi"].is_scalar
    assert cols["x"].dtype == numpy.float64 and cols["x"].location == 0
    with pytest.raises(ValueError, match="required field 'x' not found in fill arguments"):
        columns.bind_all(["x"], {})
    with pytest.raises(ValueError, match="array 'y' has len 2 but other arrays have len 3"):
        columns.bind_all(["x", "y"], {"x": numpy.zeros(3), "y": numpy.zeros(2)})
    with pytest.raises(codegen.UnsupportedError):
        columns.bind_all(["c"], {"c": ["a", "b"]})
    cols, n = columns.bind_all(["x"], {"x": numpy.arange(10.0)[::2]})      # strided -> contiguous copy
    assert n == 5 and cols["x"].keepalive.flags.c_contiguous


def test_group_table_layout_roundtrip():
    from histbook_b200.store import GroupInfo, KEY_NAN
    g = GroupInfo("groupbin", False, True, "float64", True, None)
    for key in (0.0, -2.5, 10.0, 1e300):
        assert g.key_object(g.pattern(key)) == key
    assert g.pattern(-0.0) == g.pattern(0.0)
    assert g.key_object(KEY_NAN) == "NaN" and g.pattern("NaN") == KEY_NAN
    order = g.sort_patterns([g.pattern(3.0), KEY_NAN, g.pattern(-1.0), g.pattern(0.5)])
    assert [g.key_object(p) for p in order] == [-1.0, 0.5, 3.0, "NaN"]
    gi = GroupInfo("groupby", False, False, "int64", True, None)
    assert [gi.key_object(p) for p in gi.sort_patterns([gi.pattern(5), gi.pattern(-3), gi.pattern(0)])] == [-3, 0, 5]


def test_hot_window_placement():
    from histbook_b200.fillable import choose_ragged, window_table, window_params
    rs = numpy.random.RandomState(3)
    x, y = rs.normal(0, 1, 200000), rs.normal(0, 1, 200000)
    i0 = numpy.clip(numpy.floor((x + y + 5) * 20).astype(int) + 1, 0, 201)
    i1 = numpy.clip(numpy.floor(numpy.sqrt(x * x + y * y) * 40).astype(int) + 1, 0, 201)
    h = numpy.zeros((203, 203))
    numpy.add.at(h, (i0, i1), 1)
    first, length, cov = choose_ragged(h, 6400)
    assert length.sum() <= 6400 and cov > 0.9            # a rectangle of 6400 bins holds ~0.80 of these events
    inside = sum(h[r, first[r]:first[r] + length[r]].sum() for r in range(203))
    assert abs(inside / h.sum() - cov) < 1e-12
    # the table the kernel reads: low word = offset of the row's interval, high word = first << 16 | length
    table, bins = window_table(first, length)
    assert bins == length.sum() and table.dtype == numpy.uint64
    assert numpy.array_equal(table & numpy.uint64(0xffffffff), numpy.concatenate([[0], numpy.cumsum(length)[:-1]]).astype(numpy.uint64))
    assert numpy.array_equal((table >> numpy.uint64(32)) & numpy.uint64(0xffff), length.astype(numpy.uint64))
    assert numpy.array_equal(table >> numpy.uint64(48), first.astype(numpy.uint64))
    # nothing filled yet -> no window
    first, length, cov = choose_ragged(numpy.zeros((5, 7)), 100)
    assert length.sum() == 0 and cov == 0.0
    # everything fits -> the populated span of every row
    first, length, cov = choose_ragged(numpy.array([[0, 1.0, 2, 0], [3.0, 0, 1, 0], [0, 0, 0, 0]]), 100)
    assert first.tolist() == [1, 0, 0] and length.tolist() == [2, 3, 0] and cov == 1.0
    # budget smaller than the number of populated rows -> single bins of the densest rows
    first, length, cov = choose_ragged(numpy.array([[0, 1.0, 2, 0], [3.0, 0, 1, 0], [0, 0, 0, 9.0]]), 2)
    assert length.tolist() == [0, 1, 1] and first.tolist()[1:] == [0, 3]
    # a 1-d histogram is one row
    first, length, cov = choose_ragged(numpy.array([[0, 1.0, 5, 7, 2, 0.5]]), 3)
    assert first.tolist() == [2] and length.tolist() == [3]

    class K(object):
        window = {"bpb": 16, "wmax": 1200}
    assert window_params(K, 1000) == [4800, 1000] and window_params(K, 0) == [0, 0]


def test_string_codes():
    from histbook_b200.columns import StringCodes, is_stringlike
    from histbook_b200.store import GroupInfo
    sc = StringCodes()
    a = sc.encode(["one", "two", "three", "two", "one"])
    assert a.dtype == numpy.int64 and a[0] == a[4] and a[1] == a[3] and len(set(a.tolist())) == 3
    b = sc.encode(numpy.array(["two", "zero"]))
    assert b[0] == a[1] and b[1] not in a.tolist()                  # codes are stable across fills
    assert str(sc.decode(a[2])) == "three"
    assert is_stringlike(["x"]) and is_stringlike(numpy.array(["x"])) and not is_stringlike([1.0]) and not is_stringlike(numpy.zeros(3))
    g = GroupInfo("groupby", False, False, "int64", True, None, codes=sc)
    assert g.key_object(g.pattern("two")) == "two"
    assert [str(g.key_object(p)) for p in g.sort_patterns([g.pattern(k) for k in ("two", "one", "zero", "three")])] == ["one", "three", "two", "zero"]
    assert g.from_wire(g.to_wire(g.pattern("zero"))) == g.pattern("zero")
    with pytest.raises(Exception):
        g.pattern(3)


def test_accumulation_kinds_chosen_for_the_baseline_configs():
    """fx="auto" (codegen.choose_strategies): which float64 terms become exact fixed point, which stay float64 CAS."""
    import golden_util
    from histbook_b200 import codegen, spec as hbspec

    def gen(name, **options):
        spec = golden_util.load()[name]["spec"]
        coltypes = dict((n, (numpy.dtype("int64" if n == "n" else "float64"), False)) for n in hbspec.spec_fields(spec))
        return codegen.generate(spec, coltypes, options)
    k = gen("c1")
    assert [hp.strategy for hp in k.plans] == ["smem"] and not k.fx_terms and k.threads == 1024
    k = gen("c3")                                   # 2006 rows: all but the first term, 2 blocks of 512 threads
    assert [hp.strategy for hp in k.plans] == ["smem"] and len(k.fx_terms) == 2 and k.threads == 512 and k.smem_bytes < 113 * 1024
    assert len(gen("c3", fx=False).fx_terms) == 0 and len(gen("c3", fx=True).fx_terms) == 3
    k = gen("c2")                                   # hot window, sparse variant: float64 CAS, paired REDs behind a vote
    assert k.plans[0].strategy == "window" and not k.fx_terms and "__any_sync" in k.source and "hb_red_pair" in k.source
    d = gen("c2", win_dense=True)                   # dense variant: sum(w^2) in fixed point, plain REDs outside
    assert len(d.fx_terms) == 1 and "hb_red_pair(" not in d.source.split("hb_event")[1] and d.window["wmax"] < k.window["wmax"]
    assert d.window["bpb"] == 24 and k.window["bpb"] == 16
    k = gen("c5", box=False)                        # 134 MB: global paired REDs only, no window
    assert [hp.strategy for hp in k.plans] == ["global"] and k.window is None and k.smem_bytes == 0
    k = gen("c5")                                   # default: a box window [lo_k, lo_k + n_k) per axis, placed at run time (if it would hold >= 25 %)
    assert k.plans[0].strategy == "window" and k.window["box"] == [203, 203, 203] and k.window["box_base"] == 2 and "hb_red_pair" in k.source
    k = gen("c4", sort=False)                       # big book, per-event strategies: one copy of the event code, everything privatised, all fixed point
    assert all(hp.strategy == "smem" for hp in k.plans) and len(k.fx_terms) == 48 and k.threads == 768
    assert k.nfx == 6 and k.source.count("hb_fx_decode<") == 6        # w, w^2, z, z^2, filtered w, filtered w^2: decoded once per event
    assert k.source.count("hb_event(p, hb_smem") == 1
    k = gen("c4")                                   # default: one sort group per axis expression holds its 4 histograms on bin("e", 100)
    # (w and z are staged; the filtered weight where(n >= 3, 1, 0) * w travels as one flag bit beside the sorted event numbers)
    assert k.sorted is not None and len(k.sorted["groups"]) == 8 and k.sorted["staged"] == 2 and k.sorted["flags"] == 1 and not k.fx_terms
    assert gen("c4", sort_gate=False).sorted["staged"] == 3 and gen("c4", sort_gate=False).sorted["flags"] == 0
    assert "hb_event<true>" in k.source and "hb_event<true>" not in gen("c4", sort_full=False).source      # full tiles: `live` folded away
    assert all(g["nbins"] == 103 and g["values"] == 6 and len(g["hists"]) == 4 for g in k.sorted["groups"])
    assert sorted(hp.strategy for hp in k.plans).count("sorted") == 32 and k.threads == 1024 and k.tile == 3072      # (three events per thread where it all fits)
    assert gen("c4", sort_ept=2).tile == 2048 and all(hp.strategy in ("sorted", "smem") for hp in k.plans)
    assert k.source.count("atomicAdd(") - k.source.count("hb_sadd_u32") <= 8 + 1     # one ticket per group and event
    f = gen("c4", group_fast=True)                  # the variant fillable picks when the histograms of a signature share their slot maps
    assert f.source.count("gm[") == 1 and k.source.count("gm[") == 8 and "hb_red_f64(row" not in f.source.split("hb_fill_kernel")[0]
    assert f.group_aux == k.group_aux and f.goal_aux == k.goal_aux and f.naux == k.naux and f.smem_bytes == k.smem_bytes
    k = gen("c3", sort=True)                        # forced on a small book: y and y^2 staged, y^4 = (y^2)^2 recomputed
    assert k.sorted["staged"] == 2 and k.sorted["groups"][0]["values"] == 3 and k.smem_bytes < 227 * 1024 and k.sorted["lanes"] == 128
    k = gen("c2", deterministic=True)
    assert k.deterministic and len(k.fx_terms) == 2 and "hb_fxg_add128" in k.source and "hb_red_pair(" not in k.source.split("hb_event")[1]


def test_zero_copy_ingest_of_host_containers():
    """columns.bind uses host memory in place where it can: numpy arrays, pandas Series, Arrow arrays."""
    import pandas
    import pyarrow
    from histbook_b200 import columns
    a = numpy.arange(10, dtype=numpy.float64)
    assert columns.bind("x", a).ptr == a.ctypes.data
    s = pandas.Series(a)
    assert columns.bind("x", s).ptr == s.values.ctypes.data
    arr = pyarrow.array(a)
    c = columns.bind("x", arr)
    assert c.dtype == numpy.float64 and c.length == 10 and c.ptr == arr.buffers()[1].address
    ch = pyarrow.chunked_array([a[:4], a[4:]])
    c = columns.bind("x", ch)
    assert c.length == 10 and numpy.array_equal(c.keepalive, a)
    withnull = pyarrow.array([1.0, None, 3.0])
    c = columns.bind("x", withnull)
    assert numpy.isnan(c.keepalive[1]) and c.keepalive[2] == 3.0
    with pytest.raises(Exception):
        columns.bind("n", pyarrow.array([1, None, 3]))
    i = pyarrow.array(numpy.arange(5, dtype=numpy.int32))
    assert columns.bind("n", i).dtype == numpy.int32


def test_arena_exposes_cuda_array_interface_as_property():
    """dist.allreduce hands the arena to torch.as_tensor (zero copy): the protocol needs an attribute, not a method."""
    from histbook_b200 import engine
    assert isinstance(engine.Arena.__dict__["__cuda_array_interface__"], property)


def test_cuda_array_interface_stream_key_is_recorded():
    """__cuda_array_interface__ v3 ``stream``: None and 1 (legacy default stream) need no fence, anything else is kept
    for hb_stream_wait (fillable.fill_hists).  Binding never touches the memory, so a fake exporter will do on CPU."""
    from histbook_b200 import columns, engine

    class Exporter(object):
        def __init__(self, stream):
            self.__cuda_array_interface__ = {"shape": (8,), "typestr": "<f8", "data": (0x7F0000000000, False), "version": 3, "stream": stream}

    for stream, expect in ((None, None), (1, None), (2, 2), (0x5500AA00, 0x5500AA00)):
        col = columns.bind("x", Exporter(stream))
        assert (col.location, col.length, col.ptr, col.stream, col.code) == (1, 8, 0x7F0000000000, expect, 0)
    assert engine.producer_stream() == 0                      # no CUDA context in this process: nothing to wait for


def test_box_window_placement():
    """fillable.choose_box: one interval per axis around the bulk, at most wmax bins in the box, capacity used up."""
    from histbook_b200 import fillable
    x = numpy.arange(103)
    narrow = numpy.exp(-0.5 * ((x - 40.0) / 4.0) ** 2)
    wide = numpy.exp(-0.5 * ((x - 60.0) / 12.0) ** 2)
    lo, n, coverage = fillable.choose_box([narrow, wide, narrow], 14000)
    assert numpy.prod(n) <= 14000 and numpy.prod(n) > 0.8 * 14000
    assert lo[0] <= 40 < lo[0] + n[0] and lo[1] <= 60 < lo[1] + n[1] and n[1] > n[0]      # the wide axis takes more bins
    exact = numpy.prod([m[a:a + k].sum() / m.sum() for m, a, k in zip([narrow, wide, narrow], lo, n)])
    assert abs(coverage - exact) < 1e-12 and coverage > 0.5
    assert fillable.choose_box([narrow, wide * 0.0], 100) == ([0, 0], [0, 0], 0.0)        # an empty marginal: no box
    lo, n, coverage = fillable.choose_box([narrow], 1000)                                   # room for everything
    assert lo == [0] and n == [103] and abs(coverage - 1.0) < 1e-12
