"""The product's fill path on the EMULATED engine (tests/emul_engine.py) against the UNMODIFIED reference on seeded random
histogram definitions -- the CPU twin of tests/test_random_definitions_gpu.py (same generator, same comparisons), once with
the default kernels and once with the sorted strategy of big books forced on (tickets, scatter, chunked sums, flag bits for
filtered weights ... on whatever the generator draws).  Offline hunts over 25 more seeds and six option sets (424
definitions) found no mismatch."""
import shutil
import warnings

import numpy
import pytest

histbook_b200 = pytest.importorskip("histbook_b200")
if not histbook_b200._ref.available():
    pytest.skip("the histbook package (baseline/_ref) is not available", allow_module_level=True)

import emul_engine                                    # noqa: E402
import golden_util                                    # noqa: E402
import test_oracle_live as live                       # noqa: E402
from histbook_b200 import codegen, fillable            # noqa: E402
from histbook_b200 import spec as hbspec              # noqa: E402
from oracle import fill_oracle                        # noqa: E402
from test_gpu_parity import compare                   # noqa: E402
from test_random_definitions_gpu import EXACT_EXPRS   # noqa: E402

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="the kernel emulator is compiled with g++")


@pytest.mark.parametrize("options", [{}, {"sort": True}], ids=["default", "sorted"])
def test_emulated_fill_equals_reference_on_random_definitions(options, monkeypatch):
    monkeypatch.setattr(live, "EXPRS", EXACT_EXPRS)
    rs = numpy.random.RandomState(7100 + len(options))
    accepted = []
    with histbook_b200.reference_path() as hb, warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = dict((name, getattr(hb, name)) for name in ("Hist", "bin", "intbin", "split", "cut", "profile", "groupby", "groupbin"))
        for _ in range(18):
            source = live._definition(rs)
            fills = [live._columns(rs, int(rs.choice([1, 17, 500, 4099, 9000]))) for _ in range(int(rs.choice([1, 2])))]
            try:
                hist = eval(source, dict(env))
                for cols in fills:
                    hist.fill(**cols)
            except Exception:
                continue                                    # the reference refuses this definition / these rows
            if hist._content is None:
                continue
            spec = hbspec.book_spec([hist])
            fields = hbspec.spec_fields(spec)
            accepted.append((source, spec, [dict((f, cols[f]) for f in fields) for cols in fills], golden_util.flatten(hist._content)))
    compared = 0
    with emul_engine.installed(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            fillable.set_options(**options)
            for source, spec, fills, exp in accepted:
                try:
                    sf = fillable.SpecFill(spec)
                    for cols in fills:
                        sf.fill(cols)
                    got = sf.contents()
                except codegen.UnsupportedError:
                    continue                                # documented divergences
                bound = None
                for cols in fills:
                    bound = fill_oracle.fill(spec, cols, bound, absolute=True)
                compare(source, spec, got, [exp], bound)
                compared += 1
        finally:
            fillable.set_options(**dict((k, None) for k in options))
    assert compared >= 8, (compared, len(accepted))
