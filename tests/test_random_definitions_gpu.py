"""The CUDA fill path (through the C ABI) against the UNMODIFIED reference on seeded random histogram definitions.

Companion of tests/test_oracle_live.py (same generator): for every definition the reference accepts, the device result
must have the reference's dict keys in the reference's order, bit-exact unweighted counts, and weighted / profile sums
within 1e-12 * (per-bin sum of |term|) (bound from ``oracle.fill_oracle(..., absolute=True)``).  Only expressions made of
IEEE-exact operations are drawn (+ - * / sqrt abs floor min where % sign): bin indices are then bit-exact too; functions
that go through CUDA's libm are covered by test_gpu_parity.test_libm_class_functions_flip_budget.
"""
import warnings

import numpy
import pytest

pytestmark = pytest.mark.gpu

histbook_b200 = pytest.importorskip("histbook_b200")
if not histbook_b200._ref.available():
    pytest.skip("the histbook package (baseline/_ref) is not available", allow_module_level=True)

import golden_util                                   # noqa: E402
import test_oracle_live as live                      # noqa: E402
from histbook_b200 import codegen, fillable           # noqa: E402
from histbook_b200 import spec as hbspec             # noqa: E402
from oracle import fill_oracle                       # noqa: E402
from test_gpu_parity import compare                  # noqa: E402

EXACT_EXPRS = ["x", "y", "x + y", "x - y", "x * y", "sqrt(x**2 + y**2)", "abs(x)", "x**2", "2*x + 1", "x/(1 + y**2)",
               "x + y + z", "z", "min(x, y)", "where(x > 0, x, y)", "floor(x) + 0.5", "x % 2", "sign(y)*x**2"]


@pytest.mark.parametrize("seed", range(3))
def test_device_equals_reference_on_random_definitions(seed, monkeypatch):
    monkeypatch.setattr(live, "EXPRS", EXACT_EXPRS)
    rs = numpy.random.RandomState(5000 + seed)
    accepted = []
    with histbook_b200.reference_path() as hb, warnings.catch_warnings():
        warnings.simplefilter("ignore")
        env = dict((name, getattr(hb, name)) for name in ("Hist", "bin", "intbin", "split", "cut", "profile", "groupby", "groupbin"))
        for _ in range(30):
            source = live._definition(rs)
            fills = [live._columns(rs, int(rs.choice([1, 17, 500, 4099, 20000]))) for _ in range(int(rs.choice([1, 2])))]
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
    for source, spec, fills, exp in accepted:
        try:
            sf = fillable.SpecFill(spec)
            for cols in fills:
                sf.fill(cols)
        except codegen.UnsupportedError:
            continue                                        # documented divergences
        bound = None
        for cols in fills:
            bound = fill_oracle.fill(spec, cols, bound, absolute=True)
        compare(source, spec, sf.contents(), [exp], bound)
        compared += 1
    assert compared >= 12, (compared, len(accepted))
