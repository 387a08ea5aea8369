"""Deterministic mode on the EMULATED engine (tests/emul_engine.py): the generated kernels add every float64 term as an exact
fixed-point integer (shared memory per block, then the 2 x 192-bit global accumulators), the stand-in's ``fx_finalize`` adds the
two windows with Python integers and rounds once -- so a fill must equal, BIT FOR BIT, the correctly rounded exact sum of its
terms (``fill_oracle.fill(..., exact=True)``, math.fsum per bin), whatever the launch shape.  The CPU twin of
tests/test_deterministic_gpu.py at sizes CPU threads finish in seconds."""
import shutil

import pytest

import emul_engine
import golden_util
import test_deterministic_gpu as det
from oracle import fill_oracle

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="the kernel emulator is compiled with g++")


def _fills(name, n, neg, seed=11):
    from histbook_b200 import spec as hbspec
    spec = golden_util.load()[name]["spec"]
    fields = hbspec.spec_fields(spec)
    return spec, [det._columns(fields, n, seed, neg), det._columns(fields, n // 3 + 7, seed + 1, neg)]


@pytest.mark.parametrize("name,n,neg", [("c1", 9000, False), ("c2", 15000, False), ("c2", 15000, True), ("c3", 12000, False),
                                        ("c4", 7000, False), ("c5", 9000, False)])
def test_emulated_fill_equals_the_exactly_rounded_sum(name, n, neg):
    spec, fills = _fills(name, n, neg)
    with emul_engine.installed():
        got, missed = det._run(spec, fills)
    exp = None
    for arrays in fills:
        exp = fill_oracle.fill(spec, arrays, exp, exact=True)
    assert missed == 0
    det._assert_identical(got, exp)


def test_emulated_result_does_not_depend_on_the_launch_shape():
    spec, fills = _fills("c2", 12000, True, seed=21)
    with emul_engine.installed():
        base, _ = det._run(spec, fills)
        for options in ({"threads": 256, "min_blocks": 1, "win_threads": 256, "win_min_blocks": 1, "win_smem": 64 * 1024},
                        {"window": False, "strategy": {"*": "global"}}):
            other, _ = det._run(spec, fills, **options)
            det._assert_identical(base, other)
