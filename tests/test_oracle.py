"""The CPU oracle (oracle/fill_oracle.py) against the reference's own outputs.

Bit-exact on every golden case: the oracle restates the same numpy calls in the
same order, so even float sums must be identical.
"""
import numpy
import pytest

import golden_util
from oracle import fill_oracle

CASES = golden_util.names()


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference(name):
    case = golden_util.load()[name]
    contents = None
    for arrays in golden_util.fills(case):
        contents = fill_oracle.fill(case["spec"], arrays, contents)
    expect = golden_util.outputs(case)
    assert len(contents) == len(expect)
    for got, exp in zip(contents, expect):
        flat = golden_util.flatten(got)
        assert [k for k, _ in flat] == [k for k, _ in exp]
        for (k, a), (_, b) in zip(flat, exp):
            assert a.shape == b.shape
            assert a.dtype == numpy.float64
            assert numpy.array_equal(a, b, equal_nan=True), (name, k)


def test_missing_field_message():
    spec = golden_util.load()["c2"]["spec"]
    with pytest.raises(ValueError, match="required field 'w' not found in fill arguments"):
        fill_oracle.fill(spec, {"x": numpy.zeros(3), "y": numpy.zeros(3)})


def test_length_mismatch_message():
    spec = golden_util.load()["c2"]["spec"]
    with pytest.raises(ValueError, match="has len . but other arrays have len ."):
        fill_oracle.fill(spec, {"x": numpy.zeros(3), "y": numpy.zeros(3), "w": numpy.zeros(2)})
