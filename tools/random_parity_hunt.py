"""Offline hunt with the generator of tests/test_random_definitions_gpu.py over many more seeds: device (C ABI) against the
unmodified reference (baseline/_ref), mismatches printed instead of asserted.  Usage (GPU box): python tools/random_parity_hunt.py 100 120 [dtypes]"""
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy  # noqa: E402

import histbook_b200  # noqa: E402
import golden_util  # noqa: E402
import test_oracle_live as live  # noqa: E402
from histbook_b200 import codegen, fillable  # noqa: E402
from histbook_b200 import spec as hbspec  # noqa: E402
from oracle import fill_oracle  # noqa: E402
from test_gpu_parity import compare  # noqa: E402
from test_random_definitions_gpu import EXACT_EXPRS  # noqa: E402


def cast(cols, rs):
    """Mixed column dtypes (numpy computes in the column's dtype: float32 stays float32, small ints promote like NEP 50)."""
    out = dict(cols)
    mode = rs.randint(0, 4)
    if mode == 0:
        out["x"] = out["x"].astype(numpy.float32)
    elif mode == 1:
        for name in ("x", "y", "w"):
            out[name] = out[name].astype(numpy.float32)
    elif mode == 2:
        out["z"] = numpy.nan_to_num(out["z"] * 10).astype(numpy.int16)
        out["n"] = out["n"].astype(numpy.uint8)
    else:
        out["y"] = numpy.nan_to_num(out["y"], nan=0, posinf=9, neginf=-9).astype(numpy.int64)
        out["w"] = out["w"].astype(numpy.float32)
    return out


def main():
    lo, hi = int(sys.argv[1]), int(sys.argv[2])
    dtypes = len(sys.argv) > 3 and sys.argv[3] == "dtypes"
    live.EXPRS = EXACT_EXPRS
    compared = unsupported = 0
    bad = []
    for seed in range(lo, hi):
        rs = numpy.random.RandomState(seed)
        accepted = []
        with histbook_b200.reference_path() as hb, warnings.catch_warnings():
            warnings.simplefilter("ignore")
            env = dict((name, getattr(hb, name)) for name in ("Hist", "bin", "intbin", "split", "cut", "profile", "groupby", "groupbin"))
            for _ in range(30):
                source = live._definition(rs)
                fills = [live._columns(rs, int(rs.choice([1, 17, 500, 4099, 20000]))) for _ in range(int(rs.choice([1, 2])))]
                if dtypes:
                    fills[0] = cast(fills[0], rs)
                    fills[1:] = [dict((k, v.astype(fills[0][k].dtype)) for k, v in cols.items()) for cols in fills[1:]]
                try:
                    hist = eval(source, dict(env))
                    for cols in fills:
                        hist.fill(**cols)
                except Exception:
                    continue
                if hist._content is None:
                    continue
                spec = hbspec.book_spec([hist])
                fields = hbspec.spec_fields(spec)
                accepted.append((source, spec, [dict((f, cols[f]) for f in fields) for cols in fills], golden_util.flatten(hist._content)))
        for source, spec, fills, exp in accepted:
            try:
                sf = fillable.SpecFill(spec)
                for cols in fills:
                    sf.fill(cols)
                bound = None
                for cols in fills:
                    bound = fill_oracle.fill(spec, cols, bound, absolute=True)
                compare(source, spec, sf.contents(), [exp], bound)
                compared += 1
            except codegen.UnsupportedError:
                unsupported += 1
            except Exception as err:
                bad.append((seed, source, "%s: %s" % (type(err).__name__, str(err)[:200])))
    print("seeds %d..%d: %d definitions agree, %d unsupported, %d mismatches" % (lo, hi - 1, compared, unsupported, len(bad)))
    for item in bad[:20]:
        print(item)


if __name__ == "__main__":
    main()
