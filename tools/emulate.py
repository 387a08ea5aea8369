#!/usr/bin/env python
"""Run one BASELINE configuration (or any golden case's spec) through the product's host path on the CPU emulator of the
generated kernels (tests/emul_engine.py) with the given kernel options, and compare with the oracle -- what a change of
histbook_b200/codegen.py or csrc/hb_template.cuh is put through before it is given GPU time.

    python tools/emulate.py c4 '{"sort_pair": true}' 20000
    python tools/emulate.py c2 '{}' 30000            # pilot -> hot window -> fixed point
    python tools/emulate.py c3 '{"deterministic": true}'
"""
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

import numpy                                        # noqa: E402

import emul_engine                                  # noqa: E402
import golden_util                                  # noqa: E402
from histbook_b200 import fillable                  # noqa: E402
from histbook_b200 import spec as hbspec            # noqa: E402
from oracle import fill_oracle                      # noqa: E402
from test_gpu_parity import compare                 # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c4"
    options = json.loads(sys.argv[2]) if len(sys.argv) > 2 else {}
    n = int(float(sys.argv[3])) if len(sys.argv) > 3 else 20000
    spec = golden_util.load()[name]["spec"]
    rs = numpy.random.RandomState(3)
    cols = {"x": rs.normal(0, 1, n), "y": rs.normal(0, 1, n), "z": rs.normal(0, 1, n), "w": rs.uniform(0.5, 1.5, n),
            "n": rs.poisson(3, n).astype(numpy.int64)}
    cols["x"][::997] = numpy.nan
    cols["w"][1::1009] = numpy.nan
    cols["y"][2::1013] = numpy.inf
    fields = hbspec.spec_fields(spec)
    missing = [f for f in fields if f not in cols]
    if missing:
        raise SystemExit("this tool feeds the fields x, y, z, w, n; %s needs %s" % (name, missing))
    fills = [dict((k, cols[k][:n // 2 + 7]) for k in fields), dict((k, cols[k][n // 2 + 7:]) for k in fields)]
    with emul_engine.installed():
        fillable.set_options(**options)
        try:
            sf = fillable.SpecFill(spec)
            t0 = time.time()
            for arrays in fills:
                sf.fill(arrays)
            got = sf.contents()
            stats = dict(fillable.last_stats)
            launches = [(threads, events) for full, threads, events, variant in emul_engine.LAUNCHES]
        finally:
            fillable.set_options(**dict((k, None) for k in options))
    exact = bool(options.get("deterministic"))
    exp = bound = None
    for arrays in fills:
        exp = fill_oracle.fill(spec, arrays, exp, exact=exact)
        bound = fill_oracle.fill(spec, arrays, bound, absolute=True)
    compare(name, spec, got, [golden_util.flatten(e) for e in exp], bound)
    print("%s, %d events in two fills, options %s: equal to the oracle (counts bit-exact, sums within 1e-12 * sum|term|) in %.1f s" % (
        name, n, options, time.time() - t0))
    print("launches (threads, events):", launches)
    print("last fill:", dict((k, stats.get(k)) for k in ("sorted", "group_fast", "window", "window_coverage", "deterministic") if stats.get(k) is not None))


if __name__ == "__main__":
    main()
