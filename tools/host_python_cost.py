#!/usr/bin/env python
"""Python cost of one ``SpecFill.fill`` call WITHOUT a GPU: the engine's device entry points are swapped for the test emulator
(tests/emul_engine.py), the first fill runs for real (keys, stores and kernels exist afterwards), then the launch is a no-op and
what is timed is the host path alone (column binding, plan and kernel look-up, store binding, argument marshalling).
    python tools/host_python_cost.py c1      # 14 us     python tools/host_python_cost.py c4      # 160 us (64 histograms)
The ctypes call and the driver's launch come on top on a real device (tools/host_overhead.py measures the whole on the GPU)."""
import sys, json, time, cProfile, pstats
import os
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy
import emul_engine
from histbook_b200 import fillable, engine
cases = dict((c["name"], c) for c in json.load(open(os.path.join(REPO, "tests", "golden", "cases.json")))["cases"])
which = sys.argv[1] if len(sys.argv) > 1 else "c4"
spec = cases[which]["spec"]
rs = numpy.random.RandomState(7)
n = 3000
arrays = {"x": rs.normal(0, 1, n), "y": rs.normal(0, 1, n), "z": rs.normal(0, 1, n), "w": rs.uniform(0.5, 1.5, n), "n": rs.poisson(3, n).astype(numpy.int64)}
with emul_engine.installed():
    fillable.set_options(fx=False, window=False, box=False)
    sf = fillable.SpecFill(spec)
    mine = dict((k, arrays[k]) for k in sf.fields)
    sf.fill(mine)
    real_fill = engine.fill
    # a no-op launch: what is left is the host path; the key set keeps the keys of the real first fill
    keys_state = {}
    def noop(program, columns, n, arenas, aux=(), win=(), timed=False, offset=0, exact=None):
        return {"events": int(n), "h2d_bytes": 0, "launches": 1, "grid": 1, "blocks_per_sm": 1, "ms_kernel": 0.0}
    class KS(emul_engine.KeySet):
        def clear(self): pass
    if sf.plan.keysets:
        for ks in sf.plan.keysets: ks.__class__ = KS
    engine.fill = noop
    t0 = time.perf_counter()
    N = 300
    for _ in range(N): sf.fill(mine)
    dt = (time.perf_counter() - t0) / N
    print("%s: %.1f us of Python per SpecFill.fill" % (which, dt * 1e6))
    pr = cProfile.Profile(); pr.enable()
    for _ in range(N): sf.fill(mine)
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
    fillable.set_options(fx=None, window=None, box=None)
