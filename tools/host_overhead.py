"""Host-side cost of one fill call (Python + ctypes + launch) on small device-resident columns, with a cProfile listing.
Usage (on the GPU box): python tools/host_overhead.py [c1|c2|c3] [events]"""
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

import torch  # noqa: E402

import golden_util  # noqa: E402
from histbook_b200 import engine, fillable  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 4096
    spec = golden_util.load()[name]["spec"]
    sf = fillable.SpecFill(spec)
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    cols = dict((c, torch.randn(n, generator=g, device="cuda", dtype=torch.float64).abs() + 0.5) for c in sf.plan.fields)
    for _ in range(200):
        sf.fill(cols)
    engine.sync()
    reps = 5000
    t0 = time.perf_counter()
    for _ in range(reps):
        sf.fill(cols)
    t1 = time.perf_counter()
    engine.sync()
    t2 = time.perf_counter()
    print("%s n=%d: %.1f us per fill call on the host (%.1f us incl. drain)" % (name, n, (t1 - t0) / reps * 1e6, (t2 - t0) / reps * 1e6))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(reps):
        sf.fill(cols)
    pr.disable()
    engine.sync()
    out = io.StringIO()
    pstats.Stats(pr, stream=out).sort_stats("tottime").print_stats(22)
    print(out.getvalue())


if __name__ == "__main__":
    main()
