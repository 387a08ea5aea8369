"""Where the fixed cost of one fill goes: per-block globaltimer stamps (option block_times) of the fused kernel.
Usage (GPU box): python tools/block_times.py [c2] [events]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

import numpy  # noqa: E402
import torch  # noqa: E402

import golden_util  # noqa: E402
from histbook_b200 import engine, fillable  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10**8
    fillable.set_options(block_times=True)
    spec = golden_util.load()[name]["spec"]
    sf = fillable.SpecFill(spec)
    g = torch.Generator(device="cuda")
    g.manual_seed(12345)
    cols = {}
    for c in sf.plan.fields:
        if c == "w":
            cols[c] = torch.rand(n, generator=g, device="cuda", dtype=torch.float64) + 0.5
        elif c == "n":
            cols[c] = torch.poisson(torch.full((n,), 3.0, device="cuda"), generator=g).to(torch.int64)
        else:
            cols[c] = torch.randn(n, generator=g, device="cuda", dtype=torch.float64)
    for rep in range(6):
        sf.fill(cols, timed=True)
        st = dict(fillable.last_stats)
        t = numpy.array(st["block_times"], dtype=numpy.float64)
        t0 = t[:, 0].min()
        start, loop, end = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3, (t[:, 2] - t0) / 1e3
        print("%s n=%g rep %d: kernel %.1f us (events) | block start max %.1f | loop end min %.1f med %.1f max %.1f | "
              "merge per block med %.1f max %.1f | kernel end %.1f  [us after the first block started]"
              % (name, n, rep, st["ms_kernel"] * 1e3, start.max(), loop.min(), numpy.median(loop), loop.max(),
                 numpy.median(end - loop), (end - loop).max(), end.max()))
    order = numpy.argsort(loop)
    print("slowest blocks (loop end):", [(int(b), round(float(loop[b]), 1)) for b in order[-6:]])
    print("fastest blocks (loop end):", [(int(b), round(float(loop[b]), 1)) for b in order[:6]])


if __name__ == "__main__":
    main()
