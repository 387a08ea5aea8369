#!/usr/bin/env python
"""Time Hist.project of the c5 histogram (203^3 x 2 float64 = 134 MB) on the device against the reference's route
(device -> host copy of the content, numpy.sum on the host)."""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import numpy                                          # noqa: E402
import torch                                          # noqa: E402

import histbook_b200                                  # noqa: E402
from histbook import Hist, bin                        # noqa: E402

n = 10**8
g = torch.Generator(device="cuda")
g.manual_seed(1)
cols = dict((k, torch.randn(n, generator=g, device="cuda", dtype=torch.float64)) for k in "xyz")
cols["w"] = torch.rand(n, generator=g, device="cuda", dtype=torch.float64) + 0.5
h = Hist(bin("x", 200, -5, 5), bin("y", 200, -5, 5), bin("z", 200, -5, 5), weight="w")
h.fill(**cols)
torch.cuda.synchronize()
for keep in (("x",), ("x", "y")):
    h.project(*keep)._content
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    p = h.project(*keep)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    got = p._content
    t2 = time.perf_counter()
    host = histbook_b200.peek(h)                      # device -> host copy of 134 MB
    t3 = time.perf_counter()
    exp = numpy.sum(host, tuple(i for i, a in enumerate("xyz") if a not in keep))
    t4 = time.perf_counter()
    assert numpy.allclose(got, exp, rtol=1e-12, atol=0)
    print("project%s: device %.3f ms (+ %.3f ms to read the %d-byte result)   host route: copy %.1f ms + numpy.sum %.1f ms" % (
        keep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), got.nbytes, 1e3 * (t3 - t2), 1e3 * (t4 - t3)))
