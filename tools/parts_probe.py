#!/usr/bin/env python
"""Do the two half-book kernels of c4 share the SMs?  Times one half alone, and both halves side by side."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, bench_cpu
from histbook_b200 import engine, fillable

engine.init(0)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**8
dev = torch.device("cuda", 0)
cols = bench.device_columns(torch, "c4", n, 1, dev)
spec = bench_cpu.load_case("c4")["spec"]


def timed(sf, reps=3):
    sf.fill(dict((k, v[:4096]) for k, v in cols.items()))
    sf.fill(cols)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        sf.fill(cols)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    sf.fill(cols, timed=True)
    return ms, dict(fillable.last_stats)

half = dict(spec, hists=spec["hists"][:32])
for label, sp, opts in (("both halves, parts=2", spec, {}), ("single kernel, parts=1", spec, {"parts": 1}),
                        ("first half alone, 512 threads x 1 block/SM", half, {"parts": 1, "sort_threads": 512, "sort_min_blocks": 2, "max_blocks_per_sm": 1}),
                        ("first half alone, 512 threads x 2 blocks/SM", half, {"parts": 1, "sort_threads": 512, "sort_min_blocks": 2}),
                        ("first half alone, 512 x 2 blocks/SM, stagger 20 us", half, {"parts": 1, "sort_threads": 512, "sort_min_blocks": 2, "sort_stagger_us": 20}),
                        ("first half alone, 256 threads x 4 blocks/SM", half, {"parts": 1, "sort_threads": 256, "sort_min_blocks": 4}),
                        ("first half alone, 1024 threads", half, {"parts": 1})):
    fillable.set_options(**opts)
    ms, st = timed(fillable.SpecFill(sp))
    fillable.set_options(**dict((k, None) for k in opts))
    print("%-52s %7.2f ms/fill, kernels alone %7.2f ms = %6.2f Gev/s  grid %s parts %s regs %s smem %s" % (
        label, ms, st.get("ms_kernel", 0), n / max(st.get("ms_kernel", 0), 1e-9) / 1e6, st.get("grid"), st.get("parts"), st.get("regs"), st.get("smem_bytes")))
