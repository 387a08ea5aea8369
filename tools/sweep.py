#!/usr/bin/env python
"""Time kernel-generation options against each other on the GPU (tuning aid, not a benchmark of record).

    python tools/sweep.py c2 '{"window": false}' '{"window": true, "win_threads": 512}' ...
"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import torch                                          # noqa: E402

import bench                                          # noqa: E402
import bench_cpu                                      # noqa: E402
from histbook_b200 import engine, fillable            # noqa: E402


def main():
    workload = sys.argv[1]
    n = None
    flush = False
    variants = []
    for a in sys.argv[2:]:
        if a == "flush":
            flush = True              # small inputs: flush L2 before every fill, time the launch alone (hb_fill's own events)
        elif a.startswith("n="):
            n = int(float(a[2:]))
        else:
            variants.append(json.loads(a))
    if not variants:
        variants = [{}]
    engine.init(0)
    if workload.startswith("hist:"):
        # ad-hoc: tools/sweep.py 'hist:Hist(bin("x",100,-5,5), weight="w")' n=1e8 ...   (needs histbook importable)
        import histbook_b200
        from histbook_b200 import spec as hbspec
        hb = histbook_b200.reference()
        env = dict((k, getattr(hb, k)) for k in ("Hist", "bin", "intbin", "split", "cut", "profile", "groupby", "groupbin"))
        hists = [eval(src, dict(env)) for src in workload[5:].split(";;")]
        spec = hbspec.book_spec(hists)
        fields = hbspec.spec_fields(spec)
        bench_cpu.WORKLOADS[workload] = (None, 10**8, 8 * len(fields), tuple(fields))
        n = n or 10**8
        bpe = 8 * len(fields)
    else:
        case, default_n, bpe, _ = bench_cpu.WORKLOADS[workload]
        n = n or default_n
        spec = bench_cpu.load_case(case)["spec"]
    cols = bench.device_columns(torch, workload, n, 12345, torch.device("cuda", 0))
    for opts in variants:
        fillable._options.clear()
        fillable.set_options(**opts)
        f = fillable.SpecFill(spec)
        f.fill(dict((k, v[:4096]) for k, v in cols.items()))
        f.clear()
        for _ in range(3):
            f.fill(cols)
        torch.cuda.synchronize()
        steps = 5
        if flush:
            times = []
            for _ in range(steps + 5):
                engine.flush_l2()
                f.fill(cols, timed=True)
                times.append(fillable.last_stats["ms_kernel"])
            ms = sum(sorted(times)[:steps]) / steps
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                f.fill(cols)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
        st = fillable.last_stats
        print("%s %-60s %9.3f ms  %8.2f Gev/s  %7.1f GB/s (%.1f%%)  regs=%s smem=%s grid=%s occ=%s win=%s cov=%s" % (
            workload[:40], json.dumps(opts), ms, n / ms * 1e-6, n * bpe / ms * 1e-6, 100 * n * bpe / ms * 1e-6 / 6551,
            st.get("regs"), st.get("smem_bytes"), st.get("grid"), st.get("blocks_per_sm"), st.get("window"), st.get("window_coverage")), flush=True)


if __name__ == "__main__":
    main()
