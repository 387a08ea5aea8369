#!/usr/bin/env python
"""bench.py -- fill throughput of the B200 fill path (events/s), beside the reference's numpy fill on host cores.

    python bench.py                               # N=1: the north-star config c4 (64-histogram Book, 1e9 events) through Book.fill
    python bench.py --workload c5 --steps 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \
        bench.py --gpus 8 --steps 10 --warmup 3   # one rank per GPU; the SAME total events sharded over the ranks (strong scaling)
    python bench.py --impl reference              # the reference's own CPU path, all host cores

A step = one ``Book.fill`` / ``Hist.fill`` (the histbook API, backend installed by ``import histbook_b200``) of this
rank's shard of the workload from device-resident columns, plus -- for N>1 -- the rank merge through the product's
``histbook_b200.dist.Merger`` (the distributed ``Hist.__add__``, histbook/hist.py:506-542): ONE NCCL collective per step for
the whole book.  Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import numpy                                        # noqa: E402

import bench_cpu                                    # noqa: E402

CPU_SAMPLE = {"c1": 4 * 10**6, "c2": 4 * 10**6, "c3": 2 * 10**6, "c4": 10**5, "c5": 2 * 10**6}     # events per process and step


def issue_roofline(workload):
    """Big books are bound by instruction issue, not by HBM: events/s at 4 warp instructions per clock and SM, from the warp
    instructions per event of this round's ncu capture (profiles/traffic.json)."""
    path = os.path.join(REPO, "profiles", "traffic.json")
    try:
        with open(path) as f:
            entry = json.load(f).get(workload) or {}
    except (OSError, ValueError):
        return None
    if "warp_inst_per_event" not in entry:
        return None
    out = {"bound": "issue", "warp_inst_per_event": entry["warp_inst_per_event"], "peak_events_per_s": 148 * 4 * 1.965e9 / entry["warp_inst_per_event"],
           "source": entry.get("source"), "note": "148 SMs x 4 issue slots x 1.965 GHz / warp instructions per event; the HBM figure above is what the metric asks for"}
    if "smem_wavefronts_per_event" in entry:
        # the other pipe the kernel sits on: one shared-memory wavefront per clock and SM
        out["smem_pipe"] = {"wavefronts_per_event": entry["smem_wavefronts_per_event"], "peak_events_per_s": 148 * 1.965e9 / entry["smem_wavefronts_per_event"]}
    return out


def ncu_traffic(workload, n):
    """dram bytes per launch of the fill kernel from this round's `ncu --set full` capture (profiles/traffic.json),
    scaled to this launch's events when the capture was taken at another size (the traffic is the input columns, O(N))."""
    path = os.path.join(REPO, "profiles", "traffic.json")
    try:
        with open(path) as f:
            entry = json.load(f).get(workload)
    except (OSError, ValueError):
        return None, None
    if not entry:
        return None, None
    events = int(entry["events"])
    if events == int(n):
        return int(entry["dram_bytes_per_launch"]), entry.get("source")
    return int(entry["dram_bytes_per_launch"] * (float(n) / events)), "%s, captured at %d events and scaled by N" % (entry.get("source"), events)


def peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------- clocks

class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in self.samples:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(numpy.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- reference arm

def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workload = args.workload
    procs = os.cpu_count() or 1
    per_proc = CPU_SAMPLE[workload]                 # bounded sample per step: ~1-3 s of CPU work on every core
    kind, per_rep, total = bench_cpu.run_parallel(workload, per_proc, procs, reps=args.warmup + args.steps)
    timed = per_rep[args.warmup:]
    sec = sum(timed)
    value = total * len(timed) / sec
    line = {
        "impl": "reference", "metric": "fill events/sec", "value": value, "unit": "events/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / len(timed), "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %s" % (workload, bench_cpu.DESCRIPTION[workload]),
                   "events_per_step": total, "note": "bounded sample of the workload; fill is O(N)"},
        "cpu_baseline": {"value": value, "unit": "events/s", "cores": procs, "kind": kind,
                         "sample": "%d processes x %d events per step (fill-in-parallel-then-add, README.rst:29), fill time only" % (procs, per_proc)},
        "e2e": {"value": value, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------- GPU arm

def device_columns(torch, workload, n, seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    cols = {}
    for name in bench_cpu.WORKLOADS[workload][3]:
        if name == "w":
            cols[name] = torch.rand(n, generator=g, device=device, dtype=torch.float64) + 0.5
        elif name == "n":
            cols[name] = torch.poisson(torch.full((n,), 3.0, device=device, dtype=torch.float32), generator=g).to(torch.int64)
        else:
            cols[name] = torch.randn(n, generator=g, device=device, dtype=torch.float64)
    return cols


class Target(object):
    """What a user calls: the histbook Hist (or Book of Hists) of a workload with the device backend installed by
    ``import histbook_b200`` -- or, where the histbook package itself is not importable, the spec-level driver that
    Hist.fill / Book.fill delegate to."""

    def __init__(self, workload):
        from histbook_b200 import _ref, fillable
        case = bench_cpu.load_case(bench_cpu.WORKLOADS[workload][0])
        self.spec = case["spec"]
        if _ref.available():
            import histbook_b200
            histbook_b200.install()
            hb = _ref.load()
            env = dict((n, getattr(hb, n)) for n in ("Hist", "Book", "bin", "intbin", "split", "cut", "profile", "groupby", "groupbin"))
            self.hists = [eval(src, dict(env)) for src in case["hists"]]
            self.obj = self.hists[0] if len(self.hists) == 1 else hb.Book(dict(("h%03d" % i, h) for i, h in enumerate(self.hists)))
            self.api = "histbook %s.fill(**columns) with the backend installed by `import histbook_b200`" % type(self.obj).__name__
            self.fill = lambda cols: self.obj.fill(**cols)
        else:
            self.obj = fillable.SpecFill(self.spec)
            self.hists = self.obj.hists
            self.api = "fillable.SpecFill.fill(columns) (histbook itself not importable here)"
            self.fill = self.obj.fill

    def clear(self):
        for h in self.hists:
            h.clear()

    def read(self):
        """Device -> host read of every histogram's bins, as proj/export/vega would trigger; returns the bytes read."""
        total = 0
        for h in self.hists:
            c = h._content if hasattr(h, "_content") else h.content
            total += content_bytes(c)
        return total

    def peek(self, i):
        st = self.hists[i].__dict__.get("_hb")
        return None if st is None else st.peek()


def content_bytes(c):
    if c is None:
        return 0
    if isinstance(c, numpy.ndarray):
        return c.nbytes
    return sum(content_bytes(v) for v in c.values())


def content_sum(c):
    if c is None:
        return 0.0
    if isinstance(c, numpy.ndarray):
        return float(c.sum())
    return sum(content_sum(v) for v in c.values())


def measure(torch, dist, target, cols, steps, warmup, world, merger):
    """warmup + timed steps; device time by CUDA events on the launch stream, max over ranks."""
    from histbook_b200 import fillable
    launches = 0
    for _ in range(warmup):
        target.fill(cols)
        if merger is not None:
            merger.start()
    if merger is not None:
        merger.wait()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        target.fill(cols)
        launches += fillable.last_stats.get("launches", 0) + fillable.last_stats.get("keys_launches", 0)
        if merger is not None:
            merger.start()                         # one snapshot + one NCCL collective for the whole book, on a side stream
    if merger is not None:
        merger.wait()                              # every step's merge has completed before the clock stops
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, launches


def check_merge(torch, dist, target, merger, world, fills, n_rank, workload):
    """Outside the timed region: the merged book must be the sum of the ranks' books -- every histogram's total against
    an independent all-reduce of the rank-local totals, and the unweighted histogram's total count bit-exact against
    the number of events all ranks filled."""
    merged = merger.start().contents()
    local = numpy.array([content_sum(target.peek(i)) for i in range(len(target.hists))], dtype=numpy.float64)
    t = torch.from_numpy(local.copy()).cuda()
    dist.all_reduce(t)
    expect = t.cpu().numpy()
    got = numpy.array([content_sum(c) for c in merged])
    if not numpy.all(numpy.abs(got - expect) <= 1e-9 * numpy.maximum(numpy.abs(expect), 1.0)):
        raise SystemExit("bench: merged histogram totals differ from the sum of the rank totals: %r vs %r" % (got.tolist()[:8], expect.tolist()[:8]))
    counts = {"c1": 0, "c4": 0}.get(workload)      # an unweighted histogram with every flow bin: each event lands in exactly one bin
    if counts is not None:
        n_all = torch.tensor([n_rank], device="cuda", dtype=torch.int64)
        dist.all_reduce(n_all)
        if got[counts] != float(int(n_all.item()) * fills):
            raise SystemExit("bench: merged count %r != events filled by all ranks %d" % (got[counts], int(n_all.item()) * fills))
    return {"histograms": len(merged), "totals_checked": int(len(got)), "count_exact": counts is not None,
            "collectives_per_merge": merger.collectives_last}


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from histbook_b200 import dist as hbdist, engine, fillable

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    engine.init(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"          # (a bare "NCCL version ..." line on stdout is not part of the one JSON line)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    device = torch.device("cuda", local)
    hbm, hbm_src = peaks()

    workload = args.workload
    case_name, default_n, bytes_per_event, _ = bench_cpu.WORKLOADS[workload]
    total_n = int(float(args.events)) if args.events else default_n
    strong = args.scaling == "strong"
    lo, hi = hbdist.shard(total_n, rank, world) if strong else (0, total_n)
    n = hi - lo                                                   # events this rank fills per step
    n_all = total_n if strong else total_n * world

    def setup(n_rank):
        cols = device_columns(torch, workload, n_rank, 12345 + rank, device)
        target = Target(workload)
        target.fill(dict((k, v[:4096]) for k, v in cols.items()))     # compile + first touch (excluded)
        target.clear()
        merger = hbdist.Merger(target.obj) if world > 1 else None
        return cols, target, merger

    cols, target, merger = setup(n)
    api = target.api
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, launches = measure(torch, dist, target, cols, args.steps, args.warmup, world, merger)
    clocks = sampler.stop() if rank == 0 else None
    main_stats = dict(fillable.last_stats)
    ms_step = ms / args.steps
    value = n_all / (ms_step * 1e-3)
    merge_check = None
    if world > 1:
        merge_check = check_merge(torch, dist, target, merger, world, args.warmup + args.steps, n, workload)

    # roofline of the dominant kernel: algorithmic bytes = bytes of the input columns, each read once
    if world > 1:
        ms_k, _ = measure(torch, dist, target, cols, args.steps, 1, world, None)      # fill alone (without the collective)
    else:
        ms_k = ms
    achieved = n * bytes_per_event / (ms_k / args.steps * 1e-3) / 1e9
    keys_note = ""
    if main_stats.get("keys_launches"):
        keys_note = " (+ the key-discovery kernel of the group axes, which reads the group columns once more; its time is inside the step)"

    # the other scaling mode beside it (N>1): every rank fills the whole workload size
    other_scaling = None
    if world > 1 and not args.no_weak:
        cols = target = merger = None
        torch.cuda.empty_cache()
        n2 = total_n if strong else hbdist.shard(total_n, rank, world)[1] - hbdist.shard(total_n, rank, world)[0]
        cols, target, merger = setup(n2)
        ms2, _ = measure(torch, dist, target, cols, args.steps, args.warmup, world, merger)
        t = torch.tensor([n2], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        other_scaling = {"scaling": "weak" if strong else "strong", "events_per_gpu_per_step": n2, "ms_per_step": ms2 / args.steps,
                         "value": int(t.item()) / (ms2 / args.steps * 1e-3), "unit": "events/s"}

    # ---- end to end through the public API with HOST buffers: H2D of every column + D2H of the result per step
    e2e = None
    cols = target = merger = None
    torch.cuda.empty_cache()
    if not args.no_e2e:
        e2e = run_e2e(torch, dist, workload, n if args.e2e_events is None else int(float(args.e2e_events)), rank, world, device, args)

    extras = {}
    if rank == 0 and world == 1 and not args.no_extras:
        for other in ("c1", "c2", "c3", "c5", "c4"):
            if other == workload:
                continue
            try:
                extras[other] = run_extra(torch, dist, other, device, hbm)
            except Exception as e:                    # keep the headline line even if an extra fails
                extras[other] = {"error": str(e)[:200]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        procs = os.cpu_count() or 1
        per_proc = CPU_SAMPLE[workload]
        kind, per_rep, total = bench_cpu.run_parallel(workload, per_proc, procs, reps=3)
        sec = min(per_rep[1:])
        cpu = {"value": total / sec, "unit": "events/s", "cores": procs, "kind": kind,
               "sample": "%d processes x %d events (fill-in-parallel-then-add), best of 2 after 1 warm-up, fill time only" % (procs, per_proc)}

    if rank == 0:
        stats = main_stats
        traffic, traffic_src = ncu_traffic(workload, n)
        line = {
            "metric": "fill events/sec", "value": value, "unit": "events/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: %s" % (workload, bench_cpu.DESCRIPTION[workload]), "events_per_step_all_gpus": n_all,
                       "events_per_gpu_per_step": n, "bytes_per_event": bytes_per_event, "api": api + " on device-resident torch columns",
                       "l2": "inputs (%.1f GB per GPU) larger than L2" % (n * bytes_per_event / 1e9)
                       if n * bytes_per_event > 256e6 else "inputs fit L2 and L2 is NOT flushed between steps",
                       "rank_merge": "none (1 GPU)" if world == 1 else "histbook_b200.dist.Merger: one snapshot + ONE NCCL collective for the whole book every step (plus one int64 key all-gather for group axes), on a side stream overlapping the next step's fill; all merges complete inside the timed region",
                       "merge_check": merge_check,
                       "kernel": {"regs": stats.get("regs"), "smem_bytes": stats.get("smem_bytes"), "grid": stats.get("grid"),
                                  "blocks_per_sm": stats.get("blocks_per_sm"), "hot_window": stats.get("window"),
                                  "hot_window_coverage_est": stats.get("window_coverage")}},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                         "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                         "traffic_source": traffic_src, "algorithmic_bytes_per_launch": n * bytes_per_event,
                         "kernel": "hb_fill_kernel (generated per book)" + keys_note,
                         "peak_source": hbm_src},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }
        second = issue_roofline(workload)
        if second is not None:
            second["frac"] = (n / (ms_k / args.steps * 1e-3)) / second["peak_events_per_s"]
            if "smem_pipe" in second:
                second["smem_pipe"]["frac"] = (n / (ms_k / args.steps * 1e-3)) / second["smem_pipe"]["peak_events_per_s"]
            line["roofline"]["second_bound"] = second
        if other_scaling:
            line["other_scaling"] = other_scaling
        if extras:
            line["other_workloads"] = extras
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def host_columns(torch, workload, n, seed, device, pinned):
    """Plain numpy arrays in pageable host memory (what ``numpy.random`` gives a user) -- or page-locked ones."""
    host, nbytes = {}, 0
    chunk = 1 << 26
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    for name in bench_cpu.WORKLOADS[workload][3]:
        dtype = numpy.int64 if name == "n" else numpy.float64
        if pinned:
            out_t = torch.empty(n, dtype=torch.int64 if name == "n" else torch.float64, pin_memory=True)
            out = out_t.numpy()
        else:
            out = numpy.empty(n, dtype=dtype)
        for a in range(0, n, chunk):
            m = min(chunk, n - a)
            if name == "w":
                part = torch.rand(m, generator=g, device=device, dtype=torch.float64) + 0.5
            elif name == "n":
                part = torch.poisson(torch.full((m,), 3.0, device=device, dtype=torch.float32), generator=g).to(torch.int64)
            else:
                part = torch.randn(m, generator=g, device=device, dtype=torch.float64)
            out[a:a + m] = part.cpu().numpy()
        host[name] = out
        nbytes += out.nbytes
    return host, nbytes


def run_e2e(torch, dist, workload, n, rank, world, device, args):
    """Same metric through the user-facing call with host buffers: Book.fill(x=<numpy>, ...) then read every histogram.
    Headline: plain (pageable) numpy arrays; page-locked arrays beside it."""
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 64 << 30
    bpe = bench_cpu.WORKLOADS[workload][2]
    n = int(max(1, min(n, 0.35 * avail / world / bpe)))
    out = {}
    steps = max(1, min(args.steps, 3))
    for kind in ("pageable", "pinned"):
        m = n if kind == "pageable" else max(1, min(n, (8 << 30) // bpe))        # pinning tens of GB takes longer than the measurement
        host, h2d = host_columns(torch, workload, m, 777 + rank, device, kind == "pinned")
        torch.cuda.empty_cache()
        target = Target(workload)
        d2h = 0
        for it in range(1 + steps):
            if it == 1:
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                t0 = time.perf_counter()
            target.fill(host)
            d2h = target.read()                         # device -> host read of the step's result
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([sec], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        out[kind] = {"value": world * m * steps / sec, "events_per_gpu_per_step": m, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                     "h2d_gbs_per_gpu": h2d * steps / sec / 1e9}
        api = target.api
        del host, target
    main = out["pageable"]
    return {"value": main["value"], "unit": "events/s", "h2d_bytes_per_step": main["h2d_bytes_per_step"], "d2h_bytes_per_step": main["d2h_bytes_per_step"],
            "steps": steps, "events_per_gpu_per_step": main["events_per_gpu_per_step"], "api": api + " on plain numpy arrays in pageable host memory, then ._content of every histogram",
            "h2d_gbs_per_gpu": main["h2d_gbs_per_gpu"], "host_memory": "pageable (numpy.empty)", "bound": "host->device copy of the columns",
            "pinned": dict(out["pinned"], host_memory="page-locked (torch pin_memory)")}


def run_extra(torch, dist, workload, device, hbm):
    from histbook_b200 import engine, fillable
    case_name, default_n, bytes_per_event, _ = bench_cpu.WORKLOADS[workload]
    n = default_n
    cols = device_columns(torch, workload, n, 12345, device)
    target = Target(workload)
    target.fill(dict((k, v[:4096]) for k, v in cols.items()))
    target.clear()
    steps = 3 if n >= 10**9 else 10
    if n * bytes_per_event <= 256e6:
        # small input: flush L2 between steps, time each step separately
        times, calls = [], []
        fillable.TIMED = True                       # CUDA events around the launch inside hb_fill: the kernel, without the host's launch path
        try:
            for it in range(3 + steps):
                engine.flush_l2()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                target.fill(cols)
                e1.record()
                torch.cuda.synchronize()
                if it >= 3:
                    times.append(fillable.last_stats["ms_kernel"])
                    calls.append(e0.elapsed_time(e1))
        finally:
            fillable.TIMED = False
        ms_step = sum(times) / len(times)
        l2 = "L2 flushed between steps; kernel time from CUDA events around the launch inside hb_fill; the whole Hist.fill call (Python + ctypes + launch + kernel) took %.1f us on the device timeline" % (1e3 * sum(calls) / len(calls))
    else:
        ms, _ = measure(torch, dist, target, cols, steps, 3, 1, None)
        ms_step = ms / steps
        l2 = "inputs larger than L2"
    gbs = n * bytes_per_event / (ms_step * 1e-3) / 1e9
    stats = dict(fillable.last_stats)
    del cols
    torch.cuda.empty_cache()
    traffic, traffic_src = ncu_traffic(workload, n)
    return {"workload": "%s: %s" % (workload, bench_cpu.DESCRIPTION[workload]), "events": n, "ms_per_step": ms_step, "events_per_s": n / (ms_step * 1e-3),
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": n * bytes_per_event},
            "l2": l2, "kernel": dict((k, stats.get(k)) for k in ("regs", "smem_bytes", "grid", "blocks_per_sm", "window", "window_coverage"))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(bench_cpu.WORKLOADS))
    ap.add_argument("--events", default=None, help="events per step over ALL GPUs (strong) / per GPU (weak); default: the workload's BASELINE size")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default, BASELINE.md section 4): the same total events sharded over the GPUs; weak: every GPU fills the whole size")
    ap.add_argument("--e2e-events", default=None)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-weak", action="store_true", help="N>1: skip the second run in the other scaling mode")
    ap.add_argument("--no-extras", action="store_true", help="N=1: skip the other BASELINE configs (other_workloads)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
