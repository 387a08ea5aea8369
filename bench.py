#!/usr/bin/env python
"""bench.py -- fill throughput of the B200 fill path (events/s), beside the reference's numpy fill on host cores.

    python bench.py                               # N=1, workload c2 (BASELINE.json configs[1]), 10 steps
    python bench.py --workload c5 --steps 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \
        bench.py --gpus 8 --steps 10 --warmup 3   # one rank per GPU, events sharded, ranks merged with NCCL
    python bench.py --impl reference              # the reference's own CPU path, all host cores

A step = one ``fill`` of the whole workload (every event once) from device-resident columns, plus -- for N>1 --
the rank merge (``Hist.__add__`` across ranks, histbook/hist.py:506-542) as one NCCL collective.  Prints ONE JSON line.
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


def ncu_traffic(workload, n):
    """dram bytes per launch of the fill kernel from this round's `ncu --set full` capture (profiles/traffic.json),
    or None when no capture exists for this workload at this size."""
    path = os.path.join(REPO, "profiles", "traffic.json")
    try:
        with open(path) as f:
            entry = json.load(f).get(workload)
    except (OSError, ValueError):
        return None, None
    if not entry or int(entry["events"]) != int(n):
        return None, None
    return int(entry["dram_bytes_per_launch"]), entry.get("source")


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
    # bounded sample per step: ~1-3 s of CPU work on every core
    per_proc = {"c1": 4 * 10**6, "c2": 4 * 10**6, "c3": 2 * 10**6, "c4": 10**5, "c5": 2 * 10**6}[workload]
    kind, per_rep, total = bench_cpu.run_parallel(workload, per_proc, procs, reps=args.warmup + args.steps)
    timed = per_rep[args.warmup:]
    sec = sum(timed)
    value = total * len(timed) / sec
    line = {
        "impl": "reference", "metric": "fill events/sec", "value": value, "unit": "events/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / len(timed), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
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


def measure(torch, dist, filler, cols, steps, warmup, world, merge):
    """warmup + timed steps; device time by CUDA events on the launch stream, max over ranks."""
    from histbook_b200 import fillable
    launches = 0
    finish = getattr(merge, "finish", lambda: None)
    for _ in range(warmup):
        filler.fill(cols)
        merge()
    finish()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        filler.fill(cols)
        launches += fillable.last_stats.get("launches", 0) + fillable.last_stats.get("keys_launches", 0)
        merge()
    finish()                                       # every step's merge has completed before the clock stops
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


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from histbook_b200 import engine, fillable

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    engine.init(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    device = torch.device("cuda", local)
    hbm, hbm_src = peaks()

    workload = args.workload
    case_name, default_n, bytes_per_event, _ = bench_cpu.WORKLOADS[workload]
    n = int(float(args.events)) if args.events else default_n          # per GPU (weak scaling: every rank gets the full workload size)
    spec = bench_cpu.load_case(case_name)["spec"]
    cols = device_columns(torch, workload, n, 12345 + rank, device)
    filler = fillable.SpecFill(spec)
    filler.fill(dict((k, v[:4096]) for k, v in cols.items()))     # compile + first touch (excluded)
    filler.clear()

    # rank merge = Hist.__add__ over ranks: all-reduce for small contents, reduce-scatter for large ones.
    # One merge per step.  The fill stream only snapshots the bins (device copy, in order with the kernels); the
    # collective runs on a side stream and overlaps the next step's fill; two snapshot buffers alternate, and
    # merge.finish() joins the side stream before the timed region ends.
    merged, snaps, freed = {}, {}, {}
    side = torch.cuda.Stream(device=device) if world > 1 else None
    counter = [0]

    def merge():
        if world == 1:
            return
        parity = counter[0] & 1
        counter[0] += 1
        cur = torch.cuda.current_stream()
        if parity in freed:
            cur.wait_event(freed[parity])               # the collective that read this snapshot two steps ago is done
        items = []
        for i, arena in enumerate(filler.arenas()):
            t = arena.as_tensor()
            snap = snaps.get((i, parity))
            if snap is None or snap.numel() != t.numel():
                snap = snaps[(i, parity)] = torch.empty_like(t)
            snap.copy_(t)
            items.append((i, snap))
        ready = torch.cuda.Event()
        ready.record(cur)
        side.wait_event(ready)
        with torch.cuda.stream(side):
            for i, snap in items:
                if snap.numel() * 8 >= (8 << 20) and snap.numel() % world == 0:
                    out = merged.get(i)
                    if out is None or out.numel() != snap.numel() // world:
                        out = merged[i] = torch.empty(snap.numel() // world, device=device, dtype=torch.float64)
                    dist.reduce_scatter_tensor(out, snap)
                else:
                    dist.all_reduce(snap)
            done = torch.cuda.Event()
            done.record(side)
            freed[parity] = done

    def finish():
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
    merge.finish = finish

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, launches = measure(torch, dist, filler, cols, args.steps, args.warmup, world, merge)
    clocks = sampler.stop() if rank == 0 else None
    main_stats = dict(fillable.last_stats)
    ms_step = ms / args.steps
    value = world * n / (ms_step * 1e-3)

    # roofline of the dominant (only) kernel: algorithmic bytes = bytes of the input columns, each read once
    achieved = n * bytes_per_event / (ms_step * 1e-3) / 1e9
    if world > 1:
        # kernel alone (without the collective) for the roofline
        ms_k, _ = measure(torch, dist, filler, cols, args.steps, 1, world, lambda: None)
        achieved = n * bytes_per_event / (ms_k / args.steps * 1e-3) / 1e9

    # ---- end to end through the public API with HOST (pinned) buffers: H2D of every column + D2H of the result per step
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(torch, dist, workload, spec, n if args.e2e_events is None else int(float(args.e2e_events)), rank, world, device, args)

    extras = {}
    if rank == 0 and world == 1 and args.extras:
        del cols
        torch.cuda.empty_cache()
        for other in ("c1", "c3", "c5", "c4"):
            if other == workload:
                continue
            try:
                extras[other] = run_extra(torch, dist, other, device, hbm)
            except Exception as e:                    # keep the headline line even if an extra fails
                extras[other] = {"error": str(e)[:200]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        procs = os.cpu_count() or 1
        per_proc = {"c1": 4 * 10**6, "c2": 4 * 10**6, "c3": 2 * 10**6, "c4": 10**5, "c5": 2 * 10**6}[workload]
        kind, per_rep, total = bench_cpu.run_parallel(workload, per_proc, procs, reps=3)
        sec = min(per_rep[1:])
        cpu = {"value": total / sec, "unit": "events/s", "cores": procs, "kind": kind,
               "sample": "%d processes x %d events (fill-in-parallel-then-add), best of 2 after 1 warm-up, fill time only" % (procs, per_proc)}

    if rank == 0:
        stats = main_stats
        traffic, traffic_src = ncu_traffic(workload, n)
        line = {
            "metric": "fill events/sec", "value": value, "unit": "events/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: %s" % (workload, bench_cpu.DESCRIPTION[workload]), "events_per_gpu_per_step": n,
                       "bytes_per_event": bytes_per_event, "l2": "inputs (%.1f GB per GPU) larger than L2" % (n * bytes_per_event / 1e9)
                       if n * bytes_per_event > 256e6 else "L2 flushed between steps is NOT done; inputs fit L2",
                       "rank_merge": "none (1 GPU)" if world == 1 else "NCCL all-reduce / reduce-scatter of the bin contents every step, on a side stream overlapping the next step's fill; all merges complete inside the timed region",
                       "kernel": {"regs": stats.get("regs"), "smem_bytes": stats.get("smem_bytes"), "grid": stats.get("grid"),
                                  "blocks_per_sm": stats.get("blocks_per_sm"), "hot_window": stats.get("window"),
                                  "hot_window_coverage_est": stats.get("window_coverage")}},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                         "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                         "traffic_source": traffic_src, "algorithmic_bytes_per_launch": n * bytes_per_event,
                         "kernel": "hb_fill_kernel (generated per book; the only kernel in the timed region)",
                         "peak_source": hbm_src},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }
        if extras:
            line["other_workloads"] = extras
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def api_target(workload, spec):
    """What a user would call: the histbook Hist (or Book) of the workload with the device backend installed
    (``import histbook_b200``) -- or, where the histbook package itself is not importable, the spec-level driver that
    Hist.fill / Book.fill delegate to.  Returns (fill(host_columns), read_result() -> bytes read, description)."""
    from histbook_b200 import _ref, fillable
    if _ref.available():
        import histbook_b200
        histbook_b200.install()
        hb = _ref.load()
        env = dict((n, getattr(hb, n)) for n in ("Hist", "Book", "bin", "intbin", "split", "cut", "profile", "groupby", "groupbin"))
        hists = [eval(src, dict(env)) for src in bench_cpu.load_case(bench_cpu.WORKLOADS[workload][0])["hists"]]
        target = hists[0] if len(hists) == 1 else hb.Book(dict(("h%03d" % i, h) for i, h in enumerate(hists)))

        def read():
            total = 0
            for h in hists:
                c = h._content                      # device -> host copy, as proj/export/vega would trigger
                total += c.nbytes if isinstance(c, numpy.ndarray) else sum(v.nbytes for v in c.values())
            return total
        return (lambda cols: target.fill(**cols)), read, "histbook %s.fill(**host numpy columns) + ._content (backend installed by `import histbook_b200`)" % type(target).__name__
    filler = fillable.SpecFill(spec)

    def read():
        return sum(c.nbytes for c in filler.contents() if isinstance(c, numpy.ndarray))
    return filler.fill, read, "SpecFill.fill(host numpy columns) + contents()"


def run_e2e(torch, dist, workload, spec, n, rank, world, device, args):
    """Same metric through the user-facing call with host buffers: Hist.fill(x=<numpy>, ...) then read hist content."""
    host = {}
    h2d = 0
    src = device_columns(torch, workload, n, 777 + rank, device)
    for k, v in src.items():
        pinned = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
        pinned.copy_(v)
        host[k] = pinned.numpy()
        h2d += pinned.numel() * pinned.element_size()
    del src
    torch.cuda.empty_cache()
    fill, read, api = api_target(workload, spec)
    steps = max(1, min(args.steps, 3))
    d2h = 0
    for it in range(1 + steps):
        if it == 1:
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
        fill(host)
        d2h = read()                                # device -> host read of the step's result
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([sec], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    return {"value": world * n * steps / sec, "unit": "events/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "steps": steps, "events_per_gpu_per_step": n, "api": api,
            "h2d_gbs_per_gpu": h2d * steps / sec / 1e9, "bound": "PCIe host->device copy of the columns"}


def run_extra(torch, dist, workload, device, hbm):
    from histbook_b200 import fillable
    case_name, default_n, bytes_per_event, _ = bench_cpu.WORKLOADS[workload]
    n = default_n
    if workload == "c4":
        n = 10**8
    spec = bench_cpu.load_case(case_name)["spec"]
    cols = device_columns(torch, workload, n, 12345, device)
    filler = fillable.SpecFill(spec)
    filler.fill(dict((k, v[:4096]) for k, v in cols.items()))
    filler.clear()
    steps = 3 if n >= 10**9 or workload == "c4" else 10
    if n * bytes_per_event <= 256e6:
        # small input: flush L2 between steps, time each step separately
        from histbook_b200 import engine
        times = []
        for it in range(3 + steps):
            engine.flush_l2()
            filler.fill(cols, timed=True)          # CUDA events recorded around the launch inside hb_fill (stream 0)
            if it >= 3:
                times.append(fillable.last_stats["ms_kernel"])
        ms_step = sum(times) / len(times)
        l2 = "L2 flushed between steps; kernel time from CUDA events around the launch (host launch overhead excluded)"
    else:
        ms, _ = measure(torch, dist, filler, cols, steps, 3, 1, lambda: None)
        ms_step = ms / steps
        l2 = "inputs larger than L2"
    gbs = n * bytes_per_event / (ms_step * 1e-3) / 1e9
    del cols
    torch.cuda.empty_cache()
    return {"events": n, "ms_per_step": ms_step, "events_per_s": n / (ms_step * 1e-3), "achieved_gbs": gbs, "frac": gbs / hbm,
            "l2": l2, "kernel": dict((k, fillable.last_stats.get(k)) for k in ("regs", "smem_bytes", "grid", "blocks_per_sm"))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(bench_cpu.WORKLOADS))
    ap.add_argument("--events", default=None, help="events per GPU per step (default: the workload's BASELINE size)")
    ap.add_argument("--e2e-events", default=None)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--extras", action="store_true", help="also time the other BASELINE configs (N=1 only)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
