"""CPU side of the benchmark: the reference's own numpy fill (or the oracle port) on the box's host cores.

Kept free of CUDA / torch imports so that worker processes can be spawned safely.  Used only by bench.py's
``cpu_baseline`` leg and ``--impl reference`` arm (the one place outside tests/ allowed to execute oracle/).
"""
import json
import os
import sys
import time
import warnings

import numpy

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

WORKLOADS = {
    # name: (golden case holding the spec, default events, bytes per event, columns)
    "c1": ("c1", 10**7, 8, ("x",)),
    "c2": ("c2", 10**8, 24, ("w", "x", "y")),
    "c3": ("c3", 10**8, 16, ("x", "y")),
    "c4": ("c4", 10**9, 40, ("n", "w", "x", "y", "z")),
    "c5": ("c5", 10**9, 32, ("w", "x", "y", "z")),
}
DESCRIPTION = {
    "c1": 'Hist(bin("x",100,-5,5)) unweighted',
    "c2": 'Hist(bin("x+y",200,-5,5), bin("sqrt(x**2+y**2)",200,0,5), weight="w") 2-D weighted with sumw2',
    "c3": 'Hist(bin("x",1000,-5,5), profile("y"), profile("y**2"), cut("x>0"))',
    "c4": "Book of 64 histograms (SURVEY.md App. D) filled in one pass",
    "c5": 'Hist(bin x,y,z 200^3, weight="w") 3-D weighted, 8M bins',
}


def load_case(name):
    with open(os.path.join(REPO, "tests", "golden", "cases.json")) as f:
        for c in json.load(f)["cases"]:
            if c["name"] == name:
                return c
    raise KeyError(name)


def make_columns(workload, n, seed):
    """Synthetic columns of SURVEY.md section 8d on the host (numpy)."""
    rs = numpy.random.RandomState(seed)
    cols = {}
    for name in WORKLOADS[workload][3]:
        if name == "w":
            cols[name] = rs.uniform(0.5, 1.5, n)
        elif name == "n":
            cols[name] = rs.poisson(3, n).astype(numpy.int64)
        else:
            cols[name] = rs.normal(0, 1, n)
    return cols


def _reference_hists(case):
    """The real histbook objects (reference numpy path), or None if histbook cannot be imported."""
    try:
        from histbook_b200 import _ref
        if not _ref.available():
            return None
        import histbook_b200
        histbook_b200.uninstall()
        hb = _ref.load()
    except Exception:
        return None
    env = dict((n, getattr(hb, n)) for n in ("Hist", "Book", "bin", "intbin", "split", "cut", "profile", "groupby", "groupbin"))
    hists = [eval(src, dict(env)) for src in case["hists"]]
    if len(hists) == 1:
        return hists[0]
    return hb.Book(dict(("h%03d" % i, h) for i, h in enumerate(hists)))


def worker(args):
    """Fill `n` events `reps` times; returns (kind, seconds of fill only per rep)."""
    workload, n, seed, reps = args
    warnings.filterwarnings("ignore")
    case = load_case(WORKLOADS[workload][0])
    cols = make_columns(workload, n, seed)
    target = _reference_hists(case)
    times = []
    if target is not None:
        kind = "reference"
        for _ in range(reps):
            t0 = time.perf_counter()
            target.fill(**cols)
            times.append(time.perf_counter() - t0)
    else:
        kind = "port"
        from oracle import fill_oracle
        contents = None
        for _ in range(reps):
            t0 = time.perf_counter()
            contents = fill_oracle.fill(case["spec"], cols, contents)
            times.append(time.perf_counter() - t0)
    return kind, times


def run_parallel(workload, events_per_proc, procs, reps=1, seed=12345):
    """The reference's documented parallel model (README.rst:29): P processes each fill their shard, then add.
    Returns (kind, per-rep wall time = max over workers, total events per rep)."""
    import multiprocessing
    ctx = multiprocessing.get_context("spawn")
    with ctx.Pool(procs) as pool:
        results = pool.map(worker, [(workload, events_per_proc, seed + i, reps) for i in range(procs)])
    kind = results[0][0]
    per_rep = [max(r[1][k] for r in results) for k in range(reps)]
    return kind, per_rep, events_per_proc * procs
