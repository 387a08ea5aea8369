"""The fill driver: ``Hist.fill`` / ``Book.fill`` on the device.

Stands in for ``Fillable._fill`` (histbook/fill.py:85-162), ``Hist._prefill`` / ``Hist._postfill``
(histbook/hist.py:381-504) and ``Book.fill`` (histbook/book.py:540-586): one plan per Hist or Book
(the neutral spec, built once), one generated kernel per set of column dtypes (NVRTC, cached), one
launch per fill that updates every histogram of the book in a single pass over the columns.
"""
import threading

import numpy

from histbook_b200 import codegen, columns, engine, spec as hbspec
from histbook_b200.store import GroupInfo, Store

_options = {}
_lock = threading.RLock()
last_stats = {}


def set_options(**kwargs):
    """Kernel-generation options (``smem_budget``, ``threads``, ``min_blocks``, ``strategy``); clears cached kernels."""
    _options.update(kwargs)
    for k in [k for k, v in _options.items() if v is None]:
        del _options[k]
    Plan.generation += 1


class Plan(object):
    """What has to be computed for a fixed list of histograms (built once, like Fillable.fields)."""
    generation = 0

    def __init__(self, hists=None, spec=None):
        self.spec = hbspec.book_spec(hists) if spec is None else spec
        self.nhists = len(self.spec["hists"])
        self.fields = hbspec.spec_fields(self.spec)
        self.kernels = {}           # dtype signature -> (fill Kernel, Program, keys Kernel, keys Program)
        self.generation = Plan.generation
        self.keysets = None

    def kernel_for(self, cols):
        if self.generation != Plan.generation:
            self.kernels.clear()
            self.generation = Plan.generation
        sig = tuple((n, cols[n].dtype.str, cols[n].is_scalar) for n in self.fields)
        if sig not in self.kernels:
            coltypes = dict((n, (cols[n].dtype, cols[n].is_scalar)) for n in self.fields)
            kern = codegen.generate(self.spec, coltypes, _options)
            prog = engine.Program(engine.compile_source(kern.source), kern.name, kern.threads, kern.smem_bytes)
            keys = codegen.generate_keys(self.spec, coltypes, _options)
            kprog = None
            if keys is not None:
                kprog = engine.Program(engine.compile_source(keys.source), keys.name, keys.threads, 0)
            self.kernels[sig] = (kern, prog, keys, kprog)
        return self.kernels[sig]


PILOT_EVENTS = 1 << 20          # events filled through the global path before the hot window is placed
WINDOW_MIN_COVERAGE = 0.25      # below this (estimated) fraction of events a window is not worth its merge


def choose_window(marginals, wmax):
    """Place the hot window: per axis a contiguous bin range [lo, lo+n) such that prod(n) <= wmax and the
    (independence-estimated) fraction of events inside is as large as a greedy trim can make it.

    marginals: per-axis arrays of |sum of weights| per bin (from the histogram filled so far).
    Returns ([(lo, n), ...], estimated coverage)."""
    los, his = [], []
    for m in marginals:
        nz = numpy.flatnonzero(m)
        if len(nz) == 0:
            return [(0, 0)] * len(marginals), 0.0
        los.append(int(nz[0]))
        his.append(int(nz[-1]) + 1)
    totals = [float(m.sum()) for m in marginals]

    def volume():
        v = 1
        for lo, hi in zip(los, his):
            v *= hi - lo
        return v
    vol = volume()
    while vol > wmax:
        best = None
        for k, m in enumerate(marginals):
            nk = his[k] - los[k]
            if nk <= 1:
                continue
            gain = vol // nk                       # bins saved by dropping one slice of axis k
            for end, idx in (("lo", los[k]), ("hi", his[k] - 1)):
                score = (m[idx] / totals[k]) / gain
                if best is None or score < best[0]:
                    best = (score, k, end)
        if best is None:
            return [(0, 0)] * len(marginals), 0.0
        _, k, end = best
        if end == "lo":
            los[k] += 1
        else:
            his[k] -= 1
        vol = volume()
    coverage = 1.0
    for k, m in enumerate(marginals):
        coverage *= float(m[los[k]:his[k]].sum()) / totals[k]
    return [(lo, hi - lo) for lo, hi in zip(los, his)], coverage


def window_params(kern, window):
    """p.win of hb_template.cuh: [u32 words to zero, bins in the window, lo_0, n_0, lo_1, n_1, ...]."""
    info = kern.window
    if window is None:
        return [0, 0] + [0, 0] * len(info["totbins"])
    w = 1
    for lo, n in window:
        w *= n
    out = [(w * info["bpb"] + 3) // 4, w]
    for lo, n in window:
        out += [lo, n]
    return out


def _abi_columns(kern, cols):
    return [cols[name].abi() for name, dtype, is_scalar in kern.cols]


def _store_of(hist):
    st = hist.__dict__.get("_hb")
    if st is None:
        st = hist.__dict__["_hb"] = Store(hist._shape, [], None)
    return st


def fill_hists(plan, hists, arrays, timed=False):
    """One pass over ``arrays`` filling every histogram in ``hists`` (deduplicated, in plan order)."""
    with _lock:
        cols, length = columns.bind_all(plan.fields, arrays)
        kern, prog, keys, kprog = plan.kernel_for(cols)

        # group axes: which keys does this fill hold?  (numpy.unique in calc/__init__.py:150,178,182)
        discovered = {}
        if keys is not None:
            if plan.keysets is None or len(plan.keysets) != len(keys.group_goals):
                plan.keysets = [engine.KeySet(1 << 12) for _ in keys.group_goals]
            for ks in plan.keysets:
                ks.clear()
            if length > 0:
                engine.fill(kprog, _abi_columns(keys, cols), length, [], aux=[ks.ptr for ks in plan.keysets])
            for (goalkey, info), ks in zip(keys.group_goals, plan.keysets):
                patterns, overflowed = ks.read()
                if overflowed:
                    raise codegen.UnsupportedError("a group axis produced more than {0} distinct keys in one fill".format(ks.capacity // 2))
                discovered[goalkey] = (patterns, info)

        stores = []
        aux = [0] * len(kern.group_aux)
        for hid, (hist, hspec) in enumerate(zip(hists, plan.spec["hists"])):
            st = _store_of(hist)
            if hist.__dict__.get("_copyonfill", False):       # hist.py:351-353
                st = hist.__dict__["_hb"] = _clone_store(st)
                hist.__dict__["_copyonfill"] = False
            if hspec["group"]:
                infos = []
                fill_keys = []
                for a in hspec["group"]:
                    goalkey = hbspec.tree_key(a["goal"])
                    patterns, info = discovered[goalkey]
                    infos.append(GroupInfo(a["kind"], a["keeporder"], info["kind"] == "f", info["dtype"], info["nanflow"], goalkey))
                    fill_keys.append([int(x) for x in patterns])
                st.groups = infos
                st.to_device()
                st.admit(fill_keys)
                aux[kern.group_aux[hid]] = st.table.ptr
            else:
                st.to_device()
            stores.append(st)

        stats = {"events": length, "launches": 0}
        if length > 0:
            abi = _abi_columns(kern, cols)
            arenas = [st.arena for st in stores]
            if kern.window is None:
                stats = engine.fill(prog, abi, length, arenas, aux=aux, timed=timed)
            else:
                # hot-window privatisation: place the window from what the histogram already holds; a histogram
                # that is still empty first gets a pilot slice of this fill through the global path.
                wst = stores[kern.window["hid"]]
                done = 0
                extra = 0.0
                if wst.window is None:
                    if not wst.nonempty:
                        done = min(length, max(kern.tile, PILOT_EVENTS // kern.tile * kern.tile))
                        pilot = engine.fill(prog, abi, done, arenas, aux=aux, win=window_params(kern, None), timed=timed)
                        extra = pilot["ms_kernel"]
                    marg = wst.arena.marginals(kern.window["totbins"], kern.window["ncol"], kern.window["col"])
                    window, coverage = choose_window(marg, kern.window["wmax"])
                    wst.window = window if coverage >= WINDOW_MIN_COVERAGE else False
                    wst.window_coverage = coverage
                win = window_params(kern, wst.window or None)
                stats = {"events": length, "launches": 1 if done else 0, "ms_kernel": extra}
                if length > done:
                    main = engine.fill(prog, abi, length - done, arenas, aux=aux, win=win, timed=timed, offset=done)
                    main["launches"] += stats["launches"]
                    main["ms_kernel"] += extra
                    stats = main
                stats["window"] = list(wst.window) if wst.window else None
                stats["window_coverage"] = wst.window_coverage
        for st in stores:
            st.mark_filled()
        stats["regs"] = prog.regs
        stats["smem_bytes"] = kern.smem_bytes
        stats["keys_launches"] = 1 if (keys is not None and length > 0) else 0
        last_stats.clear()
        last_stats.update(stats)
        return length


def _clone_store(st):
    """Private copy of a shared store (copy-on-fill, histbook/hist.py:86-91,351-353)."""
    out = Store(st.shape, st.groups, None)
    if st.host_valid:
        host = st.host
        out.host = _copy_host(host)
        out.host_valid = True
        out.dev_valid = False
    else:
        out.arena = st.arena.clone()
        out.tuples = type(st.tuples)(st.tuples)
        out.skeleton = _copy_skeleton(st.skeleton)
        out.table = None
        out.host_valid = False
        out.dev_valid = True
    return out


def _copy_host(content):
    if content is None:
        return None
    if isinstance(content, numpy.ndarray):
        return content.copy()
    return type(content)((k, _copy_host(v)) for k, v in content.items())


def _copy_skeleton(skel):
    if skel is None or not isinstance(skel, dict):
        return skel
    return type(skel)((k, _copy_skeleton(v)) for k, v in skel.items())


class SpecHist(object):
    """Stand-in for a histbook ``Hist`` when only a neutral spec is at hand (golden fixtures, bench
    workloads): carries the shape and the device store, nothing else."""

    def __init__(self, hspec):
        self._shape = tuple(hspec["shape"])
        self._copyonfill = False

    @property
    def content(self):
        st = self.__dict__.get("_hb")
        return None if st is None else st.get()

    def clear(self):
        self.__dict__.pop("_hb", None)


class SpecFill(object):
    """Fill driver over a neutral spec (no histbook import needed): ``SpecFill(spec).fill(x=..., y=...)``."""

    def __init__(self, spec):
        self.plan = Plan(spec=spec)
        self.hists = [SpecHist(h) for h in spec["hists"]]

    @property
    def fields(self):
        return list(self.plan.fields)

    def fill(self, arrays=None, timed=False, **more):
        if arrays is None:
            arrays = more
        elif more:
            arrays = dict(arrays, **more)
        return fill_hists(self.plan, self.hists, arrays, timed=timed)

    def contents(self):
        return [h.content for h in self.hists]

    def arenas(self):
        return [h.__dict__["_hb"].arena for h in self.hists]

    def clear(self):
        for h in self.hists:
            h.clear()
