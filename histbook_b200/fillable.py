"""The fill driver: ``Hist.fill`` / ``Book.fill`` on the device.

Stands in for ``Fillable._fill`` (histbook/fill.py:85-162), ``Hist._prefill`` / ``Hist._postfill``
(histbook/hist.py:381-504) and ``Book.fill`` (histbook/book.py:540-586): one plan per Hist or Book
(the neutral spec, built once), one generated kernel per set of column dtypes (NVRTC, cached), one
launch per fill that updates every histogram of the book in a single pass over the columns.
"""
import json
import os
import threading

import numpy

from histbook_b200 import codegen, columns, engine, spec as hbspec
from histbook_b200.store import GroupInfo, Store

# kernel-generation options; HB_OPTIONS='{"fx": true}' presets them for a whole process (tests, tuning)
_options = dict(json.loads(os.environ.get("HB_OPTIONS", "") or "{}"))
_lock = threading.RLock()
last_stats = {}
TIMED = False                   # measurement: bracket every fill's launches with CUDA events (last_stats["ms_kernel"]); costs a device sync


def set_options(**kwargs):
    """Kernel-generation options; ``name=None`` restores a default.  Changing any of them drops the cached kernels
    (they are part of the kernel source, hence of its cache key).

    Semantics: ``deterministic`` (every float64 sum the correctly rounded exact sum), ``fx_stats``.
    Accumulation: ``fx`` (False | True | "mix" | [slots] | "auto"), ``fx_auto_max_terms``, ``strategy`` ({hist id | "*":
    "smem" | "global" | "window"}), ``smem_budget``, ``window`` / ``win_dense`` / ``win_pair`` / ``win_smem`` /
    ``win_threads`` / ``win_min_blocks`` / ``win_table32`` / ``win_max_rows`` / ``win_max_hist_bytes``, ``merge_pair``,
    ``pair_red``, ``group_slots``, ``replicas``, ``f64_atomic`` ("cas" | "exch"), ``sort`` ("auto" | True | False: the sorted
    strategy of big books, codegen.SortGroup) with ``sort_lanes`` / ``sort_split`` / ``sort_ept`` / ``sort_threads`` / ``sort_max_groups`` / ``sort_smem``, ``parts``, ``box``.
    Launch shape and loop: ``threads``, ``min_blocks``, ``unroll`` (4 | 1 | 0), ``prefetch``, ``dynamic`` (False | "tail" |
    True), ``evict_first``, ``red_evict_last``.  Measurement: ``block_times``.
    Defaults are the measured best (profiles/r01_sweeps.txt); tests/test_gpu_parity.py checks that none of them changes a
    histogram."""
    _options.update(kwargs)
    for k in [k for k, v in _options.items() if v is None]:
        del _options[k]
    Plan.generation += 1


class Plan(object):
    """What has to be computed for a fixed list of histograms (built once, like Fillable.fields)."""
    generation = 0

    def __init__(self, hists=None, spec=None, options=None):
        self.spec = hbspec.book_spec(hists) if spec is None else spec
        self.options = dict(options or {})   # overrides of the process-wide options (the parts of a split book)
        self.parts = {}             # dtype signature -> None | [(Plan of the part, indices of its histograms)]
        self.discovery = {}         # dtype signature -> (keys Kernel, keys Program) of the whole book
        self.nhists = len(self.spec["hists"])
        self.fields = hbspec.spec_fields(self.spec)
        self.kernels = {}           # dtype signature -> (fill Kernel, Program, keys Kernel, keys Program, range Kernel, range Program)
        self.generated = {}         # dtype signature -> (generated fill Kernel, its options, column types): kernel_for compiles on demand
        self.generation = Plan.generation
        self.keysets = None
        self.goal_tables = None     # tree key of a group goal -> (key patterns of the last fill, their device table)
        # fields that are the direct argument of a groupby axis may hold strings: dictionary-encoded on the host
        self.groupby_fields = {}    # field name -> tree key of the groupby goal
        for h in self.spec["hists"]:
            for a in h["group"]:
                g = a["goal"]
                if a["kind"] == "groupby" and g[0] == "call" and g[1] == "histbook.groupby" and g[2][0] == "name":
                    self.groupby_fields[g[2][1]] = hbspec.tree_key(g)
        self.goalkeys = [[hbspec.tree_key(a["goal"]) for a in h["group"]] for h in self.spec["hists"]]      # per histogram, per group axis
        self.string_codes = {}      # field name -> columns.StringCodes (created at the first string-valued fill)
        self.fx_scales = {}         # dtype signature -> fixed-point scale per term (from the pilot), see _fx_scales
        self.fx_sample = {}         # dtype signature -> events the pilot of fx_scales[signature] looked at
        self.fx_stat = None         # device counter: terms that missed their fixed-point window (all fills so far)
        self.fx_seen = 0            # ... its value when last looked at
        self.fx_last = None         # (signature, events, fills) since the miss counter was last looked at

    def generated_for(self, cols, dense=False, nobox=False, gfast=False):
        """(signature, generated -- not yet compiled -- fill kernel, its options) for these column dtypes and this variant."""
        self._fresh()
        sig = (tuple((n, cols[n].dtype.str, cols[n].is_scalar) for n in self.fields) + (("dense",) if dense else ()) + (("nobox",) if nobox else ())
               + (("gfast",) if gfast else ()))
        if sig not in self.generated:
            coltypes = dict((n, (cols[n].dtype, cols[n].is_scalar)) for n in self.fields)
            options = dict(_options, **self.options)
            if dense:
                options["win_dense"] = True
            if nobox:
                options["box"] = False
            options["group_fast"] = bool(gfast)
            self.generated[sig] = (codegen.generate(self.spec, coltypes, options), options, coltypes)
        return (sig,) + self.generated[sig][:2]

    def kernel_for(self, cols, dense=False, nobox=False, gfast=False):
        """The kernels for these column dtypes.  ``dense`` selects the second variant of a hot-window kernel (see
        fill_hists): part of the window's float64 terms in fixed point, plain REDs for the few events outside.  ``gfast``
        selects the variant of a book with group axes for fills in which the histograms of a signature share their slot
        maps and hold no more slots than are privatised in shared memory (fill_hists decides per fill)."""
        sig, kern, options = self.generated_for(cols, dense=dense, nobox=nobox, gfast=gfast)
        if sig not in self.kernels:
            coltypes = self.generated[sig][2]
            prog = engine.Program(engine.compile_source(kern.source), kern.name, kern.threads, kern.smem_bytes, tile=kern.tile,
                                  max_blocks_per_sm=int(options.get("max_blocks_per_sm", 0)))
            keys, kprog = self.discovery_for(cols)
            rng = rprog = None
            if kern.fx_terms:
                rng = codegen.generate_range(self.spec, coltypes, options)
                rprog = engine.Program(engine.compile_source(rng.source), rng.name, rng.threads, rng.smem_bytes)
            self.kernels[sig] = (kern, prog, keys, kprog, rng, rprog)
        return sig, self.kernels[sig]

    def _fresh(self):
        if self.generation != Plan.generation:
            self.kernels.clear()
            self.generated.clear()
            self.parts.clear()
            self.discovery.clear()
            self.generation = Plan.generation

    def parts_for(self, cols):
        """None, or the parts this book is filled in: [(Plan, histogram indices)].  A book is split
        * when it has more histograms than one kernel's parameter block takes (each part at most PART_MAX_HISTS), or
        * when asked to (option ``parts`` = 2, 4): a big book with at least ``parts_min_groups`` sort groups as half-books,
          each filled by 512-thread blocks, two per SM with their own tiles, so that one block's arithmetic (evaluate,
          index, tickets) overlaps the other's shared-memory traffic (the sums over the sorted events) -- the one
          1024-thread block runs its phases one after the other and waits at every barrier (kernels only, 1e8 events:
          9.4 ms whole, 2 x 4.15 ms in halves; tools/parts_probe.py).  Not the default: the parts re-read the columns.
        All parts are filled in ONE pass over the columns (engine.fill_multi): every chunk is staged once."""
        self._fresh()
        sig = tuple((n, cols[n].dtype.str, cols[n].is_scalar) for n in self.fields)
        if sig not in self.parts:
            options = dict(_options, **self.options)
            want = options.get("parts", "auto")
            k, extra = 1, {}
            if self.options.get("_part") or want in (1, False) or options.get("deterministic", False):
                k = 1
            elif self.nhists > PART_MAX_HISTS:
                k = -(-self.nhists // PART_MAX_HISTS)
            if k == 1 and not self.options.get("_part") and want not in (1, False, "auto") and not options.get("deterministic", False):
                # asked for (parts=2): a big book as half-books of 512-thread blocks, two per SM.  Measured on c4: 11.4 against
                # 10.9 Gev/s for the one 1024-thread kernel -- but every part reads the columns again (2.0x the DRAM traffic
                # of the single pass), so the default stays ONE fused pass over the columns.
                coltypes = dict((n, (cols[n].dtype, cols[n].is_scalar)) for n in self.fields)
                try:
                    kern = codegen.generate(self.spec, coltypes, options)
                except codegen.UnsupportedError:
                    kern = None
                if kern is not None and kern.sorted and len(kern.sorted["groups"]) >= int(options.get("parts_min_groups", 4)):
                    k = int(want)
                    extra = {"sort_threads": int(options.get("parts_threads", 512)), "sort_min_blocks": int(options.get("parts_min_blocks", 2)),
                             "sort_ept": int(options.get("parts_ept", 2)), "sort": True}
            if k <= 1:
                self.parts[sig] = None
            else:
                n = self.nhists
                bounds = [n * i // k for i in range(k + 1)]
                parts = []
                for a, b in zip(bounds, bounds[1:]):
                    sub = dict(self.spec, hists=self.spec["hists"][a:b])
                    parts.append((Plan(spec=sub, options=dict(self.options, _part=True, **extra)), list(range(a, b))))
                self.parts[sig] = parts
        return self.parts[sig]

    def discovery_for(self, cols):
        """The key-discovery kernel of the whole book (one pass, whatever the number of parts)."""
        self._fresh()
        sig = tuple((n, cols[n].dtype.str, cols[n].is_scalar) for n in self.fields)
        if sig not in self.discovery:
            coltypes = dict((n, (cols[n].dtype, cols[n].is_scalar)) for n in self.fields)
            keys = codegen.generate_keys(self.spec, coltypes, dict(_options, **self.options))
            kprog = None
            if keys is not None:
                kprog = engine.Program(engine.compile_source(keys.source), keys.name, keys.threads, keys.smem_bytes)
            self.discovery[sig] = (keys, kprog)
        return self.discovery[sig]


PART_MAX_HISTS = 96             # histograms per kernel of a split book (HbParams holds 128 arena pointers)
MAX_GROUP_KEYS = 1 << 22        # distinct keys of one group axis in one fill (the discovery table grows up to twice this)
PILOT_EVENTS = 1 << 20          # events filled through the global path before the hot window is placed
FX_MISS_REPILOT = 0.01          # re-run the fixed-point pilot when more than this fraction of the events missed
FX_CHECK_FILLS = 64             # ... checked after this many fills
FX_CHECK_EVENTS = 4 * 10**9     # ... or this many events, whichever comes first


def _fx_scales(plan, sig, kern, rng, rprog, cols, length):
    """Scale (binary exponent of the least significant bit) of every fixed-point term of ``kern``.

    A pilot kernel takes the largest binary exponent each term reaches over the first PILOT_EVENTS events; the
    128-bit field is then placed so that values up to 4x that maximum fit (hb_template.cuh: mantissa shift
    0..50, i.e. everything within 49 binades below the pilot's maximum is added exactly, the rest goes to the
    global histogram as float64 REDs -- correct either way, the pilot only decides which path is the fast one).
    Cached per plan and dtype signature; dropped again when a fill reports too many misses."""
    if plan.fx_stat is None:
        plan.fx_stat = engine.DeviceBuffer(numpy.zeros(2, dtype=numpy.uint64))
        plan.fx_seen = 0
        plan.fx_last = None
    elif plan.fx_last is not None:
        # look at the miss counter now and then only: reading it is a device->host synchronisation
        last_sig, events, fills = plan.fx_last
        if fills >= FX_CHECK_FILLS or events >= FX_CHECK_EVENTS:
            total = int(plan.fx_stat.read(numpy.uint64)[0])
            missed, plan.fx_seen = total - plan.fx_seen, total
            if events > 0 and missed > FX_MISS_REPILOT * events:
                plan.fx_scales.pop(last_sig, None)
            plan.fx_last = None
    # (a pilot that saw only a sliver of data -- a first fill of a few thousand events -- is repeated once a fill
    #  brings four times as much, until it has seen PILOT_EVENTS)
    sample = plan.fx_sample.get(sig, 0)
    if sig in plan.fx_scales and sample < PILOT_EVENTS and min(length, PILOT_EVENTS) >= 4 * max(sample, 1):
        plan.fx_scales.pop(sig)
    if sig not in plan.fx_scales:
        maxima = engine.DeviceBuffer(numpy.zeros(max(rng.nfx, 2), dtype=numpy.uint32))
        m = min(length, PILOT_EVENTS)
        plan.fx_sample[sig] = m
        if m > 0:
            engine.fill(rprog, _abi_columns(rng, cols), m, [], aux=[maxima.ptr])
        biased = maxima.read(numpy.uint32)[:rng.nfx].astype(numpy.int64)
        biased[biased == 0] = 1023                       # nothing seen (or only zeros): assume values near 1
        # what the kernel gets is k = 1075 + s: a value with biased exponent e is shifted by e - k = e - max + 48
        scales = [int(b) - 48 for b in biased]
        while len(scales) < kern.nfx:
            scales.append(1023 - 48)
        plan.fx_scales[sig] = scales
    if plan.fx_last is not None and plan.fx_last[0] == sig:
        plan.fx_last = (sig, plan.fx_last[1] + length, plan.fx_last[2] + 1)
    else:
        plan.fx_last = (sig, length, 1)
    return plan.fx_scales[sig]

WINDOW_MIN_COVERAGE = 0.25      # below this (estimated) fraction of events a window is not worth its merge
WINDOW_DENSE_COVERAGE = 0.95    # the dense variant (fixed point inside, plain REDs outside) needs its smaller window to hold this much
WINDOW_CHECK_FILLS = 64         # a placed window is looked at again after this many fills ...
WINDOW_CHECK_EVENTS = 8 * 10**9   # ... or this many events, whichever comes first (one read of the histogram: a host synchronisation)
WINDOW_DRIFT = 0.05             # ... and placed anew when it holds this much less of the recent fills than a new one would


def _ragged_density(wst, info):
    """|sum of weights| the histogram holds per (row of the leading axes, bin of the trailing axes): what a window is placed from"""
    content = wst.arena.read(0, info["rows"] * info["trail"] * info["ncol"])
    density = numpy.abs(content.reshape(info["rows"], info["trail"], info["ncol"])[:, :, info["col"]])
    density[~numpy.isfinite(density)] = 0.0
    return density


def _place_ragged(plan, cols, kern, info, wst, density):
    """Place the ragged hot window of histogram store ``wst`` from ``density`` (choose_ragged): the dense kernel variant if its
    smaller window still holds WINDOW_DENSE_COVERAGE of it, else the sparse one if that holds WINDOW_MIN_COVERAGE, else none."""
    wst.window = False
    if _options.get("fx", "auto") == "auto" and not kern.deterministic and len(kern.plans[info["hid"]].f64_terms) >= 2:
        dinfo = plan.kernel_for(cols, dense=True)[1][0].window
        if dinfo is not None:
            first, nlen, coverage = choose_ragged(density, dinfo["wmax"])
            if coverage >= WINDOW_DENSE_COVERAGE:
                table, bins = window_table(first, nlen, dinfo["table_bits"])
                wst.window = {"table": engine.DeviceBuffer(table), "bins": bins, "rows": int((nlen > 0).sum()), "dense": True,
                              "first": first, "length": nlen}
                wst.window_coverage = coverage
    if not wst.window:
        first, nlen, coverage = choose_ragged(density, info["wmax"])
        wst.window_coverage = coverage
        if coverage >= WINDOW_MIN_COVERAGE:
            table, bins = window_table(first, nlen, info["table_bits"])
            wst.window = {"table": engine.DeviceBuffer(table), "bins": bins, "rows": int((nlen > 0).sum()), "dense": False,
                          "first": first, "length": nlen}


def _watch_window(plan, cols, kern, info, wst, length):
    """A window is placed from what the histogram held at its first fills; data that drifts away from there would fall to the
    global-RED path for good.  Every WINDOW_CHECK_FILLS fills (or WINDOW_CHECK_EVENTS events) the histogram is read once:
    the share of the fills SINCE the last look that the window caught becomes ``window_coverage`` (what last_stats reports), and
    if a window placed from those recent fills alone would hold WINDOW_DRIFT more, it replaces the old one."""
    wst.window_fills += 1
    wst.window_events += length
    if wst.window_fills < WINDOW_CHECK_FILLS and wst.window_events < WINDOW_CHECK_EVENTS:
        return
    density = _ragged_density(wst, info)
    seen, wst.window_seen = wst.window_seen, density
    wst.window_fills = wst.window_events = 0
    if seen is None or seen.shape != density.shape:
        return
    recent = numpy.abs(density - seen)
    total = float(recent.sum())
    if total <= 0.0 or not numpy.isfinite(total):
        return
    current = 0.0
    if wst.window:
        trail = numpy.arange(density.shape[1])[None, :]
        first, nlen = wst.window["first"], wst.window["length"]
        current = float(recent[(trail >= first[:, None]) & (trail < (first + nlen)[:, None])].sum()) / total
    wmax = info["wmax"]
    if wst.window and wst.window["dense"]:
        wmax = plan.kernel_for(cols, dense=True)[1][0].window["wmax"]
    best = choose_ragged(recent, wmax)[2]
    if best >= current + WINDOW_DRIFT and best >= WINDOW_MIN_COVERAGE:
        _place_ragged(plan, cols, kern, info, wst, recent)
    else:
        wst.window_coverage = current if wst.window else best


def choose_ragged(density, wmax):
    """Place the hot window: one contiguous interval [first, first + length) per row of ``density`` (rows =
    flattened leading axes, columns = flattened trailing axes; values = |sum of weights| the histogram holds so
    far) such that the intervals together have at most ``wmax`` bins and hold as much of the density as a
    threshold rule can give: every row keeps the span between its first and last bin at or above a common
    density threshold, the threshold being the lowest one that fits (binary search).

    Returns (first[rows], length[rows], coverage)."""
    density = numpy.asarray(density, dtype=numpy.float64)
    rows, trail = density.shape
    first = numpy.zeros(rows, dtype=numpy.int64)
    length = numpy.zeros(rows, dtype=numpy.int64)
    total = float(density.sum())
    if total <= 0.0 or not numpy.isfinite(total) or wmax < 1:
        return first, length, 0.0
    levels = numpy.unique(density[density > 0.0])

    def spans(t):
        mask = density >= t
        has = mask.any(axis=1)
        lo = numpy.where(has, mask.argmax(axis=1), 0)
        hi = numpy.where(has, trail - mask[:, ::-1].argmax(axis=1), 0)
        return lo, hi
    a, b = 0, len(levels) - 1               # lowest threshold index whose spans fit
    lo, hi = spans(levels[b])
    if int((hi - lo).sum()) > wmax:
        # even the single densest bin per row does not fit: keep the densest bins of the densest rows only
        order = numpy.argsort(-density.max(axis=1))[:wmax]
        first[order] = density[order].argmax(axis=1)
        length[order] = 1
    else:
        while a < b:
            mid = (a + b) // 2
            lo, hi = spans(levels[mid])
            if int((hi - lo).sum()) <= wmax:
                b = mid
            else:
                a = mid + 1
        lo, hi = spans(levels[b])
        first, length = lo.astype(numpy.int64), (hi - lo).astype(numpy.int64)
    cols = numpy.arange(trail)[None, :]
    inside = (cols >= first[:, None]) & (cols < (first + length)[:, None])
    return first, length, float(density[inside].sum()) / total


def choose_box(marginals, wmax):
    """Place a box window: one interval [lo_k, lo_k + n_k) per fixed axis with at most ``wmax`` bins in the box, around
    the bulk of the histogram.  ``marginals[k]`` = what the histogram holds, summed over all other axes.  Every axis
    takes the shortest interval that holds a common fraction f of its marginal; f is the largest that fits (bisection).
    Returns (lo[k], n[k], product of the marginal fractions = coverage if the axes were independent)."""
    margs = []
    for m in marginals:
        m = numpy.abs(numpy.asarray(m, dtype=numpy.float64))
        m[~numpy.isfinite(m)] = 0.0
        margs.append(m)
    if wmax < 1 or any(m.sum() <= 0.0 for m in margs):
        return [0] * len(margs), [0] * len(margs), 0.0

    def shortest(m, f):
        """shortest [a, b) with sum(m[a:b]) >= f * total (two pointers)"""
        need = f * m.sum()
        c = numpy.concatenate([[0.0], numpy.cumsum(m)])
        best = (0, len(m))
        a = 0
        for b in range(1, len(m) + 1):
            while a + 1 < b and c[b] - c[a + 1] >= need:
                a += 1
            if c[b] - c[a] >= need and b - a < best[1] - best[0]:
                best = (a, b)
        return best

    def boxes(f):
        return [shortest(m, f) for m in margs]

    def volume(bx):
        v = 1
        for a, b in bx:
            v *= (b - a)
        return v
    lo_f, hi_f = 0.0, 1.0
    best = boxes(0.0)
    if volume(best) > wmax:
        return [0] * len(margs), [0] * len(margs), 0.0
    for _ in range(24):
        mid = 0.5 * (lo_f + hi_f)
        bx = boxes(mid)
        if volume(bx) <= wmax:
            lo_f, best = mid, bx
        else:
            hi_f = mid
    # use what is left of the capacity: widen one axis at a time towards the heavier neighbouring bin
    grown = True
    while grown:
        grown = False
        order = sorted(range(len(margs)), key=lambda k: best[k][1] - best[k][0])
        for k in order:
            a, b = best[k]
            left = margs[k][a - 1] if a > 0 else -1.0
            right = margs[k][b] if b < len(margs[k]) else -1.0
            if left < 0 and right < 0:
                continue
            cand = (a - 1, b) if left >= right else (a, b + 1)
            trial = list(best)
            trial[k] = cand
            if volume(trial) <= wmax:
                best, grown = trial, True
                break
    coverage = 1.0
    for m, (a, b) in zip(margs, best):
        coverage *= float(m[a:b].sum() / m.sum())
    return [a for a, b in best], [b - a for a, b in best], coverage


def window_table(first, length, bits=64):
    """The kernel's row table (codegen `we`): per row one uint64 = offset of the row's interval in the window (low 32
    bits) | first in-row index << 48 | length << 32 -- or, when the kernel was generated for packed entries
    (kern.window["table_bits"] == 32), one uint32 = first << 24 | length << 16 | offset.
    Returns (table, bins in the window)."""
    first = numpy.asarray(first, dtype=numpy.uint64)
    length = numpy.asarray(length, dtype=numpy.uint64)
    offset = numpy.concatenate([[0], numpy.cumsum(length)[:-1]]).astype(numpy.uint64)
    if bits == 32:
        assert (first < 256).all() and (length < 256).all() and (offset < 65536).all()
        table = ((first << numpy.uint64(24)) | (length << numpy.uint64(16)) | offset).astype(numpy.uint32)
    else:
        table = offset | (length << numpy.uint64(32)) | (first << numpy.uint64(48))
    return table, int(length.sum())


def window_params(kern, bins):
    """p.win of hb_template.cuh: [u32 words of window to zero, bins in the window].  The kernel lays the window out
    for its full capacity (slot blocks `wmax` rows apart), so the words to zero span the capacity whenever any bin is in use."""
    return [(kern.window["wmax"] * kern.window["bpb"] + 3) // 4 if bins else 0, bins]


def _abi_columns(kern, cols):
    return [cols[name].abi() for name, dtype, is_scalar in kern.cols]


def _store_of(hist):
    st = hist.__dict__.get("_hb")
    if st is None:
        st = hist.__dict__["_hb"] = Store(hist._shape, [], None)
    return st


def _discover(plan, keys, kprog, cols, length):
    """Keys every group goal takes in this fill: {tree key of the goal: (64-bit patterns, info)} (numpy.unique in
    calc/__init__.py:150,178,182).  One small kernel over the group columns + one device hash set per goal; a set that
    fills beyond one half is replaced by one four times the size and the pass repeated."""
    discovered = {}
    if keys is None:
        return discovered
    if plan.keysets is None or len(plan.keysets) != len(keys.group_goals):
        plan.keysets = [engine.KeySet(1 << 12) for _ in keys.group_goals]
    while True:
        for ks in plan.keysets:
            ks.clear()
        if length > 0:
            engine.fill(kprog, _abi_columns(keys, cols), length, [], aux=[ks.ptr for ks in plan.keysets])
        grown = False
        for i, ((goalkey, info), ks) in enumerate(zip(keys.group_goals, plan.keysets)):
            patterns, overflowed = ks.read()
            if overflowed or 2 * len(patterns) > ks.capacity:
                # numpy.unique has no limit: a fuller table means longer probes, a full one lost keys -- take a
                # table four times the size and run the discovery again (load factor at most 1/2 when accepted)
                if ks.capacity >= MAX_GROUP_KEYS * 2:
                    raise codegen.UnsupportedError("a group axis produced more than {0} distinct keys in one fill".format(MAX_GROUP_KEYS))
                plan.keysets[i] = engine.KeySet(ks.capacity * 4)
                grown = True
            discovered[goalkey] = (patterns, info)
        if not grown:
            return discovered


def perfect_multiplier(patterns, bits):
    """An odd 64-bit multiplier M for which (key * M mod 2**64) >> (64 - bits) differs for all of ``patterns``, or 0 if none
    of the candidates does (with n keys over 2**bits >= n*n/2 entries about one candidate in three works)."""
    m = 0x9E3779B97F4A7C15
    for _ in range(4096):
        if len(set(((p * m) & 0xFFFFFFFFFFFFFFFF) >> (64 - bits) for p in patterns)) == len(patterns):
            return m
        m = (m * 0xD1342543DE82EF95 + 0x2545F4914F6CDD1D) & 0xFFFFFFFFFFFFFFFF | 1
    return 0


def goal_table(patterns):
    """Device table of one group goal (hb_template.cuh hb_key_find): [n, M, shift, npad, the fill's keys as ascending
    unsigned 64-bit patterns (npad of them, even, at least 8), and -- for up to 64 keys -- 2**bits (key, position) entries
    addressed by (key * M) >> shift; M = 0: no hash, the kernel bisects the keys]."""
    patterns = tuple(patterns)
    n = len(patterns)
    npad = max(8, n + (n & 1))
    bits, mult = 0, 0
    if 0 < n <= 64:
        bits = 4
        while (1 << bits) < max(16, n * n // 2):
            bits += 1
        mult = perfect_multiplier(patterns, bits)
    words = [n, mult, 64 - bits, npad] + list(patterns) + [0] * (npad - n)
    if mult:
        entries = [(0, 0xFFFFFFFFFFFFFFFF)] * (1 << bits)
        for pos, p in enumerate(patterns):
            entries[((p * mult) & 0xFFFFFFFFFFFFFFFF) >> (64 - bits)] = (p, pos)
        for key, pos in entries:
            words.extend((key, pos))
    return words


def _bind_stores(plan, kern, hists, discovered):
    """The stores of ``hists`` made current on the device, group keys of this fill admitted; returns (stores, aux
    pointers of the kernel's group tables: per goal the fill's keys, per group histogram its dense slot map)."""
    aux = [0] * (len(kern.group_aux) + len(kern.goal_aux))
    if kern.goal_aux:
        # per group goal: [number of keys, 0, this fill's keys as ascending unsigned 64-bit patterns] (codegen._goal_lookup)
        if plan.goal_tables is None:
            plan.goal_tables = {}
        for goalkey, slot in kern.goal_aux.items():
            patterns = tuple(sorted(int(x) for x in discovered[goalkey][0]))
            cached = plan.goal_tables.get(goalkey)
            if cached is None or cached[0] != patterns:
                cached = plan.goal_tables[goalkey] = (patterns, engine.DeviceBuffer(numpy.array(goal_table(patterns), dtype=numpy.uint64)))
            aux[slot] = cached[1].ptr
    stores = []
    for hid, (hist, hspec) in enumerate(zip(hists, plan.spec["hists"])):
        st = _store_of(hist)
        if hist.__dict__.get("_copyonfill", False):       # hist.py:351-353
            st = hist.__dict__["_hb"] = _clone_store(st)
            hist.__dict__["_copyonfill"] = False
        if hspec["group"]:
            infos = []
            fill_keys = []
            for a, goalkey in zip(hspec["group"], plan.goalkeys[hid]):
                patterns, info = discovered[goalkey]
                codes = None
                for field, gk in plan.groupby_fields.items():
                    if gk == goalkey and field in plan.string_codes:
                        codes = plan.string_codes[field]
                infos.append(GroupInfo(a["kind"], a["keeporder"], info["kind"] == "f", info["dtype"], info["nanflow"], goalkey, codes))
                fill_keys.append([int(x) for x in patterns])
            st.groups = infos
            st.to_device()
            st.admit(fill_keys)
            # histograms that group by the same goals usually hold the same keys in the same slots (they were filled together):
            # a null pointer tells the kernel to use the slot its signature's first histogram looked up (codegen._group_lookup)
            leader = kern.group_leader.get(hid, hid)
            if leader != hid and stores[leader].table is not None and stores[leader].table[2] == st.table[2]:
                aux[kern.group_aux[hid]] = 0
            else:
                aux[kern.group_aux[hid]] = st.table[1].ptr
        else:
            st.to_device()
        stores.append(st)
    return stores, aux


def _group_fast(plan, kern, stores, aux, cols):
    """Whether this fill may run the ``gfast`` variant of a book's kernel (codegen option group_fast): every histogram that
    shares its group goals with an earlier one also shares that one's slot map (``_bind_stores`` passed a null map), and
    every histogram with group axes that is privatised in shared memory holds no more slab slots than are privatised there
    -- then the kernel needs neither the per-histogram map loads nor the global-memory path of the slots beyond. The
    usual case for the histograms of a book, which are always filled together (c4: 11 keys, 16 privatised slots)."""
    if not kern.group_leader or not _options.get("group_fast", True) or kern.window is not None or kern.deterministic:
        return False
    if any(aux[kern.group_aux[hid]] != 0 for hid in kern.group_leader):
        return False
    vkern = plan.generated_for(cols, gfast=True)[1]
    if vkern.group_aux != kern.group_aux or vkern.goal_aux != kern.goal_aux or vkern.naux != kern.naux:
        return False
    for hid in vkern.group_aux:
        hp = vkern.plans[hid]
        if hp.strategy == "smem" and len(stores[hid].tuples) > hp.gcap:
            return False
    return True


def _fill_parts(plan, parts, hists, cols, length, timed):
    """A split book (Plan.parts_for): one discovery pass, one pass over the columns, one kernel per part and chunk."""
    on_device = [c for c in cols.values() if c.location == 1]
    if on_device:
        dev = engine.init()
        for c in on_device:
            if c.device is not None and c.device != dev:
                raise codegen.UnsupportedError("field {0!r} lives on cuda:{1} but this process fills on cuda:{2} (one process per GPU)".format(c.name, c.device, dev))
            if c.stream is not None:
                engine.stream_wait(0, c.stream)
    side = engine.producer_stream() if length > 0 and on_device else 0
    if side:
        engine.stream_wait(0, side)
    keys, kprog = plan.discovery_for(cols)
    discovered = _discover(plan, keys, kprog, cols, length)
    order = list(plan.fields)
    launches, finals, all_stores, kerns = [], [], [], []
    for sub, idx in parts:
        sub.string_codes, sub.groupby_fields = plan.string_codes, plan.groupby_fields
        sig, (kern, prog, _k, _kp, rng, rprog) = sub.kernel_for(cols)
        if kern.window is not None:
            raise codegen.UnsupportedError("a part of a split book wants a hot window")       # (windows are for single large histograms)
        stores, aux = _bind_stores(sub, kern, [hists[i] for i in idx], discovered)
        aux = aux + [0] * (kern.naux - len(aux))
        fx = []
        if kern.fx_terms:
            fx = _fx_scales(sub, sig, kern, rng, rprog, cols, length)
            aux[kern.fx_aux] = sub.fx_stat.ptr
        launches.append({"program": prog, "colmap": [order.index(name) for name, dtype, is_scalar in kern.cols],
                         "arenas": [st.arena for st in stores], "aux": aux, "win": [0, 0] + fx, "exact": None})
        all_stores.extend(stores)
        kerns.append((kern, prog))
    stats = {"events": length, "launches": 0}
    if length > 0:
        stats = engine.fill_multi(launches, [cols[name].abi() for name in order], length, timed=timed)
    if side:
        engine.stream_wait(side, 0)
    for st in all_stores:
        st.mark_filled()
    kern, prog = kerns[0]
    stats.update({"deterministic": False, "sorted": kern.sorted, "regs": prog.regs, "smem_bytes": kern.smem_bytes, "parts": len(parts),
                  "keys_launches": 1 if (keys is not None and length > 0) else 0})
    last_stats.clear()
    last_stats.update(stats)
    return length


def fill_hists(plan, hists, arrays, timed=False):
    """One pass over ``arrays`` filling every histogram in ``hists`` (deduplicated, in plan order)."""
    with _lock:
        timed = timed or TIMED
        arrays = _encode_strings(plan, arrays)
        cols, length = columns.bind_all(plan.fields, arrays)
        parts = plan.parts_for(cols)
        if parts is not None:
            return _fill_parts(plan, parts, hists, cols, length, timed)
        # (generated, not compiled yet: which variant of the kernel runs is decided below, once the group keys are known)
        sig, kern, _kopts = plan.generated_for(cols)
        keys, kprog = plan.discovery_for(cols)
        # device columns produced on another stream (torch's current one): CUDA does not order a non-blocking stream
        # with the engine's, so the fill waits for the producer here and the producer for the fill at the end
        on_device = [c for c in cols.values() if c.location == 1]
        if on_device:
            dev = engine.init()
            for c in on_device:
                if c.device is not None and c.device != dev:
                    raise codegen.UnsupportedError("field {0!r} lives on cuda:{1} but this process fills on cuda:{2} (one process per GPU)".format(c.name, c.device, dev))
                if c.stream is not None:                      # __cuda_array_interface__ v3: wait for the exporter's stream
                    engine.stream_wait(0, c.stream)
        side = engine.producer_stream() if length > 0 and on_device else 0
        if side:
            engine.stream_wait(0, side)

        # group axes: which keys does this fill hold?  (numpy.unique in calc/__init__.py:150,178,182)
        discovered = _discover(plan, keys, kprog, cols, length)

        stores, aux = _bind_stores(plan, kern, hists, discovered)

        gfast = _group_fast(plan, kern, stores, aux, cols)
        sig, (kern, prog, _keys, _kprog, rng, rprog) = plan.kernel_for(cols, gfast=gfast)

        stats = {"events": length, "launches": 0}
        used = (sig, kern, prog, rng, rprog)

        def prepare(variant):
            """Everything one kernel variant needs for its launches: ABI columns, aux pointers, fixed-point scales,
            exact accumulators (deterministic mode) and the function that folds them into the contents."""
            vsig, vkern, vprog, vrng, vrprog = variant
            vaux = list(aux) + [0] * (vkern.naux - len(aux))
            fx = []
            if vkern.fx_terms:
                fx = _fx_scales(plan, vsig, vkern, vrng, vrprog, cols, length)
                vaux[vkern.fx_aux] = plan.fx_stat.ptr
            if vkern.times_aux is not None:
                plan.block_times = engine.DeviceBuffer(numpy.zeros(3 * 4096, dtype=numpy.uint64))
                vaux[vkern.times_aux] = plan.block_times.ptr
            exact = None
            if vkern.deterministic and vkern.exact:
                # deterministic mode: two zeroed 192-bit accumulators (upper / lower exponent window) per (row, float64 slot)
                exact = []
                for hid, st in enumerate(stores):
                    slots = vkern.exact.get(hid)
                    if not slots:
                        exact.append(None)
                        continue
                    need = (st.arena.size // st.shape[-1]) * len(slots) * 6
                    if st.exact is None or st.exact.size != need:
                        st.exact = engine.Arena(need)
                    exact.append(st.exact)

            def finalize():
                if exact is not None:
                    for hid, st in enumerate(stores):
                        slots = vkern.exact.get(hid)
                        if slots:
                            engine.fx_finalize(st.arena, st.exact, st.arena.size // st.shape[-1], st.shape[-1],
                                               [fx[i] for i, mask in slots], [mask for i, mask in slots])
            return {"kern": vkern, "prog": vprog, "abi": _abi_columns(vkern, cols), "aux": vaux, "fx": fx, "exact": exact, "finalize": finalize}

        def launch(v, n, win, offset=0):
            out = engine.fill(v["prog"], v["abi"], n, arenas, aux=v["aux"], win=win + v["fx"] + v.get("box", []), timed=timed, offset=offset, exact=v["exact"])
            v["finalize"]()
            return out

        if length > 0:
            arenas = [st.arena for st in stores]
            if kern.window is None:
                stats = launch(prepare(used), length, [0, 0])
            else:
                # Hot-window privatisation: place the window from what the histogram already holds; a histogram that
                # is still empty first gets a pilot slice of this fill through the global path.  Two kernel variants:
                # "sparse" (float64 CAS inside, lane-paired REDs outside) and -- default options only -- "dense"
                # (all but the first float64 term in fixed point, so 1.5x the bytes per bin; plain REDs outside),
                # taken when its smaller window still holds WINDOW_DENSE_COVERAGE of the events
                # (c2: 156 -> 171 Gev/s at 97.8 %; a 303x303 profile at ~60 %: 132 -> 92, so it stays sparse).
                info = kern.window
                wst = stores[info["hid"]]
                done = 0
                extra = 0.0
                key = (info["rows"], info["trail"], info["wmax"])
                if info.get("box") is not None and (wst.window is None or wst.window_key != key):
                    # box window (c5-class histograms): placed from the marginal distributions, which are summed on the device
                    if not wst.nonempty:
                        base = prepare(used)
                        done = min(length, max(kern.tile, PILOT_EVENTS // kern.tile * kern.tile))
                        extra = launch(base, done, [0, 0])["ms_kernel"]
                    dims = info["box"]
                    margs = []
                    for a in range(len(dims)):
                        proj = wst.arena.project(dims, info["ncol"], [k == a for k in range(len(dims))])
                        margs.append(proj.read().reshape(dims[a], info["ncol"])[:, info["col"]])
                    lo, nn, coverage = choose_box(margs, info["wmax"])
                    wst.window_key = key
                    wst.window_coverage = coverage
                    bins = int(numpy.prod(nn, dtype=numpy.int64)) if coverage >= WINDOW_MIN_COVERAGE else 0
                    wst.window = {"box": [int(x) for pair in zip(lo, nn) for x in pair], "bins": bins, "rows": len(dims), "dense": False} if bins else False
                elif wst.window is None or wst.window_key != key:
                    if not wst.nonempty:
                        base = prepare(used)
                        base["aux"][info["aux"]] = 0          # null row table = no window
                        done = min(length, max(kern.tile, PILOT_EVENTS // kern.tile * kern.tile))
                        extra = launch(base, done, [0, 0])["ms_kernel"]
                    density = _ragged_density(wst, info)
                    wst.window_key = key
                    _place_ragged(plan, cols, kern, info, wst, density)
                    wst.window_seen, wst.window_fills, wst.window_events = density, 0, 0
                elif info.get("box") is None:
                    _watch_window(plan, cols, kern, info, wst, length)
                if wst.window and wst.window["dense"]:
                    dsig, (dkern, dprog, _k, _kp, drng, drprog) = plan.kernel_for(cols, dense=True)
                    used = (dsig, dkern, dprog, drng, drprog)
                elif info.get("box") is not None and wst.window is False:
                    # no box worth its merge (c5: the largest one holds 9 % of the events): the plain kernel, which is not
                    # tied to one 1024-thread block per SM by a shared-memory window it would not use (c5: 108 -> 113 Gev/s)
                    dsig, (dkern, dprog, _k, _kp, drng, drprog) = plan.kernel_for(cols, nobox=True)
                    used = (dsig, dkern, dprog, drng, drprog)
                v = prepare(used)
                vinfo = used[1].window
                box_tail = []
                if vinfo is None:
                    win = [0, 0]
                elif vinfo.get("box") is not None:
                    win = window_params(used[1], wst.window["bins"]) if wst.window else [0, 0]
                    box_tail = list(wst.window["box"]) if wst.window else [0] * (2 * len(vinfo["box"]))
                elif wst.window:
                    v["aux"][vinfo["aux"]] = wst.window["table"].ptr
                    win = window_params(used[1], wst.window["bins"])
                else:
                    v["aux"][vinfo["aux"]] = 0
                    win = [0, 0]
                v["box"] = box_tail
                stats = {"events": length, "launches": 1 if done else 0, "ms_kernel": extra}
                if length > done:
                    main = launch(v, length - done, win, offset=done)
                    main["launches"] += stats["launches"]
                    main["ms_kernel"] += extra
                    stats = main
                stats["window"] = {"bins": wst.window["bins"], "rows": wst.window["rows"], "dense": wst.window["dense"]} if wst.window else None
                stats["window_coverage"] = wst.window_coverage
        kern, prog = used[1], used[2]
        if side:
            engine.stream_wait(side, 0)
        for st in stores:
            st.mark_filled()
        if kern.times_aux is not None and length > 0:
            stats["block_times"] = plan.block_times.read(numpy.uint64).reshape(-1, 3)[:stats.get("grid", 0)].tolist()
        if kern.fx_terms and _options.get("fx_stats", False):
            stats["fx_missed_total"] = int(plan.fx_stat.read(numpy.uint64)[0])
        stats["deterministic"] = kern.deterministic
        stats["sorted"] = kern.sorted
        stats["group_fast"] = bool(gfast)
        stats["regs"] = prog.regs
        stats["smem_bytes"] = kern.smem_bytes
        stats["keys_launches"] = 1 if (keys is not None and length > 0) else 0
        last_stats.clear()
        last_stats.update(stats)
        return length


class _Overlay(object):
    """``arrays`` with a few fields replaced (dict-like: __getitem__ raising KeyError is all bind_all needs)."""

    def __init__(self, base, replaced):
        self.base, self.replaced = base, replaced

    def __getitem__(self, name):
        if name in self.replaced:
            return self.replaced[name]
        return self.base[name]


def _encode_strings(plan, arrays):
    """String-valued groupby fields -> int64 codes (histbook_groupby on strings, calc/__init__.py:149-154)."""
    replaced = {}
    for field in plan.groupby_fields:
        try:
            value = arrays[field]
        except KeyError:
            continue
        if columns.is_stringlike(value):
            replaced[field] = plan.string_codes.setdefault(field, columns.StringCodes()).encode(value)
        elif field in plan.string_codes:
            raise codegen.UnsupportedError("field {0!r} was filled with strings before; a group axis cannot mix string and numeric keys".format(field))
    return _Overlay(arrays, replaced) if replaced else arrays


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
        self.ngroup = len(hspec["group"])
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
