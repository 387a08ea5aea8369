"""Device-resident bin contents behind ``Hist._content``.

The reference keeps ``Hist._content`` as ``None`` (never filled), a float64 ndarray of shape
``hist.shape``, or -- with group axes -- a nested dict of such arrays keyed by group value
(histbook/hist.py:72-99,256,381-390).  Everything on the read side touches the bins only through that
attribute (histbook/proj.py:129-135,498-640; histbook/hist.py:506-607,700-741), so it is the seam: here
it becomes a property over a ``Store`` that owns a float64 arena on the device and hands out / takes in
host copies on demand.

Coherence (deliberately pessimistic, because callers mutate the host object in place, e.g.
``self._content += other._content`` in hist.py:561): reading ``_content`` downloads and makes the *host*
copy authoritative; the next ``fill`` uploads it again.  Back-to-back fills never leave the device.
"""
import collections

import numpy

from histbook_b200 import engine
from histbook_b200.codegen import UnsupportedError

KEY_NAN = 0x7FF8000000000000        # hb_template.cuh HB_KEY_NAN
KEY_EMPTY = 0xFFF0DEADBEEFCAFE


class GroupInfo(object):
    """Static description of one group axis of a histogram."""
    __slots__ = ("kind", "keeporder", "floatkey", "dtype", "nanflow", "goalkey", "codes")

    def __init__(self, kind, keeporder, floatkey, dtype, nanflow, goalkey, codes=None):
        self.codes = codes            # columns.StringCodes when the axis groups by a string field (patterns are codes)
        self.kind = kind              # "groupby" | "groupbin"
        self.keeporder = keeporder
        self.floatkey = floatkey      # True: pattern is float64 bits; False: int64 bits
        self.dtype = numpy.dtype(dtype)
        self.nanflow = nanflow
        self.goalkey = goalkey

    def key_object(self, pattern):
        """64-bit pattern -> the dict key the reference would use (calc/__init__.py:150-154,170-187)."""
        if self.codes is not None:
            return self.codes.decode(pattern)
        if self.floatkey:
            if pattern == KEY_NAN:
                return "NaN" if self.kind == "groupbin" else self.dtype.type(numpy.nan)
            value = numpy.array([pattern], dtype=numpy.uint64).view(numpy.float64)[0]
            return self.dtype.type(value)
        value = numpy.array([pattern], dtype=numpy.uint64).view(numpy.int64)[0]
        return self.dtype.type(value)

    def pattern(self, key):
        """dict key -> 64-bit pattern (inverse of key_object)."""
        if self.codes is not None:
            if not isinstance(key, (str, bytes)):
                raise UnsupportedError("a group axis cannot mix string and numeric keys")
            return self.codes.code(key)
        if isinstance(key, str):
            if key == "NaN" and self.floatkey:
                return KEY_NAN
            raise UnsupportedError("string group keys are not supported by the B200 fill backend")
        if self.floatkey:
            v = numpy.float64(key)
            if v != v:
                return KEY_NAN
            if v == 0.0:
                v = numpy.float64(0.0)
            return int(numpy.array([v]).view(numpy.uint64)[0])
        return int(numpy.array([int(key)], dtype=numpy.int64).view(numpy.uint64)[0])

    def to_wire(self, pattern):
        """What identifies this key on another rank (string codes are rank-local)."""
        return ("s", self.codes.strings[int(pattern)]) if self.codes is not None else int(pattern)

    def from_wire(self, wire):
        return self.codes.code(wire[1]) if self.codes is not None else int(wire)

    def sort_patterns(self, patterns):
        """Order numpy.unique would give the keys of one fill: ascending by value, NaN last."""
        if self.codes is not None:
            return [int(p) for p in sorted(patterns, key=lambda p: self.codes.strings[int(p)])]
        arr = numpy.array(sorted(patterns), dtype=numpy.uint64)
        if self.floatkey:
            vals = arr.view(numpy.float64)
            order = numpy.argsort(vals, kind="stable")        # NaN sorts last
        else:
            order = numpy.argsort(arr.view(numpy.int64), kind="stable")
        return [int(x) for x in arr[order]]


class Store(object):
    """Contents of one histogram: host object and/or device arena."""

    def __init__(self, shape, groups, host=None):
        self.shape = tuple(shape)
        self.blocksize = int(numpy.prod(self.shape, dtype=numpy.int64))
        self.groups = groups              # [GroupInfo]
        self.host = host                  # None | ndarray | nested dict
        self.host_valid = True
        self.arena = None
        self.dev_valid = False
        # group bookkeeping (device side)
        self.tuples = collections.OrderedDict()      # tuple(patterns) -> slot
        self.skeleton = None                          # nested OrderedDict pattern -> ... -> slot, reference insertion order
        self.table = None                             # (fill key lists, engine.DeviceBuffer, the map's bytes): dense slot map of the last fill's key combinations
        # hot-window privatisation (fillable.choose_window): None = not placed yet, False = not worth it
        self.window = None            # {"table": DeviceBuffer, "bins", "rows"}
        self.window_key = None        # (rows, trail, wmax) of the kernel the window was placed for
        self.window_coverage = None
        self.window_seen = None       # |sum of weights| per (row, trailing bin) when the window was placed or last checked (fillable._watch_window)
        self.window_fills = 0         # fills / events since then
        self.window_events = 0
        self._admitted = None         # the key lists of the last fill (admit)
        self.exact = None             # deterministic mode: engine.Arena of 128-bit accumulators (zero between fills)
        self.nonempty = False                         # the device copy holds at least one fill / uploaded content

    # ------------------------------------------------------------------ host view
    def get(self):
        if not self.host_valid:
            self.host = self._download()
            self.host_valid = True
            self.dev_valid = False        # the caller may mutate what we hand out
        return self.host

    def peek(self):
        """Host copy without giving up the device copy (the returned object must not be mutated)."""
        if self.host_valid:
            return self.host
        return self._download()

    # ------------------------------------------------------------------ device side
    def to_device(self):
        """Make the device copy current (upload or zero-fill); called before every fill."""
        if self.dev_valid:
            return
        if not self.groups:
            if self.arena is None:
                self.arena = engine.Arena(self.blocksize)
                if self.host is not None:
                    self.arena.write(self._check(self.host))
            elif self.host is None:
                self.arena.clear()
            else:
                self.arena.write(self._check(self.host))
            self.nonempty = self.host is not None
            self.window = None
            self.window_seen = None
        else:
            self._upload_groups()
        self.dev_valid = True

    def mark_filled(self):
        self.host = None
        self.host_valid = False
        self.dev_valid = True
        self.nonempty = True

    def _check(self, array):
        array = numpy.asarray(array, dtype=numpy.float64)
        if array.shape != self.shape:
            raise ValueError("histogram content has shape {0}, expected {1}".format(array.shape, self.shape))
        return array

    def _download(self):
        if not self.groups:
            return self.arena.read().reshape(self.shape)
        flat = self.arena.read(0, len(self.tuples) * self.blocksize) if len(self.tuples) else numpy.empty(0)

        def build(node, level):
            out = collections.OrderedDict() if (self.groups[level].kind == "groupby" and self.groups[level].keeporder) else {}
            for pattern, sub in node.items():
                key = self.groups[level].key_object(pattern)
                if level + 1 == len(self.groups):
                    out[key] = flat[sub * self.blocksize:(sub + 1) * self.blocksize].reshape(self.shape).copy()
                else:
                    out[key] = build(sub, level + 1)
            return out
        return build(self.skeleton if self.skeleton is not None else {}, 0)

    # ------------------------------------------------------------------ group axes
    def _upload_groups(self):
        """Nested host dict -> slab + skeleton (after the host object was set or mutated)."""
        self.tuples = collections.OrderedDict()
        self.skeleton = collections.OrderedDict()
        blocks = []

        def walk(node, level, path, skel):
            for key, sub in node.items():
                pattern = self.groups[level].pattern(key)
                if level + 1 == len(self.groups):
                    slot = len(blocks)
                    blocks.append(self._check(sub).reshape(-1))
                    self.tuples[path + (pattern,)] = slot
                    skel[pattern] = slot
                else:
                    skel[pattern] = collections.OrderedDict()
                    walk(sub, level + 1, path + (pattern,), skel[pattern])
        if self.host is not None:
            walk(self.host, 0, (), self.skeleton)
        n = max(len(blocks), 1) * self.blocksize
        if self.arena is None:
            self.arena = engine.Arena(n)
        else:
            self.arena.clear()
            if self.arena.size < n:
                self.arena.resize(n)
        if blocks:
            self.arena.write(numpy.concatenate(blocks))
        self.table = None

    def set_tuples(self, order):
        """Adopt ``order`` (key-pattern tuples, slot = position) as the slab layout: tuples, nested skeleton in the
        reference's insertion order, lookup table dropped."""
        self.tuples = collections.OrderedDict((tup, slot) for slot, tup in enumerate(order))
        skeleton = collections.OrderedDict()
        for tup, slot in self.tuples.items():
            node = skeleton
            for k in tup[:-1]:
                node = node.setdefault(k, collections.OrderedDict())
            node[tup[-1]] = slot
        self.skeleton = skeleton
        self.table = None

    def admit(self, fill_keys):
        """Register the keys one fill produced.  ``fill_keys[a]`` = patterns seen on axis ``a`` in this fill.

        hist.py:458-499: the reference loops over *all* uniques of every axis at every nesting level and
        creates missing entries, i.e. after a fill the dict holds the cross product of this fill's per-axis
        key sets (zero blocks where no row landed).  Insertion order = numpy.unique order per axis."""
        if self.table is not None and self.arena is not None and self._admitted == fill_keys:
            return                                   # the same keys as the last fill: slots, slab and slot map stand
        if self.skeleton is None:
            self.skeleton = collections.OrderedDict()
        ordered = [g.sort_patterns(k) for g, k in zip(self.groups, fill_keys)]
        before = len(self.tuples)

        def walk(level, path, skel):
            for pattern in ordered[level]:
                if level + 1 == len(self.groups):
                    if pattern not in skel:
                        slot = len(self.tuples)
                        self.tuples[path + (pattern,)] = slot
                        skel[pattern] = slot
                else:
                    if pattern not in skel:
                        skel[pattern] = collections.OrderedDict()
                    walk(level + 1, path + (pattern,), skel[pattern])
        walk(0, (), self.skeleton)
        need = max(len(self.tuples), 1) * self.blocksize
        if self.arena.size < need:
            grow = max(need, 2 * self.arena.size)
            self.arena.resize(grow)
        self._slot_map(ordered)
        self._admitted = [list(k) for k in fill_keys]

    def _slot_map(self, ordered):
        """Dense map for the fill kernel (codegen._group_lookup): position of the event's key on every axis among THIS
        fill's keys (ascending unsigned patterns, the order of the goal tables fillable.fill_hists uploads) -> slab
        slot.  ``ordered`` is the per-axis key lists in any order; rebuilt only when the fill's key lists change."""
        axes = tuple(tuple(sorted(set(int(p) for p in keys))) for keys in ordered)
        if self.table is not None and self.table[0] == axes:
            return
        sizes = [len(x) for x in axes]
        total = int(numpy.prod(sizes, dtype=numpy.int64)) if sizes else 0
        slotmap = numpy.full(max(total, 1), -1, dtype=numpy.int32)
        if total:
            import itertools
            for comb, tup in enumerate(itertools.product(*axes)):
                slotmap[comb] = self.tuples[tup]
        self.table = (axes, engine.DeviceBuffer(slotmap), slotmap.tobytes())
