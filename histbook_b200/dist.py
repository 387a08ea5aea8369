"""Multi-GPU: fill in parallel, then add.

The reference's only parallel model is "fill separately, combine with ``+``" (README.rst:29;
``Hist.__add__`` histbook/hist.py:506-542, ``GenericBook.__add__`` histbook/book.py:440-453): events are
independent and bin contents are a commutative sum.  Here: one process per GPU (``torch.distributed``, NCCL
over NVLink), every rank fills its own contiguous shard of the events into private device-resident
contents, and ONE collective per histogram adds the ranks -- ``all_reduce(SUM)`` on the float64 arena.
Histograms with group axes (dict contents) first agree on the union of their keys, in an order that is the
same on every rank, then reduce the re-laid-out slabs.

Nothing here touches the data path of a fill: there is no collective while events are being processed.
"""
import collections

import numpy


def shard(n, rank, world):
    """Contiguous event range [lo, hi) of rank ``rank`` out of ``world`` (sizes differ by at most one)."""
    return n * rank // world, n * (rank + 1) // world


def shard_columns(arrays, rank, world):
    """Slice every 1-d column to this rank's shard (scalars are passed through)."""
    n = None
    for v in arrays.values():
        if hasattr(v, "shape") and len(v.shape) == 1:
            n = int(v.shape[0])
            break
    if n is None:
        return dict(arrays)
    lo, hi = shard(n, rank, world)
    return dict((k, v[lo:hi] if hasattr(v, "shape") and len(v.shape) == 1 else v) for k, v in arrays.items())


def _dist():
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised (launch with torchrun / init_process_group)")
    return dist


# ----------------------------------------------------------------------------- host contents (any backend)

def union_keys(keysets):
    """Rank-independent order of the union of per-rank key lists: first-seen order over ranks 0..R-1."""
    out = collections.OrderedDict()
    for keys in keysets:
        for k in keys:
            out.setdefault(k, None)
    return list(out)


def allreduce_host(content, group=None):
    """Sum a host content object (ndarray | nested dict of ndarrays | None) over all ranks -- the distributed
    form of ``Hist.__add__``'s ``add`` (hist.py:513-537).  Works with any backend (gloo on CPU tensors)."""
    import torch
    dist = _dist()
    world = dist.get_world_size(group)

    def paths(node, prefix=()):
        if node is None:
            return []
        if isinstance(node, numpy.ndarray):
            return [(prefix, node)]
        out = []
        for k, v in node.items():
            out.extend(paths(v, prefix + (k,)))
        return out

    mine = paths(content)
    gathered = [None] * world
    dist.all_gather_object(gathered, [(p, a.shape) for p, a in mine], group=group)
    shapes = {}
    for rank_list in gathered:
        for p, shp in rank_list:
            shapes.setdefault(p, shp)
    order = union_keys([[p for p, _ in rank_list] for rank_list in gathered])
    if not order:
        return content
    have = dict(mine)
    flat = numpy.concatenate([(have[p] if p in have else numpy.zeros(shapes[p])).reshape(-1).astype(numpy.float64) for p in order])
    t = torch.from_numpy(flat)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, group=group)
    flat = t.cpu().numpy()
    if order == [()]:
        return flat.reshape(shapes[()])
    out = {} if not isinstance(content, collections.OrderedDict) else collections.OrderedDict()
    off = 0
    for p in order:
        size = int(numpy.prod(shapes[p]))
        node = out
        for k in p[:-1]:
            node = node.setdefault(k, {})
        node[p[-1]] = flat[off:off + size].reshape(shapes[p])
        off += size
    return out


# ----------------------------------------------------------------------------- one collective per Hist / Book

KEYCAP0 = 1024              # int64 words of the key exchange buffer (doubled by every rank alike when one rank needs more)
SCATTER_BYTES = 8 << 20     # contents at least this large are reduce-scattered, and gathered only when somebody reads them


def _hists_of(target):
    if isinstance(target, (list, tuple)):
        return list(target)
    plan = getattr(target, "__dict__", {}).get("_hb_plan")
    if isinstance(plan, tuple) and len(plan) == 3 and hasattr(target, "itervalues"):
        return plan[2]                                         # a Book that has been filled: the histogram list of its plan (no walk)
    if hasattr(target, "hists"):                               # fillable.SpecFill
        return list(target.hists)
    if hasattr(target, "itervalues"):                          # Book: one Hist stored under several names counts once
        seen = set()
        return [h for h in target.itervalues(recursive=True, onlyhist=True) if not (id(h) in seen or seen.add(id(h)))]
    return [target]


def _store(hist):
    st = hist.__dict__.get("_hb")
    if st is None:
        from histbook_b200.store import Store
        st = hist.__dict__["_hb"] = Store(hist._shape, [], None)
    return st


def _static_groups(hist, st):
    """Number of group axes of a histogram -- from its definition, not from what this rank happened to fill (a rank
    with an empty shard never ran the group discovery, yet has to join the same collectives as the others)."""
    g = getattr(hist, "_group", None)
    if g is not None:
        return len(g)
    if getattr(hist, "ngroup", None) is not None:              # fillable.SpecHist
        return hist.ngroup
    return len(st.groups)


def _host_tuples(st, content):
    """Nested host dict -> [(pattern tuple, flat block)] in insertion order (keys -> 64-bit patterns)."""
    out = []

    def walk(node, level, path):
        for key, sub in node.items():
            pattern = st.groups[level].pattern(key)
            if level + 1 == len(st.groups):
                out.append((path + (pattern,), numpy.asarray(sub, dtype=numpy.float64).reshape(-1)))
            else:
                walk(sub, level + 1, path + (pattern,))
    if content:
        walk(content, 0, ())
    return out


def nested_content(groups, order, blocks, shape):
    """Pattern tuples (in dict insertion order) + their blocks -> the nested dict ``Hist._content`` holds for group
    axes (histbook/hist.py:458-499): one level per group axis, keys as the reference would spell them."""
    def fresh(level):
        return collections.OrderedDict() if (groups[level].kind == "groupby" and groups[level].keeporder) else {}
    root = fresh(0)
    for tup, block in zip(order, blocks):
        node = root
        for level, pattern in enumerate(tup):
            key = groups[level].key_object(pattern)
            if level + 1 == len(tup):
                node[key] = numpy.array(block, dtype=numpy.float64).reshape(shape)
            else:
                if key not in node:
                    node[key] = fresh(level + 1)
                node = node[key]
    return root


class Merger(object):
    """The rank merge of a Hist, a Book (or a fillable.SpecFill) as ONE collective -- the distributed form of
    ``sum(hists)`` (``Hist.__add__`` histbook/hist.py:506-542, ``GenericBook.__add__`` histbook/book.py:440-453).

    ``start()`` snapshots every histogram's bins into one flat float64 buffer (device copies in order with the fill
    kernels; group-axis slabs are re-laid out on the device in the union order of all ranks' keys, agreed on with one
    fixed-size int64 all-gather) and enqueues one ``all_reduce`` -- or, for contents of ``scatter_bytes`` and more, one
    ``reduce_scatter`` that leaves the sum bin-sharded over the ranks until somebody reads it -- on a side stream, so
    the next fill overlaps it.  The rank's own histograms are not touched: like ``a + b`` the merge yields a new
    value (``contents()``), unless ``apply()`` writes it back (``allreduce`` below).

    Works on device-resident contents over NCCL and on host-resident contents over any backend (gloo on CPU: the
    world-size-2 tests)."""

    def __init__(self, target, group=None, scatter_bytes=SCATTER_BYTES, device=None):
        self.target = target
        self.group = group
        self.scatter_bytes = scatter_bytes
        self.keycap = KEYCAP0
        self.flat = None
        self.shard = None
        self.sharded = False
        self.side = None
        self.kstream = None             # stream of the key exchange
        self._kbuf = None               # ... and its buffers
        self.done = None
        self.device = device            # None: decided at the first start() (NCCL -> device, anything else -> host)
        self.layout = None              # [(hist, store, kind, offset, size, union order)] of the last start()
        self._sig = None                # what the last start() saw (stores, arenas, key counts): reuse of its layout
        self._entries = None
        self._unions = None
        self._fixed = None
        self.collectives = 0            # collectives issued so far
        self.collectives_last = 0       # ... by the last start()
        self._maps = {}

    # ------------------------------------------------------------------ helpers
    def _use_device(self):
        if self.device is None:
            self.device = _dist().get_backend(self.group) == "nccl"
        return self.device

    def _exchange_keys(self, mine, stringy):
        """mine: per group histogram a list of int64 pattern tuples; stringy: per histogram whether this rank knows its
        keys to be strings (dictionary codes are rank-local: those keys travel separately).  Returns (per histogram the
        union order over ranks -- first seen, ranks 0..R-1, the same list on every rank --, per histogram whether ANY
        rank called it stringy).  One all-gather of a fixed-size int64 tensor."""
        import torch
        dist = _dist()
        world = dist.get_world_size(self.group)
        while True:
            words = [len(mine)]
            for tuples, flag in zip(mine, stringy):
                words.extend([len(tuples), len(tuples[0]) if tuples else 0, 1 if flag else 0])
            for tuples in mine:
                for t in tuples:
                    words.extend(int(x) for x in t)
            need = len(words) + 1
            buf = numpy.zeros(self.keycap, dtype=numpy.uint64)
            buf[0] = need
            if need <= self.keycap:
                buf[1:need] = numpy.array(words, dtype=numpy.uint64)
            t = torch.from_numpy(buf.view(numpy.int64))
            if self._use_device():
                # on a stream of its own: the exchange must not queue behind the fill kernel that is still running on the
                # compute stream -- the host would wait for it and do the rest of the merge's host work with the GPU idle
                # (c4 at 8 GPUs: 10.5 ms per step against 8.8 ms of kernel before this)
                if self.kstream is None:
                    self.kstream = torch.cuda.Stream()
                if self._kbuf is None or self._kbuf[0].numel() != self.keycap:
                    # page-locked and device buffers of the exchange, allocated once (a cudaHostAlloc per merge costs more than the collective)
                    self._kbuf = (torch.empty(self.keycap, dtype=torch.int64).pin_memory(),
                                  torch.empty(self.keycap, dtype=torch.int64, device="cuda"),
                                  torch.empty(world * self.keycap, dtype=torch.int64, device="cuda"),
                                  torch.empty(world * self.keycap, dtype=torch.int64).pin_memory())
                pin, td, out, host = self._kbuf
                pin.copy_(t)
                with torch.cuda.stream(self.kstream):
                    td.copy_(pin, non_blocking=True)
                    dist.all_gather_into_tensor(out, td, group=self.group)
                    host.copy_(out, non_blocking=True)
                self.kstream.synchronize()
                rows = [r.view(numpy.uint64) for r in host.numpy().reshape(world, -1)]
            else:
                gathered = [torch.empty_like(t) for _ in range(world)]
                dist.all_gather(gathered, t, group=self.group)
                rows = [g.numpy().view(numpy.uint64) for g in gathered]
            self.collectives += 1
            most = max(int(r[0]) for r in rows)
            if most > self.keycap:                       # every rank sees the same numbers and grows alike
                while self.keycap < most:
                    self.keycap *= 2
                continue
            per_rank, flags = [], [False] * len(mine)
            for r in rows:
                pos = 1
                n = int(r[pos]); pos += 1
                if n != len(mine):
                    raise RuntimeError("ranks disagree on the histograms being merged ({0} vs {1} with group axes)".format(n, len(mine)))
                shapes = []
                for h in range(n):
                    shapes.append((int(r[pos]), int(r[pos + 1])))
                    flags[h] = flags[h] or bool(r[pos + 2])
                    pos += 3
                lists = []
                for count, k in shapes:
                    lists.append([tuple(int(x) for x in r[pos + i * k: pos + (i + 1) * k]) for i in range(count)])
                    pos += count * k
                per_rank.append(lists)
            return [union_keys([per_rank[r][h] for r in range(world)]) for h in range(len(mine))], flags

    def _exchange_wire(self, mine):
        """String group keys are dictionary codes local to a rank: they travel as the strings themselves (pickled)."""
        dist = _dist()
        world = dist.get_world_size(self.group)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine, group=self.group)
        self.collectives += 1
        return [union_keys([gathered[r][h] for r in range(world)]) for h in range(len(mine))]

    # ------------------------------------------------------------------ the merge
    def start(self):
        import torch
        dist = _dist()
        world = dist.get_world_size(self.group)
        device = self._use_device()
        hists = _hists_of(self.target)
        stores = [_store(h) for h in hists]
        before = self.collectives
        # repeated merges of a book that is being filled: the stores, their arenas and key lists are usually the same as
        # last time -- then everything below except the key exchange, the copies and the collective is reused
        sig = tuple((id(st), st.dev_valid, st.host_valid, None if st.arena is None else st.arena.ptr, len(st.tuples)) for st in stores) if device else None
        reuse = device and sig is not None and sig == self._sig
        self._sig = sig

        # 1. what every histogram holds here: fixed-size blocks, or (pattern tuple -> block) for group axes
        entries = self._entries[0] if reuse else []
        group_entries = self._entries[1] if reuse else []
        for h, st in ([] if reuse else zip(hists, stores)):
            ngroup = _static_groups(h, st)
            if ngroup == 0:
                if device:
                    st.to_device()
                entries.append([h, st, "fixed", None, None])
                continue
            if st.groups and device:
                st.to_device()
                tuples, blocks = list(st.tuples.keys()), None
            elif st.groups:
                pairs = _host_tuples(st, st.peek() if (st.host_valid or st.arena is not None) else None)
                tuples, blocks = [p for p, b in pairs], [b for p, b in pairs]
            else:
                if isinstance(st.host, dict) and st.host_valid and len(st.host) > 0:
                    raise RuntimeError("a group-axis histogram holds a host dict that never went through a fill here: use dist.allreduce_host for it")
                tuples, blocks = [], []                    # never filled here (no call of fill at all): joins with no keys
            group_entries.append(len(entries))
            entries.append([h, st, "group", tuples, blocks])

        # 2. agree on the union of the keys: one int64 all-gather; string keys (rank-local dictionary codes) travel pickled
        unions = {}
        if group_entries:
            stringy = [any(g.codes is not None for g in entries[i][1].groups) for i in group_entries]
            orders, flags = self._exchange_keys([[] if f else entries[i][3] for i, f in zip(group_entries, stringy)], stringy)
            for i, order in zip(group_entries, orders):
                unions[i] = order
            wire = [i for i, f in zip(group_entries, flags) if f]
            if wire:
                mine = []
                for i in wire:
                    st = entries[i][1]
                    mine.append([tuple(g.to_wire(p) for g, p in zip(st.groups, t)) for t in entries[i][3]] if st.groups else [])
                for i, order in zip(wire, self._exchange_wire(mine)):
                    st = entries[i][1]
                    unions[i] = [tuple(g.from_wire(w) for g, w in zip(st.groups, tup)) for tup in order] if st.groups else list(order)

        if not reuse:
            self._entries = (entries, group_entries)
        reuse = reuse and unions == self._unions and self.flat is not None
        self._unions = unions

        # 3. flat layout: [fixed blocks ...][group slabs in union order ...]
        off = 0
        layout = []
        for i, (h, st, kind, tuples, blocks) in enumerate(entries):
            if kind == "fixed":
                layout.append((h, st, kind, off, st.blocksize, None, tuples, blocks))
                off += st.blocksize
        for i, (h, st, kind, tuples, blocks) in enumerate(entries):
            if kind == "group":
                order = unions[i]
                layout.append((h, st, kind, off, len(order) * st.blocksize, order, tuples, blocks))
                off += len(order) * st.blocksize
        total = off
        scatter = world > 1 and device and total * 8 >= self.scatter_bytes
        padded = (total + world - 1) // world * world if scatter else total

        # 4. snapshot
        if device:
            from histbook_b200 import engine
            dev = torch.device("cuda", engine.init())
            if self.flat is None or self.flat.numel() != padded or not self.flat.is_cuda:
                self.flat = torch.empty(padded, device=dev, dtype=torch.float64)
                self._fixed, reuse = None, False
            cur = torch.cuda.current_stream(dev)
            if self.done is not None:
                cur.wait_event(self.done)                  # the previous collective has read the buffer
            # ONE launch copies every arena -- group-axis slabs slot by slot into the union order -- into the buffer
            # (hb_arena_snapshot; 56 + 8 small torch copies were ~0.4 ms of launches per merge of c4)
            if not (reuse and self._fixed is not None):
                segments, holes = [], []
                for (h, st, kind, off_, size, order, tuples, blocks) in layout:
                    if kind == "fixed":
                        segments.append((st.arena, 0, off_, size))
                    elif size:
                        bs = st.blocksize
                        have = dict((t, slot) for slot, t in enumerate(tuples))
                        for j, t in enumerate(order):
                            if t in have:
                                segments.append((st.arena, have[t] * bs, off_ + j * bs, bs))
                            else:
                                holes.append((off_ + j * bs, bs))          # a key only other ranks hold: zeros here
                self._fixed = (engine.Snapshot(segments), holes)
            snap, holes = self._fixed
            for o, c in holes:
                self.flat[o:o + c].zero_()
            snap.run(self.flat.data_ptr())
            if padded > total:
                self.flat[total:].zero_()
        else:
            flat = numpy.zeros(padded, dtype=numpy.float64)
            for (h, st, kind, off_, size, order, tuples, blocks) in layout:
                if kind == "fixed":
                    c = st.peek() if (st.host_valid or st.arena is not None) else None
                    if c is not None:
                        flat[off_:off_ + size] = numpy.asarray(c, dtype=numpy.float64).reshape(-1)
                else:
                    have = dict((t, b) for t, b in zip(tuples, blocks))
                    for j, t in enumerate(order):
                        if t in have:
                            flat[off_ + j * st.blocksize: off_ + (j + 1) * st.blocksize] = have[t]
            self.flat = torch.from_numpy(flat)

        # 5. ONE collective, on a side stream: the next fill overlaps it
        self.layout = layout
        self.total = total
        self.sharded = False
        if world > 1 or device:
            if device:
                if self.side is None:
                    self.side = torch.cuda.Stream(device=self.flat.device)
                ready = torch.cuda.Event()
                ready.record(torch.cuda.current_stream(self.flat.device))
                self.side.wait_event(ready)
                with torch.cuda.stream(self.side):
                    if scatter:
                        if self.shard is None or self.shard.numel() != padded // world:
                            self.shard = torch.empty(padded // world, device=self.flat.device, dtype=torch.float64)
                        dist.reduce_scatter_tensor(self.shard, self.flat, group=self.group)
                        self.sharded = True
                    else:
                        dist.all_reduce(self.flat, group=self.group)
                    self.done = torch.cuda.Event()
                    self.done.record(self.side)
            else:
                dist.all_reduce(self.flat, group=self.group)
            self.collectives += 1
        self.collectives_last = self.collectives - before
        return self

    def wait(self):
        """Order the caller's stream after the collective (device side; the host does not block)."""
        if self.done is not None:
            import torch
            torch.cuda.current_stream(self.flat.device).wait_event(self.done)
        return self

    def _gathered(self):
        """The whole merged buffer on this rank (all-gather of the shards after a reduce-scatter)."""
        import torch
        self.wait()
        if self.sharded:
            _dist().all_gather_into_tensor(self.flat, self.shard, group=self.group)
            self.collectives += 1
            self.sharded = False
        return self.flat

    def contents(self):
        """Host contents of the merged histograms, in the order of the target's histograms: ndarray, or the nested dict
        of a group-axis histogram."""
        flat = self._gathered()
        host = flat.cpu().numpy() if flat.is_cuda else flat.numpy()
        out = {}
        for (h, st, kind, off, size, order, tuples, blocks) in self.layout:
            if kind == "fixed":
                out[id(h)] = host[off:off + size].reshape(st.shape).copy()
            else:
                bs = st.blocksize
                out[id(h)] = nested_content(st.groups, order, [host[off + j * bs: off + (j + 1) * bs] for j in range(len(order))], st.shape) if st.groups else {}
        return [out[id(h)] for h in _hists_of(self.target)]

    def apply(self):
        """Write the merged bins back into the target's histograms: afterwards every rank holds the total."""
        import torch
        flat = self._gathered()
        if not flat.is_cuda:
            host = flat.numpy()
            for (h, st, kind, off, size, order, tuples, blocks) in self.layout:
                if kind == "fixed":
                    st.host = host[off:off + size].reshape(st.shape).copy()
                elif st.groups:
                    bs = st.blocksize
                    st.host = nested_content(st.groups, order, [host[off + j * bs: off + (j + 1) * bs] for j in range(len(order))], st.shape)
                else:
                    continue
                st.host_valid, st.dev_valid = True, False
            return self
        fixed = [(off, size, st) for (h, st, kind, off, size, order, tuples, blocks) in self.layout if kind == "fixed"]
        if fixed:
            torch._foreach_copy_([st.arena.as_tensor() for off, size, st in fixed], [flat[off:off + size] for off, size, st in fixed])
            for off, size, st in fixed:
                st.mark_filled()
        for (h, st, kind, off, size, order, tuples, blocks) in self.layout:
            if kind != "group":
                continue
            if not st.groups:
                if size:
                    raise RuntimeError("this rank never filled a group-axis histogram the other ranks hold keys for: call fill "
                                       "(empty columns will do) on every rank before merging, so that the key types are known")
                continue
            if st.arena is None:
                st.to_device()
            if st.arena.size < max(size, 1):
                st.arena.resize(max(size, 1))
            if size:
                st.arena.as_tensor()[:size].copy_(flat[off:off + size])
            st.set_tuples(order)
            self._maps.pop(id(st), None)
            st.mark_filled()
        return self


def allreduce(target, group=None):
    """Add the contents of ``target`` (a histbook Hist or Book, or a fillable.SpecFill) over all ranks, in place:
    afterwards every rank holds the total, like ``sum(hists)`` in the reference.  One collective for the whole book
    (plus one small all-gather of the group keys when it has group axes)."""
    hists = _hists_of(target)
    rest = []
    for h in hists:
        st = _store(h)
        if not st.groups and isinstance(st.host, dict) and st.host_valid:
            st.host = allreduce_host(st.host, group)       # dict content that never went through a fill: reduce on the host
        else:
            rest.append(h)
    return Merger(rest, group=group, scatter_bytes=1 << 62).start().apply()
