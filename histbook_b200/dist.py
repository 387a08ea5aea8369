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


# ----------------------------------------------------------------------------- device contents (NCCL)

def _allreduce_store(st, group):
    dist = _dist()
    if st.arena is None or not st.dev_valid:
        st.to_device()
    if not st.groups:
        dist.all_reduce(st.arena.as_tensor(), group=group)
        st.mark_filled()
        return
    # group axes: agree on the union of key tuples, re-lay the slab out in that order, then reduce
    world = dist.get_world_size(group)
    gathered = [None] * world
    # (string group keys are dictionary codes local to a rank: they travel as the strings themselves)
    wire = [tuple(g.to_wire(p) for g, p in zip(st.groups, tup)) for tup in st.tuples.keys()]
    dist.all_gather_object(gathered, wire, group=group)
    order = [tuple(g.from_wire(w) for g, w in zip(st.groups, tup)) for tup in union_keys(gathered)]
    host = st.arena.read(0, len(st.tuples) * st.blocksize).reshape(len(st.tuples), st.blocksize) if st.tuples else numpy.zeros((0, st.blocksize))
    slab = numpy.zeros((max(len(order), 1), st.blocksize))
    for slot, tup in enumerate(order):
        if tup in st.tuples:
            slab[slot] = host[st.tuples[tup]]
    if st.arena.size < slab.size:
        st.arena.resize(slab.size)
    st.arena.write(slab.reshape(-1))
    st.tuples = collections.OrderedDict((tup, slot) for slot, tup in enumerate(order))
    skeleton = collections.OrderedDict()
    for tup, slot in st.tuples.items():
        node = skeleton
        for k in tup[:-1]:
            node = node.setdefault(k, collections.OrderedDict())
        node[tup[-1]] = slot
    st.skeleton = skeleton
    st.table = None
    t = st.arena.as_tensor()[:slab.size]
    dist.all_reduce(t, group=group)
    st.mark_filled()


def allreduce(target, group=None):
    """Add the contents of ``target`` (a histbook Hist or Book, or a fillable.SpecFill) over all ranks, in
    place and on the device: afterwards every rank holds the total, like ``sum(hists)`` in the reference."""
    hists = None
    if hasattr(target, "hists"):                               # SpecFill
        hists = list(target.hists)
    elif hasattr(target, "itervalues"):                        # Book
        seen = set()
        hists = [h for h in target.itervalues(recursive=True, onlyhist=True) if not (id(h) in seen or seen.add(id(h)))]
    else:
        hists = [target]
    for h in hists:
        st = h.__dict__.get("_hb")
        if st is None:
            from histbook_b200.store import Store
            st = h.__dict__["_hb"] = Store(h._shape, [], None)
        if st.groups or (isinstance(st.host, dict) and st.host_valid):
            if not st.groups:
                # dict content that never went through a device fill: reduce on the host
                st.host = allreduce_host(st.host, group)
                continue
        _allreduce_store(st, group)
