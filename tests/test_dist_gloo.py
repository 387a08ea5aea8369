"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: event sharding and the rank merge.

Each rank fills its shard with the CPU oracle (standing in for the device fill, which needs a GPU) and the
ranks are merged with histbook_b200.dist.allreduce_host -- the same union-of-keys + all_reduce(SUM) logic the
NCCL path uses for device arenas.  The merged result must equal one fill of all events (bit-exact for counts).
"""
import os
import socket
import sys

import numpy
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, case_name, out_dir):
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import golden_util
    from histbook_b200 import dist as hbdist
    from oracle import fill_oracle
    case = golden_util.load()[case_name]
    arrays = golden_util.fills(case)[0]
    mine = hbdist.shard_columns(arrays, rank, world)
    contents = fill_oracle.fill(case["spec"], mine)
    merged = [hbdist.allreduce_host(c) for c in contents]
    flat = []
    for c in merged:
        flat.append([(k, a) for k, a in golden_util.flatten(c)])
    numpy.save(os.path.join(out_dir, "rank%d.npy" % rank), numpy.array(flat, dtype=object), allow_pickle=True)
    dist.destroy_process_group()


def test_shard_ranges_partition_the_events():
    from histbook_b200 import dist as hbdist
    for n in (0, 1, 7, 1000, 10**9 + 7):
        for world in (1, 2, 3, 8):
            ranges = [hbdist.shard(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1


def test_union_keys_is_rank_independent():
    from histbook_b200 import dist as hbdist
    assert hbdist.union_keys([[(1,), (3,)], [(2,), (1,)], []]) == [(1,), (3,), (2,)]


@pytest.mark.parametrize("case_name", ["c1", "c2", "c3", "groupbin_nan_refill", "groupby_groupbin_num", "c4"])
def test_two_ranks_fill_then_add(case_name, tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import golden_util
    from oracle import fill_oracle
    port = _free_port()
    mp.spawn(_worker, args=(2, port, case_name, str(tmp_path)), nprocs=2, join=True)
    case = golden_util.load()[case_name]
    arrays = golden_util.fills(case)[0]
    whole = fill_oracle.fill(case["spec"], arrays)
    bound = fill_oracle.fill(case["spec"], arrays, absolute=True)
    r0 = numpy.load(os.path.join(str(tmp_path), "rank0.npy"), allow_pickle=True)
    r1 = numpy.load(os.path.join(str(tmp_path), "rank1.npy"), allow_pickle=True)
    for hid, (w, b) in enumerate(zip(whole, bound)):
        fw, fb = dict(golden_util.flatten(w)), dict(golden_util.flatten(b))
        for merged in (r0[hid], r1[hid]):
            got = dict((tuple(k), a) for k, a in merged)
            # a rank only knows the keys of its own shard; the union over ranks must be at least the keys that got entries
            assert set(fw) >= set(got) or set(got) >= set(k for k in fw if fw[k].any())
            for k in fw:
                if k not in got:
                    assert not fw[k].any()
                    continue
                assert numpy.all(numpy.abs(got[k] - fw[k]) <= 1e-12 * fb[k] + 0.0)
                if case["spec"]["hists"][hid]["weight"] is None:
                    sumw = case["spec"]["hists"][hid]["sumw"]
                    assert numpy.array_equal(got[k][..., sumw], fw[k][..., sumw])


# ------------------------------------------------------------------------------------------------
# dist.Merger: the one-collective merge of a whole book (host-resident contents, gloo)

def _host_filler(spec, contents, arrays):
    """A fillable.SpecFill whose stores hold ``contents`` on the host, with the group-axis descriptions a device fill
    would have recorded (codegen only, no GPU)."""
    from histbook_b200 import codegen, fillable, spec as hbspec
    from histbook_b200.store import GroupInfo, Store
    sf = fillable.SpecFill(spec)
    coltypes = dict((n, (numpy.asarray(arrays[n]).dtype, numpy.ndim(arrays[n]) == 0)) for n in sf.fields)
    keys = codegen.generate_keys(spec, coltypes)
    infos = dict(keys.group_goals) if keys is not None else {}
    for h, hspec, c in zip(sf.hists, spec["hists"], contents):
        groups = []
        for a in hspec["group"]:
            gk = hbspec.tree_key(a["goal"])
            info = infos[gk]
            groups.append(GroupInfo(a["kind"], a["keeporder"], info["kind"] == "f", info["dtype"], info["nanflow"], gk, None))
        h.__dict__["_hb"] = Store(hspec["shape"], groups, c)
    return sf


def _merger_worker(rank, world, port, case_name, out_dir, lazy_rank, emulated=False):
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import contextlib
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import golden_util
    from histbook_b200 import dist as hbdist, fillable
    from oracle import fill_oracle
    case = golden_util.load()[case_name]
    arrays = golden_util.fills(case)[0]
    engine_scope = contextlib.nullcontext()
    if emulated:
        import emul_engine
        engine_scope = emul_engine.installed()            # the rank's fill really runs: generated kernels on CPU threads
    with engine_scope:
        if lazy_rank == rank:
            sf = fillable.SpecFill(case["spec"])              # this rank never calls fill: no store at all
        else:
            if lazy_rank is None:
                mine = hbdist.shard_columns(arrays, rank, world)
            else:
                mine = arrays                                 # the other rank fills everything
            if emulated:
                sf = fillable.SpecFill(case["spec"])
                sf.fill(mine)
            else:
                sf = _host_filler(case["spec"], fill_oracle.fill(case["spec"], mine), mine)
        merger = hbdist.Merger(sf).start()
        merged = merger.contents()
    flat = [[(k, a) for k, a in golden_util.flatten(c)] for c in merged]
    numpy.save(os.path.join(out_dir, "rank%d.npy" % rank), numpy.array([flat, merger.collectives], dtype=object), allow_pickle=True)
    dist.destroy_process_group()


@pytest.mark.parametrize("case_name,lazy_rank,emulated", [("c2", None, False), ("groupbin_nan_refill", None, False), ("groupby_groupbin_num", None, False),
                                                          ("c4", None, False), ("c4", 1, False), ("book_fill", None, False),
                                                          ("c4", None, True), ("c4", 1, True)])
def test_merger_is_one_collective_and_equals_one_fill(case_name, lazy_rank, emulated, tmp_path):
    """Two ranks, contents on the host, gloo: the whole book merges with ONE all_reduce (plus one key all-gather when it
    has group axes), the result equals one fill of everything, and a rank that never filled (``lazy_rank``) joins the
    same collectives instead of hanging the others.  ``emulated``: the ranks do not take their contents from the oracle but
    fill their shards through the product's host path on the emulated engine (tests/emul_engine.py) -- shard, fill, merge, as
    under torchrun on the GPUs, with the kernels on CPU threads and gloo for NCCL."""
    import shutil
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import golden_util
    from oracle import fill_oracle
    if emulated and shutil.which("g++") is None:
        pytest.skip("the kernel emulator is compiled with g++")
    port = _free_port()
    mp.spawn(_merger_worker, args=(2, port, case_name, str(tmp_path), lazy_rank, emulated), nprocs=2, join=True)
    case = golden_util.load()[case_name]
    arrays = golden_util.fills(case)[0]
    whole = fill_oracle.fill(case["spec"], arrays)
    bound = fill_oracle.fill(case["spec"], arrays, absolute=True)
    has_groups = any(h["group"] for h in case["spec"]["hists"])
    for r in range(2):
        flat, collectives = numpy.load(os.path.join(str(tmp_path), "rank%d.npy" % r), allow_pickle=True)
        assert collectives == (2 if has_groups else 1)
        if r == lazy_rank:
            continue                                      # cannot name the keys (it never saw a column), but did not hang
        for hid, (w, b) in enumerate(zip(whole, bound)):
            fw, fb = dict(golden_util.flatten(w)), dict(golden_util.flatten(b))
            got = dict((tuple(k), a) for k, a in flat[hid])
            for k in fw:
                if k not in got:
                    assert not fw[k].any()
                    continue
                assert numpy.all(numpy.abs(got[k] - fw[k]) <= 1e-12 * fb[k] + 0.0)
                if case["spec"]["hists"][hid]["weight"] is None:
                    sumw = case["spec"]["hists"][hid]["sumw"]
                    assert numpy.array_equal(got[k][..., sumw], fw[k][..., sumw])
