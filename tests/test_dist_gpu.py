"""One process per GPU (NCCL): fill in parallel, then ``dist.allreduce`` -- the device form of the reference's
"fill in parallel, then add" (README.rst:29, Hist.__add__ histbook/hist.py:506-542).  The two-rank case is skipped with fewer than two GPUs
(``gpurun --gpus 2 -- python -m pytest tests/test_dist_gpu.py -m gpu``); the host logic of the same merge is covered on
CPU by tests/test_dist_gloo.py.

Each rank fills its contiguous shard of the events on its own GPU; after the merge every rank must hold the result of
one fill of all events: counts bit-exact, float sums within 1e-12 of the per-bin sum of |term|, group keys (numeric
and string) in the union over ranks.
"""
import os
import socket
import sys

import numpy
import pytest

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _columns(n, seed):
    rs = numpy.random.RandomState(seed)
    names = numpy.array(["ttbar", "wjets", "zjets", "qcd", "data", "rare"])
    # "rare" only occurs in the second half: the ranks see different key sets
    c = names[numpy.where(numpy.arange(n) < n // 2, rs.randint(0, 5, n), rs.randint(0, 6, n))]
    return {"x": rs.normal(0, 1, n), "y": rs.normal(0, 1, n), "w": rs.uniform(0.5, 1.5, n), "n": rs.poisson(3.0, n).astype(numpy.int64), "c": c}


def _make(hb):
    return [hb.Hist(hb.bin("x", 100, -5, 5)),
            hb.Hist(hb.bin("x + y", 200, -5, 5), hb.bin("sqrt(x**2 + y**2)", 200, 0, 5), weight="w"),
            hb.Hist(hb.bin("x", 1000, -5, 5), hb.cut("x > 0"), hb.profile("y"), hb.profile("y**2")),
            hb.Hist(hb.groupbin("n", 2), hb.bin("x", 50, -5, 5)),
            hb.Hist(hb.groupby("c"), hb.bin("x", 20, -3, 3), weight="w")]


def _worker(rank, world, port, n, out_dir):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import histbook_b200
    from histbook_b200 import dist as hbdist, engine
    engine.init(rank)
    hb = histbook_b200.reference()
    cols = _columns(n, 5)
    lo, hi = hbdist.shard(n, rank, world)
    mine = dict((k, v[lo:hi]) for k, v in cols.items())
    book = hb.Book(dict(("h%d" % i, h) for i, h in enumerate(_make(hb))))
    book.fill(**mine)
    hbdist.allreduce(book)
    out = {}
    for name in sorted(book.keys()):
        c = book[name]._content
        out[name] = c if isinstance(c, numpy.ndarray) else dict((str(k), v) for k, v in c.items())
    numpy.save(os.path.join(out_dir, "rank%d.npy" % rank), numpy.array([out], dtype=object), allow_pickle=True)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2])
def test_fill_then_allreduce(tmp_path, world):
    """world = 1 runs everywhere (one rank, NCCL group of one: the zero-copy arena tensor and the collective call are
    still exercised); world = 2 needs two GPUs."""
    if _gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    sys.path.insert(0, REPO)
    import histbook_b200
    n = 400003
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    cols = _columns(n, 5)
    with histbook_b200.reference_path() as hb:              # one fill of everything on the reference's numpy path
        ref = _make(hb)
        expected, bound = {}, {}
        for i, h in enumerate(ref):
            h.fill(**cols)
            c = h._content
            expected["h%d" % i] = c if isinstance(c, numpy.ndarray) else dict((str(k), v) for k, v in c.items())
    ranks = [numpy.load(os.path.join(str(tmp_path), "rank%d.npy" % r), allow_pickle=True)[0] for r in range(world)]
    for got in ranks:
        assert sorted(got) == sorted(expected)
        for name, exp in expected.items():
            g = got[name]
            if isinstance(exp, dict):
                assert sorted(g) == sorted(exp), name
                pairs = [(g[k], exp[k]) for k in exp]
            else:
                pairs = [(g, exp)]
            for a, e in pairs:
                assert a.shape == e.shape
                tol = 1e-12 * numpy.maximum(numpy.abs(e), 1.0) * 50        # |sum| is a lower bound of sum|term|: generous factor, still far below a count
                assert numpy.all(numpy.abs(a - e) <= tol), name
        # unweighted histograms: bit-exact
        assert numpy.array_equal(got["h0"], expected["h0"])
        for k in expected["h3"]:
            assert numpy.array_equal(got["h3"][k], expected["h3"][k])
    # all ranks hold the same thing, bit for bit
    for name in expected:
        a, b = ranks[0][name], ranks[-1][name]
        if isinstance(a, dict):
            for k in a:
                assert numpy.array_equal(a[k], b[k])
        else:
            assert numpy.array_equal(a, b)
