"""histbook_b200.ingest: fills from Parquet / Arrow IPC / .npy files and from iterables of batches, batch by batch through the
product's host path on the emulated engine (tests/emul_engine.py), against the oracle fed the same batches."""
import shutil

import numpy
import pytest

import emul_engine
import golden_util
from oracle import fill_oracle
from test_gpu_parity import compare

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="the kernel emulator is compiled with g++")


def _columns(n, seed):
    rs = numpy.random.RandomState(seed)
    return {"x": rs.normal(0, 1, n), "y": rs.normal(0, 1, n), "z": rs.normal(0, 1, n), "w": rs.uniform(0.5, 1.5, n),
            "n": rs.poisson(3, n).astype(numpy.int64), "unused": rs.normal(0, 1, n)}


def _expect(spec, batches):
    exp = bound = None
    for arrays in batches:
        mine = dict((k, arrays[k]) for k in ("x", "y", "z", "w", "n"))
        exp = fill_oracle.fill(spec, mine, exp)
        bound = fill_oracle.fill(spec, mine, bound, absolute=True)
    return [golden_util.flatten(e) for e in exp], bound


def _slices(cols, bounds):
    return [dict((k, v[a:b]) for k, v in cols.items()) for a, b in zip(bounds, bounds[1:])]


def test_fill_parquet_in_row_batches(tmp_path):
    pa = pytest.importorskip("pyarrow")
    import pyarrow.parquet as pq
    from histbook_b200 import fillable, ingest
    spec = golden_util.load()["c4"]["spec"]
    cols = _columns(7000, 1)
    path = tmp_path / "events.parquet"
    pq.write_table(pa.table(cols), str(path), row_group_size=2500)
    with emul_engine.installed():
        sf = fillable.SpecFill(spec)
        nb, nev = ingest.fill_parquet(sf, str(path), batch_rows=2500)
        got = sf.contents()
    assert (nb, nev) == (3, 7000)
    exp, bound = _expect(spec, _slices(cols, [0, 2500, 5000, 7000]))
    compare("c4", spec, got, exp, bound)


def test_fill_arrow_ipc_and_several_files(tmp_path):
    pa = pytest.importorskip("pyarrow")
    from histbook_b200 import fillable, ingest
    spec = golden_util.load()["c4"]["spec"]
    parts = [_columns(3000, 2), _columns(1500, 3)]
    paths = []
    for i, cols in enumerate(parts):
        path = tmp_path / ("part%d.arrow" % i)
        table = pa.table(cols)
        with pa.OSFile(str(path), "wb") as sink:
            with pa.ipc.new_file(sink, table.schema) as writer:
                for batch in table.to_batches(max_chunksize=2000):
                    writer.write_batch(batch)
        paths.append(str(path))
    with emul_engine.installed():
        sf = fillable.SpecFill(spec)
        nb, nev = ingest.fill_arrow_ipc(sf, paths)
        got = sf.contents()
    assert (nb, nev) == (3, 4500)
    exp, bound = _expect(spec, _slices(parts[0], [0, 2000, 3000]) + [parts[1]])
    compare("c4", spec, got, exp, bound)


def test_fill_npy_memory_maps_and_plain_iterables(tmp_path):
    from histbook_b200 import fillable, ingest
    spec = golden_util.load()["c4"]["spec"]
    cols = _columns(5000, 4)
    files = {}
    for name, arr in cols.items():
        files[name] = str(tmp_path / (name + ".npy"))
        numpy.save(files[name], arr)
    with emul_engine.installed():
        sf = fillable.SpecFill(spec)
        assert ingest.fill_npy(sf, files, batch_rows=2048) == (3, 5000)
        got = sf.contents()
        sf2 = fillable.SpecFill(spec)
        assert ingest.fill_batches(sf2, iter(_slices(cols, [0, 2048, 4096, 5000])), read_ahead=1) == (3, 5000)
        got2 = sf2.contents()
    exp, bound = _expect(spec, _slices(cols, [0, 2048, 4096, 5000]))
    compare("c4", spec, got, exp, bound)
    compare("c4", spec, got2, exp, bound)


def test_errors_travel_from_the_reader_thread_and_missing_fields_are_the_references_error(tmp_path):
    from histbook_b200 import fillable, ingest
    spec = golden_util.load()["c4"]["spec"]

    def broken():
        yield _columns(100, 5)
        raise IOError("disk on fire")
    with emul_engine.installed():
        sf = fillable.SpecFill(spec)
        with pytest.raises(IOError):
            ingest.fill_batches(sf, broken())
        cols = _columns(100, 6)
        del cols["w"]
        with pytest.raises(ValueError) as err:
            ingest.fill_batches(sf, [cols])
        assert "required field 'w' not found in fill arguments" in str(err.value)
    with pytest.raises(TypeError):
        ingest.fill_batches(object(), [])
    lengths = {"x": str(tmp_path / "x.npy"), "y": str(tmp_path / "y.npy")}
    numpy.save(lengths["x"], numpy.zeros(3))
    numpy.save(lengths["y"], numpy.zeros(4))
    with pytest.raises(ValueError):
        list(ingest.npy_batches(lengths, ["x", "y"]))
