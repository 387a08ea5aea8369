"""File-backed and batched columnar ingest: fill a ``Hist`` / ``Book`` from columns that do not sit
in memory as a whole -- Parquet files, Arrow IPC / Feather files, ``.npy`` files, or any iterable of column mappings.

The reference fills from anything with ``__getitem__`` (histbook/util/__init__.py:38-47, fill.py:124-146) and documents chunked
refill as the way to process more events than fit in memory (README.rst:29); a histogram's contents are additive over fills
(hist.py:429-456), so a stream of batches is a sequence of ordinary ``fill`` calls.  What this module adds is the plumbing
around them:

  * only the fields the histograms need are read (``target.fields``, histbook/fill.py:41-59);
  * Arrow batches are handed over without a copy where their layout allows it (columns.bind uses the values buffer in place);
  * a read-ahead thread decodes batch k + 1 while batch k is staged and filled (``hb_fill_multi`` overlaps the host copy, the
    DMA and the kernels of ONE fill; across fills the file read is what would otherwise serialise).

Every batch goes through the target's own ``fill`` -- the device path, no CPU fallback.
"""
import queue
import threading

import numpy


def _fields_of(target):
    fields = getattr(target, "fields", None)
    if fields is None:
        raise TypeError("{0!r} has no `fields`: not a histbook Hist / Book (or SpecFill)".format(type(target).__name__))
    return list(fields)


def _read_ahead(batches, depth):
    """iterate ``batches`` on a worker thread, ``depth`` batches ahead of the consumer"""
    if depth <= 0:
        for b in batches:
            yield b
        return
    q = queue.Queue(maxsize=depth)
    done = object()
    stop = threading.Event()

    def work():
        try:
            for b in batches:
                while not stop.is_set():
                    try:
                        q.put(b, timeout=0.1)
                        break
                    except queue.Full:
                        continue
                if stop.is_set():
                    return
            q.put(done)
        except BaseException as err:           # re-raised in the consumer
            q.put(err)

    t = threading.Thread(target=work, name="hb-ingest-read-ahead", daemon=True)
    t.start()
    try:
        while True:
            item = q.get()
            if item is done:
                return
            if isinstance(item, BaseException):
                raise item
            yield item
    finally:
        stop.set()


def fill_batches(target, batches, read_ahead=2):
    """Fill ``target`` (Hist, Book, or anything with ``fill`` and ``fields``) from an iterable of column mappings, one ``fill``
    per batch; returns (batches, events).  ``read_ahead`` batches are produced on a worker thread while the current one fills."""
    fields = _fields_of(target)
    nb = nev = 0
    for batch in _read_ahead(iter(batches), read_ahead):
        cols = dict((name, batch[name]) for name in fields if _has(batch, name))
        target.fill(cols)
        nb += 1
        nev += _length(cols)
    return nb, nev


def _has(batch, name):
    try:
        batch[name]
        return True
    except (KeyError, IndexError):
        return False               # a missing field: the fill raises the reference's ValueError (or takes pi / e / inf / nan)


def _length(cols):
    for v in cols.values():
        try:
            if numpy.ndim(v) == 1 or hasattr(v, "__len__"):
                return len(v)
        except TypeError:
            continue
    return 0


def _arrow_columns(batch, fields):
    """one Arrow RecordBatch / Table -> {field: Arrow array} for the fields present (columns.bind takes them zero-copy)"""
    names = set(batch.schema.names)
    out = {}
    for name in fields:
        if name in names:
            col = batch.column(name)
            if hasattr(col, "combine_chunks") and getattr(col, "num_chunks", 1) != 1:
                col = col.combine_chunks()
            out[name] = col
    return out


def parquet_batches(paths, fields, batch_rows=1 << 22):
    """Arrow record batches of at most ``batch_rows`` rows from one or several Parquet files, only the columns in ``fields``"""
    import pyarrow.parquet as pq
    if isinstance(paths, (str, bytes)) or hasattr(paths, "__fspath__"):
        paths = [paths]
    for path in paths:
        f = pq.ParquetFile(path)
        present = [name for name in fields if name in f.schema_arrow.names]
        for batch in f.iter_batches(batch_size=int(batch_rows), columns=present):
            yield _arrow_columns(batch, fields)


def fill_parquet(target, paths, batch_rows=1 << 22, read_ahead=2):
    """Fill from Parquet files in batches of ``batch_rows`` rows; returns (batches, events)."""
    return fill_batches(target, parquet_batches(paths, _fields_of(target), batch_rows), read_ahead)


def arrow_ipc_batches(paths, fields):
    """the record batches of Arrow IPC files (Feather v2), memory-mapped: a batch's buffers are pages of the file"""
    import pyarrow as pa
    import pyarrow.ipc
    if isinstance(paths, (str, bytes)) or hasattr(paths, "__fspath__"):
        paths = [paths]
    for path in paths:
        with pa.memory_map(str(path), "r") as source:
            reader = pa.ipc.open_file(source)
            for i in range(reader.num_record_batches):
                yield _arrow_columns(reader.get_batch(i), fields)


def fill_arrow_ipc(target, paths, read_ahead=2):
    """Fill from Arrow IPC (Feather v2) files, one ``fill`` per record batch; returns (batches, events)."""
    return fill_batches(target, arrow_ipc_batches(paths, _fields_of(target)), read_ahead)


def npy_batches(files, fields, batch_rows=1 << 24):
    """slices of ``batch_rows`` events of memory-mapped ``.npy`` files, ``files`` = {field: path}; all files one length"""
    maps = dict((name, numpy.load(files[name], mmap_mode="r")) for name in fields if name in files)
    lengths = set(len(m) for m in maps.values())
    if len(lengths) > 1:
        raise ValueError("the .npy files have different lengths: {0}".format(dict((k, len(v)) for k, v in maps.items())))
    n = lengths.pop() if lengths else 0
    for start in range(0, n, int(batch_rows)):
        yield dict((name, m[start:start + int(batch_rows)]) for name, m in maps.items())


def fill_npy(target, files, batch_rows=1 << 24, read_ahead=0):
    """Fill from memory-mapped ``.npy`` files ({field: path}) in slices of ``batch_rows`` events; returns (batches, events).
    (No read-ahead by default: a slice of a memory map is read by the staging copy itself.)"""
    return fill_batches(target, npy_batches(files, _fields_of(target), batch_rows), read_ahead)
