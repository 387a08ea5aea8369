"""ctypes binding of libhbfill.so (include/hbfill.h) -- the only door to the device.

There is deliberately no fallback: if the shared library is missing, or no CUDA device is
present, every compute entry point raises ``EngineError``.
"""
import ctypes
import hashlib
import os
import sys
import threading

import numpy

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhbfill.so")
TEMPLATE_PATH = os.path.join(_HERE, "csrc", "hb_template.cuh")

HB_FILL_TIMED = 1
MAX_COLS, MAX_HISTS, MAX_AUX = 32, 128, 64


class EngineError(RuntimeError):
    pass


class hb_column(ctypes.Structure):
    _fields_ = [("ptr", ctypes.c_void_p), ("dtype", ctypes.c_int32), ("location", ctypes.c_int32),
                ("is_scalar", ctypes.c_int32), ("reserved", ctypes.c_int32), ("scalar_bits", ctypes.c_uint64)]


class hb_launch(ctypes.Structure):
    _fields_ = [("prog", ctypes.c_void_p), ("ncols", ctypes.c_int32), ("colmap", ctypes.POINTER(ctypes.c_int32)),
                ("nhists", ctypes.c_int32), ("hists", ctypes.POINTER(ctypes.c_void_p)), ("exact", ctypes.POINTER(ctypes.c_void_p)),
                ("naux", ctypes.c_int32), ("aux", ctypes.POINTER(ctypes.c_void_p)), ("nwin", ctypes.c_int32),
                ("win", ctypes.POINTER(ctypes.c_int32))]


class hb_stats(ctypes.Structure):
    _fields_ = [("events", ctypes.c_int64), ("h2d_bytes", ctypes.c_int64), ("launches", ctypes.c_int32),
                ("grid", ctypes.c_int32), ("blocks_per_sm", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("ms_kernel", ctypes.c_double)]


# every symbol include/hbfill.h declares: name -> (restype, argtypes)
_P = ctypes.c_void_p
_PP = ctypes.POINTER(ctypes.c_void_p)
_I = ctypes.c_int
_I64 = ctypes.c_int64
_IP = ctypes.POINTER(ctypes.c_int)
SYMBOLS = {
    "hb_last_error": (ctypes.c_char_p, []),
    "hb_version": (_I, []),
    "hb_nvrtc_version": (_I, []),
    "hb_init": (_I, [_I]),
    "hb_device_info": (_I, [_IP, _IP, ctypes.POINTER(_I64), ctypes.POINTER(_I64)]),
    "hb_set_stream": (_I, [_P]),
    "hb_stream_wait": (_I, [_P, _P]),
    "hb_sync": (_I, []),
    "hb_set_chunk_events": (_I, [_I64]),
    "hb_program_compile": (_I, [ctypes.c_char_p, ctypes.c_char_p, _I, _PP, ctypes.POINTER(ctypes.c_size_t)]),
    "hb_program_load": (_I, [_P, ctypes.c_size_t, ctypes.c_char_p, _I, _I, _PP]),
    "hb_program_destroy": (_I, [_P]),
    "hb_program_set_tile": (_I, [_P, _I]),
    "hb_program_set_max_blocks_per_sm": (_I, [_P, _I]),
    "hb_fill_multi": (_I, [_I, ctypes.POINTER(hb_launch), _I, ctypes.POINTER(hb_column), _I64, _I, ctypes.POINTER(hb_stats)]),
    "hb_set_copy_threads": (_I, [_I]),
    "hb_program_info": (_I, [_P, _IP, _IP, _IP]),
    "hb_free": (None, [_P]),
    "hb_arena_create": (_I, [_I64, _PP]),
    "hb_arena_destroy": (_I, [_P]),
    "hb_arena_clear": (_I, [_P]),
    "hb_arena_clone": (_I, [_P, _PP]),
    "hb_arena_resize": (_I, [_P, _I64]),
    "hb_arena_read": (_I, [_P, _I64, _I64, _P]),
    "hb_arena_write": (_I, [_P, _I64, _I64, _P]),
    "hb_arena_axpy": (_I, [_P, _P, ctypes.c_double]),
    "hb_arena_axpy_at": (_I, [_P, _I64, _P, _I64, _I64, ctypes.c_double]),
    "hb_arena_scale": (_I, [_P, ctypes.c_double]),
    "hb_arena_snapshot": (_I, [_I, _PP, ctypes.POINTER(_I64), ctypes.POINTER(_I64), ctypes.POINTER(_I64), _P]),
    "hb_arena_project": (_I, [_P, _I, ctypes.POINTER(_I64), _I, ctypes.POINTER(ctypes.c_int32), _PP]),
    "hb_arena_slice": (_I, [_P, _I64, _I64, _I64, _I64, _I64, _I64, _PP]),
    "hb_arena_table": (_I, [_P, _I64, _I, _I, _I, _I, _I, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32), _PP]),
    "hb_arena_ptr": (_P, [_P]),
    "hb_arena_size": (_I64, [_P]),
    "hb_fill": (_I, [_P, _I, ctypes.POINTER(hb_column), _I64, _I, _PP, _I, _PP, _I, ctypes.POINTER(ctypes.c_int32), _I, ctypes.POINTER(hb_stats)]),
    "hb_fill_exact": (_I, [_P, _I, ctypes.POINTER(hb_column), _I64, _I, _PP, _PP, _I, _PP, _I, ctypes.POINTER(ctypes.c_int32), _I, ctypes.POINTER(hb_stats)]),
    "hb_fx_finalize": (_I, [_P, _P, _I64, _I, _I, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_uint64)]),
    "hb_keyset_create": (_I, [_I, _PP]),
    "hb_keyset_destroy": (_I, [_P]),
    "hb_keyset_clear": (_I, [_P]),
    "hb_keyset_ptr": (_P, [_P]),
    "hb_keyset_read": (_I, [_P, _P, _I, _IP, _IP]),
    "hb_flush_l2": (_I, []),
    "hb_memcpy_h2d": (_I, [_P, _P, ctypes.c_size_t]),
    "hb_memcpy_d2h": (_I, [_P, _P, ctypes.c_size_t]),
    "hb_device_malloc": (_I, [ctypes.c_size_t, _PP]),
    "hb_device_free": (_I, [_P]),
}

_lib = None
_lock = threading.Lock()
_initialized_device = None


def lib():
    """The loaded shared library (raises EngineError if it has not been built)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.isfile(LIB_PATH):
                    raise EngineError("{0} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                      "(histbook_b200 has no CPU fallback)".format(LIB_PATH))
                handle = ctypes.CDLL(LIB_PATH)
                for name, (restype, argtypes) in SYMBOLS.items():
                    fn = getattr(handle, name)
                    fn.restype = restype
                    fn.argtypes = argtypes
                _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().hb_last_error()
        raise EngineError("libhbfill error {0}: {1}".format(rc, msg.decode(errors="replace") if msg else "?"))


def init(device=None):
    """Select the CUDA device (LOCAL_RANK by default) -- required before any device call."""
    global _initialized_device
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0")) if _initialized_device is None else _initialized_device
    if _initialized_device != device:
        check(lib().hb_init(int(device)))
        _initialized_device = device
    return device


def device_info():
    init()
    sm, smem, free, total = ctypes.c_int(), ctypes.c_int(), ctypes.c_int64(), ctypes.c_int64()
    check(lib().hb_device_info(ctypes.byref(sm), ctypes.byref(smem), ctypes.byref(free), ctypes.byref(total)))
    return {"sm_count": sm.value, "max_smem_optin": smem.value, "free_bytes": free.value, "total_bytes": total.value}


def sync():
    check(lib().hb_sync())


def flush_l2():
    init()
    check(lib().hb_flush_l2())


# ----------------------------------------------------------------------------- compile (NVRTC, no GPU needed)

_template_text = None
_cubin_cache = {}


def template_text():
    global _template_text
    if _template_text is None:
        with open(TEMPLATE_PATH, "rb") as f:
            _template_text = f.read()
    return _template_text


def cache_dir():
    d = os.environ.get("HB_CACHE_DIR", os.path.join(os.path.expanduser("~"), ".cache", "histbook_b200"))
    try:
        os.makedirs(d, exist_ok=True)
    except OSError:
        return None
    return d


def compile_source(source, lineinfo=True):
    """CUDA source text -> sm_100a cubin bytes.  Cached in memory and on disk, keyed by
    sha256(template + source)."""
    if isinstance(source, str):
        source = source.encode()
    key = hashlib.sha256(template_text() + b"\0" + source + (b"L" if lineinfo else b"")).hexdigest()
    if key in _cubin_cache:
        return _cubin_cache[key]
    d = cache_dir()
    path = os.path.join(d, key + ".cubin") if d else None
    if path and os.path.isfile(path):
        with open(path, "rb") as f:
            cubin = f.read()
        _cubin_cache[key] = cubin
        return cubin
    out, size = ctypes.c_void_p(), ctypes.c_size_t()
    check(lib().hb_program_compile(source, template_text(), 1 if lineinfo else 0, ctypes.byref(out), ctypes.byref(size)))
    try:
        cubin = ctypes.string_at(out.value, size.value)
    finally:
        lib().hb_free(out)
    _cubin_cache[key] = cubin
    if path:
        try:
            tmp = path + ".%d.tmp" % os.getpid()
            with open(tmp, "wb") as f:
                f.write(cubin)
            os.replace(tmp, path)
        except OSError:
            pass
    return cubin


# ----------------------------------------------------------------------------- handles

class Program(object):
    """A loaded fill kernel."""

    def __init__(self, cubin, kernel_name, threads, smem_bytes, tile=None, max_blocks_per_sm=0):
        init()
        self._h = ctypes.c_void_p()
        self._cubin = cubin
        check(lib().hb_program_load(cubin, len(cubin), kernel_name.encode(), int(threads), int(smem_bytes), ctypes.byref(self._h)))
        if tile is not None and int(tile) != 4 * int(threads):
            check(lib().hb_program_set_tile(self._h, int(tile)))
        if max_blocks_per_sm:
            check(lib().hb_program_set_max_blocks_per_sm(self._h, int(max_blocks_per_sm)))
        regs, ssm, occ = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        check(lib().hb_program_info(self._h, ctypes.byref(regs), ctypes.byref(ssm), ctypes.byref(occ)))
        self.regs, self.static_smem, self.blocks_per_sm = regs.value, ssm.value, occ.value
        self.threads, self.smem_bytes = threads, smem_bytes

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.hb_program_destroy(h)


class Arena(object):
    """float64 bin contents on the device (one per histogram)."""

    def __init__(self, ndoubles=None, _handle=None):
        init()
        if _handle is not None:
            self._h = _handle
        else:
            self._h = ctypes.c_void_p()
            check(lib().hb_arena_create(int(ndoubles), ctypes.byref(self._h)))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.hb_arena_destroy(h)

    @property
    def size(self):
        return lib().hb_arena_size(self._h)

    @property
    def ptr(self):
        return lib().hb_arena_ptr(self._h)

    def clear(self):
        check(lib().hb_arena_clear(self._h))

    def clone(self):
        out = ctypes.c_void_p()
        check(lib().hb_arena_clone(self._h, ctypes.byref(out)))
        return Arena(_handle=out)

    def resize(self, ndoubles):
        check(lib().hb_arena_resize(self._h, int(ndoubles)))

    def read(self, offset=0, count=None):
        if count is None:
            count = self.size - offset
        out = numpy.empty(int(count), dtype=numpy.float64)
        if count:
            check(lib().hb_arena_read(self._h, int(offset), int(count), out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def write(self, array, offset=0):
        array = numpy.ascontiguousarray(array, dtype=numpy.float64).reshape(-1)
        if array.size:
            check(lib().hb_arena_write(self._h, int(offset), int(array.size), array.ctypes.data_as(ctypes.c_void_p)))

    def axpy(self, other, alpha=1.0):
        check(lib().hb_arena_axpy(self._h, other._h, float(alpha)))

    def axpy_at(self, dst_off, other, src_off, count, alpha=1.0):
        check(lib().hb_arena_axpy_at(self._h, int(dst_off), other._h, int(src_off), int(count), float(alpha)))

    def scale(self, alpha):
        check(lib().hb_arena_scale(self._h, float(alpha)))

    def project(self, shape, ncol, keep):
        """Sum over the fixed axes with keep[k] false (Hist.project on the device); returns a new Arena."""
        shape = [int(x) for x in shape]
        arr = (ctypes.c_int64 * len(shape))(*shape)
        karr = (ctypes.c_int32 * len(shape))(*[1 if k else 0 for k in keep])
        out = ctypes.c_void_p()
        check(lib().hb_arena_project(self._h, len(shape), arr, int(ncol), karr, ctypes.byref(out)))
        return Arena(_handle=out)

    def slice(self, outer, length, inner, start, step, count):
        """[outer][length][inner] -> [outer][count][inner], element k of the middle axis from start + k * step (Hist.select on
        the device); returns a new Arena."""
        out = ctypes.c_void_p()
        check(lib().hb_arena_slice(self._h, int(outer), int(length), int(inner), int(start), int(step), int(count), ctypes.byref(out)))
        return Arena(_handle=out)

    def table(self, nrows, ncol, sumw, sumw2, count, error, effcount, profiles):
        """Hist.table's columns per bin (histbook/proj.py:577-627) computed on the device; ``profiles`` = [(sumwx column, sumwx2
        column)]; sumw2 = None for an unweighted histogram.  Returns a new Arena of nrows x ncolumns doubles."""
        xs = (ctypes.c_int32 * max(len(profiles), 1))(*[int(a) for a, b in profiles])
        x2s = (ctypes.c_int32 * max(len(profiles), 1))(*[int(b) for a, b in profiles])
        flags = (1 if count else 0) | (2 if error else 0) | (4 if effcount else 0)
        out = ctypes.c_void_p()
        check(lib().hb_arena_table(self._h, int(nrows), int(ncol), int(sumw), -1 if sumw2 is None else int(sumw2), flags, len(profiles), xs, x2s,
                                   ctypes.byref(out)))
        return Arena(_handle=out)

    @property
    def __cuda_array_interface__(self):
        return {"shape": (int(self.size),), "typestr": "<f8", "data": (int(self.ptr or 0), False), "version": 3, "strides": None}

    def as_tensor(self):
        """Zero-copy torch view (for torch.distributed collectives over NCCL); cached until the arena moves or grows."""
        import torch
        ptr, size = self.ptr, self.size
        cached = getattr(self, "_view", None)
        if cached is None or cached[0] != ptr or cached[1] != size:
            cached = self._view = (ptr, size, torch.as_tensor(self, device="cuda:%d" % init()))
        return cached[2]


class KeySet(object):
    def __init__(self, capacity=1 << 12):
        init()
        self.capacity = capacity
        self._h = ctypes.c_void_p()
        check(lib().hb_keyset_create(int(capacity), ctypes.byref(self._h)))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.hb_keyset_destroy(h)

    @property
    def ptr(self):
        return lib().hb_keyset_ptr(self._h)

    def clear(self):
        check(lib().hb_keyset_clear(self._h))

    def read(self):
        keys = numpy.empty(self.capacity, dtype=numpy.uint64)
        n, over = ctypes.c_int(), ctypes.c_int()
        check(lib().hb_keyset_read(self._h, keys.ctypes.data_as(ctypes.c_void_p), self.capacity, ctypes.byref(n), ctypes.byref(over)))
        return keys[:n.value].copy(), bool(over.value)


class DeviceBuffer(object):
    """Small raw device allocation (group-key lookup tables)."""

    def __init__(self, array):
        init()
        array = numpy.ascontiguousarray(array)
        self.nbytes = array.nbytes
        self._p = ctypes.c_void_p()
        check(lib().hb_device_malloc(max(array.nbytes, 8), ctypes.byref(self._p)))
        if array.nbytes:
            check(lib().hb_memcpy_h2d(self._p, array.ctypes.data_as(ctypes.c_void_p), array.nbytes))

    @property
    def ptr(self):
        return self._p.value

    def read(self, dtype):
        """The buffer's current device contents as a numpy array of ``dtype``."""
        out = numpy.empty(self.nbytes // numpy.dtype(dtype).itemsize, dtype=dtype)
        if out.nbytes:
            check(lib().hb_memcpy_d2h(out.ctypes.data_as(ctypes.c_void_p), self._p, out.nbytes))
        return out

    def write(self, array):
        array = numpy.ascontiguousarray(array)
        assert array.nbytes <= max(self.nbytes, 8)
        if array.nbytes:
            check(lib().hb_memcpy_h2d(self._p, array.ctypes.data_as(ctypes.c_void_p), array.nbytes))

    def __del__(self):
        p, self._p = getattr(self, "_p", None), None
        if p and _lib is not None:
            _lib.hb_device_free(p)


# ----------------------------------------------------------------------------- streams

def producer_stream():
    """The CUDA stream device columns handed to ``fill`` are being produced on, if it is not the engine's own (stream 0):
    torch's current stream when torch has a CUDA context.  Returns 0 otherwise."""
    torch = sys.modules.get("torch")
    if torch is None or not torch.cuda.is_initialized():
        return 0
    raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)       # the handle without building a Stream object (~0.3 us)
    if raw is not None:
        return int(raw(init()) or 0)
    return int(torch.cuda.current_stream().cuda_stream or 0)


def stream_wait(waiter, signaler):
    """hb_stream_wait: 0 stands for the engine's compute stream."""
    check(lib().hb_stream_wait(ctypes.c_void_p(waiter or None), ctypes.c_void_p(signaler or None)))


# ----------------------------------------------------------------------------- fill

_ITEMSIZE = [8, 4, 8, 4, 2, 1, 8, 4, 2, 1, 1]


def fx_finalize(content, exact, nrows, ncol, ks, masks):
    """Round the exact accumulators of one histogram into its float64 content and zero them (deterministic mode)."""
    n = len(ks)
    karr = (ctypes.c_int32 * n)(*[int(x) for x in ks])
    marr = (ctypes.c_uint64 * n)(*[int(x) for x in masks])
    check(lib().hb_fx_finalize(content._h, exact._h, int(nrows), int(ncol), n, karr, marr))


def fill(program, columns, n, arenas, aux=(), win=(), timed=False, offset=0, exact=None):
    """columns: list of (ptr:int, dtype_code, location 0 host / 1 device) or ("scalar", dtype_code, bits).
    ``offset``/``n`` select the event range [offset, offset + n) of every column.  ``exact``: per histogram an
    Arena of 128-bit accumulators or None (deterministic mode)."""
    init()
    ncols = len(columns)
    structs = []
    for c in columns:
        if c[0] == "scalar":
            structs.append(hb_column(None, c[1], 0, 1, 0, c[2]))
        else:
            structs.append(hb_column((c[0] + offset * _ITEMSIZE[c[1]]) if c[0] else c[0], c[1], c[2], 0, 0, 0))
    cols = (hb_column * max(ncols, 1))(*structs)
    hist = (ctypes.c_void_p * max(len(arenas), 1))(*[a._h for a in arenas])
    auxp = (ctypes.c_void_p * max(len(aux), 1))(*[int(x) for x in aux])
    winp = (ctypes.c_int32 * max(len(win), 1))(*[int(x) for x in win])
    stats = hb_stats()
    if exact is None:
        check(lib().hb_fill(program._h, ncols, cols, int(n), len(arenas), hist, len(aux), auxp, len(win), winp,
                            HB_FILL_TIMED if timed else 0, ctypes.byref(stats)))
    else:
        ex = (ctypes.c_void_p * max(len(arenas), 1))(*[(a._h if a is not None else None) for a in exact])
        check(lib().hb_fill_exact(program._h, ncols, cols, int(n), len(arenas), hist, ex, len(aux), auxp, len(win), winp,
                                  HB_FILL_TIMED if timed else 0, ctypes.byref(stats)))
    return {"events": stats.events, "h2d_bytes": stats.h2d_bytes, "launches": stats.launches, "grid": stats.grid,
            "blocks_per_sm": stats.blocks_per_sm, "ms_kernel": stats.ms_kernel}


def fill_multi(parts, columns, n, timed=False, offset=0):
    """One pass over ``columns`` for several programs at once (hb_fill_multi): ``parts`` = list of dicts with keys
    program, colmap (indices into ``columns`` per column slot of the program), arenas, aux, win, exact (or None).
    Every host chunk is staged once; the parts' kernels run side by side on their own streams."""
    init()
    ncols = len(columns)
    structs = []
    for c in columns:
        if c[0] == "scalar":
            structs.append(hb_column(None, c[1], 0, 1, 0, c[2]))
        else:
            structs.append(hb_column((c[0] + offset * _ITEMSIZE[c[1]]) if c[0] else c[0], c[1], c[2], 0, 0, 0))
    cols = (hb_column * max(ncols, 1))(*structs)
    launches = (hb_launch * len(parts))()
    keep = []
    for q, part in zip(launches, parts):
        arenas, aux, win, exact = part["arenas"], part.get("aux", ()), part.get("win", ()), part.get("exact")
        colmap = (ctypes.c_int32 * max(len(part["colmap"]), 1))(*[int(i) for i in part["colmap"]])
        hist = (ctypes.c_void_p * max(len(arenas), 1))(*[a._h for a in arenas])
        auxp = (ctypes.c_void_p * max(len(aux), 1))(*[int(x) for x in aux])
        winp = (ctypes.c_int32 * max(len(win), 1))(*[int(x) for x in win])
        ex = None
        if exact is not None:
            ex = (ctypes.c_void_p * max(len(arenas), 1))(*[(a._h if a is not None else None) for a in exact])
        keep.append((colmap, hist, auxp, winp, ex))
        q.prog = part["program"]._h
        q.ncols, q.colmap = len(part["colmap"]), colmap
        q.nhists, q.hists = len(arenas), hist
        q.exact = ex if ex is not None else ctypes.POINTER(ctypes.c_void_p)()
        q.naux, q.aux = len(aux), auxp
        q.nwin, q.win = len(win), winp
    stats = hb_stats()
    check(lib().hb_fill_multi(len(parts), launches, ncols, cols, int(n), HB_FILL_TIMED if timed else 0, ctypes.byref(stats)))
    return {"events": stats.events, "h2d_bytes": stats.h2d_bytes, "launches": stats.launches, "grid": stats.grid,
            "blocks_per_sm": stats.blocks_per_sm, "ms_kernel": stats.ms_kernel}


class Snapshot(object):
    """A prepared hb_arena_snapshot call: [(arena, source offset, destination offset, count)] -> one launch."""

    def __init__(self, segments):
        n = len(segments)
        self.n = n
        self.keep = [a for a, so, do, c in segments]
        self.src = (ctypes.c_void_p * max(n, 1))(*[a._h for a, so, do, c in segments])
        self.src_off = (ctypes.c_int64 * max(n, 1))(*[int(so) for a, so, do, c in segments])
        self.dst_off = (ctypes.c_int64 * max(n, 1))(*[int(do) for a, so, do, c in segments])
        self.count = (ctypes.c_int64 * max(n, 1))(*[int(c) for a, so, do, c in segments])

    def run(self, dst_ptr):
        if self.n:
            check(lib().hb_arena_snapshot(self.n, self.src, self.src_off, self.dst_off, self.count, ctypes.c_void_p(int(dst_ptr))))
