"""TEST INFRASTRUCTURE ONLY: a stand-in for histbook_b200.engine that runs the generated kernels on the CPU emulator
(tests/emul_util.py), so that the product's HOST path -- fillable.fill_hists, key discovery, goal tables, slot maps, the
choice of the kernel variant, Store -- can be exercised end to end in the `-m "not gpu"` suite and compared with the oracle.

``installed()`` is a context manager that swaps the handful of engine entry points the fill path uses (compile_source,
Program, Arena, KeySet, DeviceBuffer, fill, init, streams) for emulated ones and restores them afterwards.  "Device
memory" is numpy memory of this process.  Nothing here is reachable from the product: without this context manager the
engine raises EngineError on a machine without a GPU (tests/test_host_logic.py::test_no_cpu_fallback)."""
import contextlib
import ctypes

import numpy

import emul_util
from histbook_b200 import engine, fillable

_ITEMSIZE = [8, 4, 8, 4, 2, 1, 8, 4, 2, 1, 1]
HB_KEY_EMPTY = 0xFFF0DEADBEEFCAFE
VEC_OK = 1
LAUNCHES = []                   # (kernel name, threads, events) of every emulated launch, for the tests to look at


def compile_source(source, lineinfo=True):
    return source if isinstance(source, bytes) else source.encode()


class Program(object):
    def __init__(self, cubin, kernel_name, threads, smem_bytes, tile=None, max_blocks_per_sm=0):
        self.source = cubin.decode()
        self.lib = emul_util.build(self.source)
        self.lib.hb_emu_run.argtypes = [ctypes.POINTER(emul_util.HbParams), ctypes.c_int, ctypes.c_int]
        self.threads, self.smem_bytes, self.tile = int(threads), int(smem_bytes), tile
        self.regs, self.static_smem, self.blocks_per_sm = 0, 0, 1


class Arena(object):
    def __init__(self, ndoubles=None, _array=None):
        self.a = numpy.zeros(int(ndoubles), dtype=numpy.float64) if _array is None else _array

    @property
    def size(self):
        return self.a.size

    @property
    def ptr(self):
        return self.a.ctypes.data

    @property
    def _h(self):
        return self

    def clear(self):
        self.a[:] = 0.0

    def clone(self):
        return Arena(_array=self.a.copy())

    def resize(self, ndoubles):
        grown = numpy.zeros(int(ndoubles), dtype=numpy.float64)
        grown[:min(self.a.size, grown.size)] = self.a[:grown.size]
        self.a = grown

    def read(self, offset=0, count=None):
        count = self.a.size - offset if count is None else count
        return self.a[int(offset):int(offset) + int(count)].copy()

    def write(self, array, offset=0):
        array = numpy.ascontiguousarray(array, dtype=numpy.float64).reshape(-1)
        self.a[int(offset):int(offset) + array.size] = array

    # (numpy restatements of the engine's utility kernels in csrc/hbfill.cu -- enough for the API glue above them to run; the
    #  kernels themselves are covered by the GPU tests only)
    def axpy(self, other, alpha=1.0):
        n = min(self.a.size, other.a.size)
        self.a[:n] = self.a[:n] + float(alpha) * other.a[:n] if alpha != 1.0 else self.a[:n] + other.a[:n]

    def axpy_at(self, dst_off, other, src_off, count, alpha=1.0):
        d, q, c = int(dst_off), int(src_off), int(count)
        self.a[d:d + c] = self.a[d:d + c] + (other.a[q:q + c] if alpha == 1.0 else float(alpha) * other.a[q:q + c])

    def scale(self, alpha):
        self.a[:] = self.a * float(alpha)

    def slice(self, outer, length, inner, start, step, count):
        full = self.a[:int(outer) * int(length) * int(inner)].reshape(int(outer), int(length), int(inner))
        idx = int(start) + int(step) * numpy.arange(int(count))
        return Arena(_array=numpy.ascontiguousarray(full[:, idx, :]).reshape(-1))

    def table(self, nrows, ncol, sumw, sumw2, count, error, effcount, profiles):
        rows = self.a[:int(nrows) * int(ncol)].reshape(int(nrows), int(ncol))
        w = rows[:, sumw]
        w2 = rows[:, sumw2] if sumw2 is not None else None
        good = w > 0.0
        out = []
        with numpy.errstate(all="ignore"):
            eff = (w * w) / w2 if w2 is not None else w
            if count:
                out.append(w)
                if error:
                    out.append(numpy.sqrt(w2 if w2 is not None else w))
            if effcount:
                out.append(numpy.where(good, eff, 0.0))
            for x, x2 in profiles:
                mean = rows[:, x] / w
                out.append(numpy.where(good, mean, 0.0))
                if error:
                    out.append(numpy.where(good, numpy.sqrt((rows[:, x2] / w - mean * mean) / eff), 0.0))
        return Arena(_array=numpy.ascontiguousarray(numpy.stack(out, axis=1)).reshape(-1))

    def project(self, shape, ncol, keep):
        """hb_arena_project: sum over the fixed axes with keep[k] false"""
        full = self.a[:int(numpy.prod(shape)) * int(ncol)].reshape(tuple(int(x) for x in shape) + (int(ncol),))
        drop = tuple(k for k, kept in enumerate(keep) if not kept)
        return Arena(_array=numpy.ascontiguousarray(full.sum(axis=drop)).reshape(-1))


class KeySet(object):
    """hb_template.cuh: [0] capacity, [2] overflow flag, [4...] slots (HB_KEY_EMPTY = free)"""

    def __init__(self, capacity=1 << 12):
        self.capacity = capacity
        self.t = numpy.empty(4 + capacity, dtype=numpy.uint64)
        self.clear()

    @property
    def ptr(self):
        return self.t.ctypes.data

    def clear(self):
        self.t[:4] = (self.capacity, 0, 0, 0)
        self.t[4:] = HB_KEY_EMPTY

    def read(self):
        slots = self.t[4:]
        return slots[slots != numpy.uint64(HB_KEY_EMPTY)].copy(), bool(self.t[2])


class DeviceBuffer(object):
    def __init__(self, array):
        self.a = numpy.ascontiguousarray(array).copy()
        self.nbytes = self.a.nbytes

    @property
    def ptr(self):
        return self.a.ctypes.data

    def read(self, dtype):
        return self.a.view(numpy.uint8)[:self.nbytes].view(dtype).copy()

    def write(self, array):
        array = numpy.ascontiguousarray(array)
        self.a.view(numpy.uint8)[:array.nbytes] = array.view(numpy.uint8).reshape(-1)


def fx_finalize(content, exact, nrows, ncol, ks, masks):
    """csrc/hbfill.cu hb_fx_finalize, with Python integers: per (row, float64 slot) the two 192-bit two's-complement windows
    (upper: lsb 2^s, lower: lsb 2^(s - 51), s = k - 1075) are added exactly, rounded to float64 ONCE (Fraction -> float is
    correctly rounded) and added to the slot's destination columns; the accumulators are zeroed."""
    from fractions import Fraction
    words = exact.a.view(numpy.uint64)
    nslots = len(ks)
    used = words[:int(nrows) * nslots * 6].reshape(-1, 6)
    for i in numpy.flatnonzero(used.any(axis=1)):
        w = [int(x) for x in used[i]]

        def signed(a):
            v = a[0] | (a[1] << 64) | (a[2] << 128)
            return v - (1 << 192) if v >> 191 else v
        total = signed(w[:3]) * (1 << 51) + signed(w[3:])
        row, j = divmod(int(i), nslots)
        v = float(Fraction(total) * Fraction(2) ** (int(ks[j]) - 1075 - 51))
        used[i] = 0
        m, c = int(masks[j]), 0
        while m:
            if m & 1:
                content.a[row * int(ncol) + c] += v
            m >>= 1
            c += 1


def fill(program, columns, n, arenas, aux=(), win=(), timed=False, offset=0, exact=None):
    p = emul_util.HbParams()
    for hid, a in enumerate(exact or ()):
        if a is not None:
            p.fxg[hid] = a.ptr                      # deterministic mode: the histogram's exact accumulators
    for slot, c in enumerate(columns):
        if c[0] == "scalar":
            p.scal[slot] = c[2]
        else:
            assert c[2] == 0, "emulated fills take host columns"
            p.cols[slot] = (c[0] + offset * _ITEMSIZE[c[1]]) if c[0] else 0
    for hid, a in enumerate(arenas):
        p.hist[hid] = a.ptr
    for k, x in enumerate(aux):
        p.aux[k] = int(x) or None
    for k, x in enumerate(win):
        p.win[k] = int(x)
    p.n = int(n)
    p.vec_ok = VEC_OK                           # 1: the four-events-per-thread path of the per-event kernels (their loads have a host form)
    grid = 2 if n > (program.tile or 4 * program.threads) else 1        # (a second block only where it gets a tile: threads are not free here)
    LAUNCHES.append((program.source.count("hb_event<true>") > 0, program.threads, int(n), "group_fast" if program.source.count("gm[") <= 1 and "gd_" in program.source else ""))
    if n > 0:
        assert program.lib.hb_emu_run(ctypes.byref(p), grid, program.threads) == 0
    return {"events": int(n), "h2d_bytes": 0, "launches": 1 if n > 0 else 0, "grid": grid, "blocks_per_sm": 1, "ms_kernel": 0.0}


def fill_multi(parts, columns, n, timed=False, offset=0):
    """hb_fill_multi: several programs over the same columns in one pass -- here one emulated launch per part"""
    out = None
    for part in parts:
        cols = [columns[i] for i in part["colmap"]]
        st = fill(part["program"], cols, n, part["arenas"], aux=part.get("aux", ()), win=part.get("win", ()), timed=timed, offset=offset,
                  exact=part.get("exact"))
        if out is None:
            out = st
        else:
            out["launches"] += st["launches"]
    return out or {"events": int(n), "h2d_bytes": 0, "launches": 0, "grid": 0, "blocks_per_sm": 0, "ms_kernel": 0.0}


@contextlib.contextmanager
def installed():
    names = ["compile_source", "Program", "Arena", "KeySet", "DeviceBuffer", "fill", "fill_multi", "fx_finalize", "init", "producer_stream", "stream_wait", "sync"]
    saved = dict((n, getattr(engine, n)) for n in names)
    engine.compile_source = compile_source
    engine.Program = Program
    engine.Arena = Arena
    engine.KeySet = KeySet
    engine.DeviceBuffer = DeviceBuffer
    engine.fill = fill
    engine.fx_finalize = fx_finalize
    engine.fill_multi = fill_multi
    engine.init = lambda device=None: 0
    engine.producer_stream = lambda: 0
    engine.stream_wait = lambda waiter, signaler: None
    engine.sync = lambda: None
    fillable.Plan.generation += 1               # no kernel compiled for the real engine survives into the emulated one
    del LAUNCHES[:]
    try:
        yield
    finally:
        for n, f in saved.items():
            setattr(engine, n, f)
        fillable.Plan.generation += 1
