"""TEST INFRASTRUCTURE ONLY: runs a *generated* fill kernel on CPU threads (tests/emul/cuda_emul.h) and returns the
histograms it filled, so that the code generator's logic can be checked against the oracle without a GPU.

What is emulated: the kernel text codegen.generate() produces + csrc/hb_template.cuh, compiled with g++ (-ffp-contract=off:
no FMA, like NVRTC's --fmad=false) behind a shim that maps the CUDA execution model onto std::threads.  The host side that
the product runs around a launch (key discovery on the device, slot maps, arenas: fillable.py / store.py / hbfill.cu) is
replaced by a few lines of numpy here -- the keys come from the oracle -- so this covers the kernels, not the engine.
Never used by the product path; the GPU tests remain the parity tests proper.
"""
import ctypes
import hashlib
import os
import re
import subprocess
import tempfile

import numpy

from histbook_b200 import codegen, fillable, spec as hbspec

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
TEMPLATE = os.path.join(REPO, "histbook_b200", "csrc", "hb_template.cuh")
CACHE = os.path.join(tempfile.gettempdir(), "hb_emul_cache")

HB_KEY_NAN = 0x7FF8000000000000


class HbParams(ctypes.Structure):
    # csrc/hb_template.cuh struct HbParams
    _fields_ = [("cols", ctypes.c_void_p * 32), ("scal", ctypes.c_uint64 * 32), ("hist", ctypes.c_void_p * 128),
                ("fxg", ctypes.c_void_p * 128), ("aux", ctypes.c_void_p * 64), ("n", ctypes.c_int64), ("vec_ok", ctypes.c_int),
                ("flags", ctypes.c_int), ("win", ctypes.c_int * 128), ("sched", ctypes.c_void_p)]


# (exact text in csrc/hb_template.cuh, host form)
PTX_FORMS = [
    ("    hb_loader<T>::load4(reinterpret_cast<const T*>(base) + tile_first, tid, nthreads, o);",
     "    for (int k = 0; k < 4; ++k) o[k] = reinterpret_cast<const T*>(base)[tile_first + (hb_i64)(k >> 1) * 2 * nthreads + (hb_i64)tid * 2 + (k & 1)];"),
    ('    asm("bfe.u32 %0, %1, 20, 11;" : "=r"(e) : "r"(hi));', "    e = hb_emu_bfe_20_11(hi);"),
    ('    if (neg) asm("sub.cc.u32 %0, 0, %0;\\n\\tsubc.cc.u32 %1, 0, %1;\\n\\tsubc.u32 %2, 0, %2;" : "+r"(q0), "+r"(q1), "+r"(q2));',
     "    if (neg) hb_emu_neg96(q0, q1, q2);"),
    ('    asm("{\\n\\t.reg .u32 t;\\n\\tadd.cc.u32 t, %1, %2;\\n\\taddc.u32 %0, 0, 0;\\n\\t}" : "=r"(c) : "r"(o), "r"(f.x0));',
     "    c = hb_emu_carry(o, f.x0);"),
    ('    asm("add.cc.u32 %0, %3, %5;\\n\\taddc.cc.u32 %1, %4, 0;\\n\\taddc.u32 %2, 0, 0;"\n        : "=r"(a1), "=r"(a2), "=r"(pre) : "r"(f.x1), "r"(f.x2), "r"(carry));',
     "    hb_emu_stage1a(a1, a2, pre, f.x1, f.x2, carry);"),
    ('    asm("{\\n\\t.reg .u32 t;\\n\\tadd.cc.u32 t, %2, %3;\\n\\taddc.cc.u32 %0, %0, 0;\\n\\taddc.u32 %1, %1, 0;\\n\\t}"\n        : "+r"(a2), "+r"(pre) : "r"(o), "r"(a1));',
     "    hb_emu_stage1b(a2, pre, o, a1);"),
    ('    asm("{\\n\\t.reg .u32 u;\\n\\tadd.cc.u32 u, %1, %2;\\n\\taddc.u32 %0, %3, %4;\\n\\t}" : "=r"(top) : "r"(o), "r"(a2), "r"((hb_u32)(t >> 32)), "r"(f.ext));',
     "    top = hb_emu_stage2(o, a2, (hb_u32)(t >> 32), f.ext);"),
]

RUNNER = r"""
extern "C" int hb_emu_run(const HbParams* p, int grid, int threads)
{
    blockDim.x = (unsigned)threads;
    gridDim.x = (unsigned)grid;
    hb_emu_failed.store(0);
    for (int b = 0; b < grid && !hb_emu_failed.load(); ++b) {
        std::vector<std::thread> pool;
        pool.reserve(threads);
        for (int t = 0; t < threads; ++t)
            pool.emplace_back([=] { threadIdx.x = (unsigned)t; blockIdx.x = (unsigned)b; hb_fill_kernel(*p); });
        for (auto& th : pool) th.join();
    }
    return hb_emu_failed.load();           // 1: a barrier timed out (cuda_emul.h)
}
"""


def host_source(kernel_source):
    """The generated CUDA text as one host C++ translation unit."""
    with open(TEMPLATE) as f:
        template = f.read()
    template = template.replace("#pragma once", "")
    # inline PTX -> host forms (tests/emul/cuda_emul.h).  Every substitution must hit exactly once: a change of the template
    # that the emulator does not know about fails here, not silently in a test.
    template = template.replace("void hb_red_f64(double* addr, double v) {", "void hb_red_f64_ptx(double* addr, double v) {")
    for ptx, host in PTX_FORMS:
        assert template.count(ptx) == 1, "hb_template.cuh changed: no unique match for %r" % ptx[:60]
        template = template.replace(ptx, host)
    # what is left: the vector loads inside hb_loader (hb_load4 is redirected above them), the L2 prefetch hint and the
    # cache-hinted RED of hb_red_f64_ptx (not called)
    template = re.sub(r"asm\s+volatile\s*\(", "HB_EMU_NOASM(", template)
    template = template.replace('asm("createpolicy', 'HB_EMU_NOASM("createpolicy')
    assert not re.search(r"\basm\s*\(", template), "an asm statement of hb_template.cuh has no host form"
    src = kernel_source.replace('#include "hb_template.cuh"', '#include "cuda_emul.h"\n' + template)
    src, nbar = re.subn(r'asm volatile\("bar\.sync %0, %1;" :: "r"\(([^)]*)\), "r"\(([^)]*)\) : "memory"\);', r"hb_emu_bar(\1, \2);", src)
    assert "asm volatile" not in src.split("hb_fill_kernel")[-1], "an inline-PTX statement of the generated kernel has no host form"
    src = src.replace("extern __shared__ __align__(16) unsigned char hb_smem[];", "")
    return src + RUNNER


CXXFLAGS = ["-std=c++17", "-O1", "-ffp-contract=off", "-fno-strict-aliasing", "-fwrapv", "-w", "-pthread", "-fPIC"]


def _shim_dir(shim):
    """a directory holding the shim and its precompiled header (the standard headers it pulls in are 40 % of a kernel's compile time)"""
    d = os.path.join(CACHE, "shim_" + hashlib.sha256((" ".join(CXXFLAGS) + shim).encode()).hexdigest()[:16])
    if not os.path.isfile(os.path.join(d, "cuda_emul.h.gch")):
        tmp = d + ".%d.tmp" % os.getpid()
        os.makedirs(tmp, exist_ok=True)
        with open(os.path.join(tmp, "cuda_emul.h"), "w") as f:
            f.write(shim)
        try:
            subprocess.check_call(["g++"] + CXXFLAGS + ["-x", "c++-header", os.path.join(tmp, "cuda_emul.h"), "-o", os.path.join(tmp, "cuda_emul.h.gch")])
        except (OSError, subprocess.CalledProcessError):
            pass                                       # no precompiled header: the plain include still works
        try:
            os.rename(tmp, d)
        except OSError:
            pass                                       # another process was first
    return d if os.path.isdir(d) else os.path.join(HERE, "emul")


def build(kernel_source):
    os.makedirs(CACHE, exist_ok=True)
    src = host_source(kernel_source)
    with open(os.path.join(HERE, "emul", "cuda_emul.h")) as f:
        shim = f.read()
    key = hashlib.sha256((shim + src).encode()).hexdigest()[:24]
    lib = os.path.join(CACHE, key + ".so")
    if not os.path.isfile(lib):
        cpp = os.path.join(CACHE, key + ".cpp")
        with open(cpp, "w") as f:
            f.write(src)
        cmd = ["g++"] + CXXFLAGS + ["-shared", "-I", _shim_dir(shim), "-o", lib + ".%d.tmp" % os.getpid(), cpp]
        subprocess.check_call(cmd)
        os.replace(lib + ".%d.tmp" % os.getpid(), lib)
    return ctypes.CDLL(lib)


def keybits(value):
    """hb_template.cuh hb_keybits for one float64 / int64 key of the oracle's dict"""
    if isinstance(value, str):
        assert value == "NaN"
        return HB_KEY_NAN
    if isinstance(value, (float, numpy.floating)):
        v = numpy.float64(value)
        if v != v:
            return HB_KEY_NAN
        if v == 0.0:
            v = numpy.float64(0.0)
        return int(numpy.array([v]).view(numpy.uint64)[0])
    return int(numpy.array([value], dtype=numpy.int64).view(numpy.uint64)[0])


def run(spec, arrays, expected, options=None, grid=2, null_follower_maps=True):
    """Fill ``spec`` from ``arrays`` with the emulated kernel; ``expected`` (the oracle's contents) supplies the keys of
    the group axes.  Returns (kernel, contents in the oracle's form)."""
    options = dict(options or {})
    fields = hbspec.spec_fields(spec)
    cols = dict((n, numpy.ascontiguousarray(arrays[n])) for n in fields)
    coltypes = dict((n, (cols[n].dtype, False)) for n in fields)
    kern = codegen.generate(spec, coltypes, options)
    assert kern.window is None and not kern.fx_terms and not kern.deterministic, "the emulator runs kernels without hot windows and fixed point"
    lib = build(kern.source)
    lib.hb_emu_run.argtypes = [ctypes.POINTER(HbParams), ctypes.c_int, ctypes.c_int]
    p = HbParams()
    n = len(next(iter(cols.values()))) if cols else 0
    keep = []
    for slot, (name, dtype, is_scalar) in enumerate(kern.cols):
        assert cols[name].dtype == dtype and not is_scalar
        p.cols[slot] = cols[name].ctypes.data
    # group goals: the keys of this fill as ascending unsigned patterns (fillable._bind_stores, store.Store._slot_map)
    goal_patterns = {}
    for hid, h in enumerate(spec["hists"]):
        level = expected[hid]
        for a in h["group"]:
            gk = hbspec.tree_key(a["goal"])
            pats = sorted(set(keybits(k) for k in level.keys()))
            goal_patterns.setdefault(gk, pats)
            assert goal_patterns[gk] == pats, "histograms that share a group goal hold different keys"
            level = next(iter(level.values())) if len(level) else {}
    for gk, slot in kern.goal_aux.items():
        table = numpy.array(fillable.goal_table(goal_patterns[gk]), dtype=numpy.uint64)
        keep.append(table)
        p.aux[slot] = table.ctypes.data
    arenas, slotkeys = [], []
    for hid, h in enumerate(spec["hists"]):
        shape = tuple(h["shape"])
        if h["group"]:
            axes = [goal_patterns[hbspec.tree_key(a["goal"])] for a in h["group"]]
            total = int(numpy.prod([len(x) for x in axes]))
            slotmap = numpy.arange(max(total, 1), dtype=numpy.int32)          # slot = position of the key combination
            keep.append(slotmap)
            leader = kern.group_leader.get(hid, hid)
            p.aux[kern.group_aux[hid]] = 0 if (leader != hid and null_follower_maps) else slotmap.ctypes.data
            arena = numpy.zeros((max(total, 1),) + shape, dtype=numpy.float64)
            slotkeys.append(axes)
        else:
            arena = numpy.zeros(shape, dtype=numpy.float64)
            slotkeys.append(None)
        arenas.append(arena)
        p.hist[hid] = arena.ctypes.data
    p.n = n
    p.vec_ok = 1
    p.flags = 0
    rc = lib.hb_emu_run(ctypes.byref(p), int(grid), int(kern.threads))
    assert rc == 0
    out = []
    for hid, h in enumerate(spec["hists"]):
        if not h["group"]:
            out.append(arenas[hid])
            continue
        assert len(h["group"]) == 1, "nested group axes: not needed by the emulator's cases"
        d = {}
        for key in expected[hid].keys():
            d[key] = arenas[hid][slotkeys[hid][0].index(keybits(key))]
        out.append(d)
    return kern, out
