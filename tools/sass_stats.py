#!/usr/bin/env python
"""Static SASS statistics of a generated fill kernel (no GPU needed): NVRTC-compiles the kernel of a golden case with the
given options, disassembles the cubin with cuobjdump and counts instructions per phase of the sorted big-book kernel
(event phase = from the first column load to the first block barrier of the tile loop, scatter, summing loops), spills
and registers.  Used to judge a code-generation change before it is given GPU time.

    python tools/sass_stats.py c4 '{"sort_full": true}'
"""
import collections
import json
import os
import re
import subprocess
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import numpy                                                        # noqa: E402
from histbook_b200 import codegen, engine, spec as hbspec           # noqa: E402


def disassemble(cubin):
    with tempfile.NamedTemporaryFile(suffix=".cubin", delete=False) as f:
        f.write(cubin)
        path = f.name
    try:
        text = subprocess.check_output(["cuobjdump", "-sass", path]).decode()
        usage = subprocess.check_output(["cuobjdump", "-res-usage", path]).decode()
    finally:
        os.unlink(path)
    ins = []
    for line in text.splitlines():
        m = re.match(r"^\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if not m:
            continue
        txt = m.group(2).strip()
        t = txt.split()
        op = t[1] if t[0].startswith("@") else t[0]
        ins.append((int(m.group(1), 16), op, txt))
    return ins, usage


def main():
    case = sys.argv[1] if len(sys.argv) > 1 else "c4"
    options = json.loads(sys.argv[2]) if len(sys.argv) > 2 else {}
    with open(os.path.join(REPO, "tests", "golden", "cases.json")) as f:
        cases = dict((c["name"], c) for c in json.load(f)["cases"])
    spec = cases[case]["spec"]
    coltypes = dict((n, (numpy.dtype("int64" if n == "n" else "float64"), False)) for n in hbspec.spec_fields(spec))
    kern = codegen.generate(spec, coltypes, options)
    cubin = engine.compile_source(kern.source)
    if len(sys.argv) > 3:
        with open(sys.argv[3], "w") as f:
            f.write(kern.source)
    ins, usage = disassemble(cubin)
    # keep the main kernel only (the first function of the dump is the fill kernel; helpers such as the sqrt slow path follow)
    print("kernel smem %d B, threads %d, sorted %s" % (kern.smem_bytes, kern.threads, kern.sorted))
    for line in usage.splitlines():
        if "REG" in line:
            print(line.strip())
    print("instructions: %d" % len(ins))
    bars = [i for i, (a, op, txt) in enumerate(ins) if op.startswith("BAR")]
    ldg = [i for i, (a, op, txt) in enumerate(ins) if op.startswith("LDG")]
    print("barriers at", bars[:8], "... first LDG at", ldg[0] if ldg else None)
    if len(bars) >= 4 and ldg:
        # bars[0] = after zeroing; bars[1] = end of the event phase; bars[2] = after the scan; bars[3] = after the scatter
        phases = [("event", ldg[0], bars[1]), ("scan", bars[1], bars[2]), ("scatter", bars[2], bars[3])]
        ept = kern.sorted["ept"] if kern.sorted else 1
        for name, a, b in phases:
            hist = collections.Counter(op.split(".")[0] for addr, op, txt in ins[a:b])
            print("%-8s %5d instructions (%6.1f per event)  %s" % (name, b - a, (b - a) / float(ept),
                  " ".join("%s:%d" % kv for kv in hist.most_common(14))))
        if len(bars) >= 6:
            a, b = bars[3], bars[4]
            hist = collections.Counter(op.split(".")[0] for addr, op, txt in ins[a:b])
            print("%-8s %5d instructions (first group incl. find + loop + flush)  %s" % ("sum[0]", b - a, " ".join("%s:%d" % kv for kv in hist.most_common(12))))
    spills = collections.Counter(op.split(".")[0] for addr, op, txt in ins if op.startswith("STL") or op.startswith("LDL"))
    print("spills:", dict(spills))


if __name__ == "__main__":
    main()
