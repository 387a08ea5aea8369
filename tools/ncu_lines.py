#!/usr/bin/env python
"""Per-phase, per-opcode and per-source-line breakdown of one kernel of an .ncu-rep (SASS page), joined with the line
table of the matching cubin (nvdisasm -g).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep build/c4.cubin build/c4.cu [top]
"""
import collections
import csv
import re
import subprocess
import sys


def main():
    rep, cubin, cu = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, data = rows[1], rows[2:]
    ix = dict((h, i) for i, h in enumerate(hdr))
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    cur, insts = None, []
    for line in dis.splitlines():
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            insts.append((int(m.group(1), 16), m.group(2), cur))
    inst = [int(r[ix["Instructions Executed"]] or 0) for r in data]
    samp = [int(r[ix["# Samples"]] or 0) for r in data]
    wave = [int(r[ix["L1 Wavefronts Shared"]] or 0) for r in data]
    ti, ts, tw = sum(inst), sum(samp), max(sum(wave), 1)
    print("warp instructions %d, samples %d, shared wavefronts %d, rows %d (cubin %d)" % (ti, ts, tw, len(data), len(insts)))
    bars = [i for i, r in enumerate(data) if "BAR" in r[ix["Source"]]]
    prev = 0
    for b in bars + [len(data) - 1]:
        a, e = prev, b + 1
        print("  rows %5d-%5d  inst %5.1f%%  samples %5.1f%%  smem wavefronts %5.1f%%   ..%s" % (
            a, e, 100.0 * sum(inst[a:e]) / ti, 100.0 * sum(samp[a:e]) / ts, 100.0 * sum(wave[a:e]) / tw, data[b][ix["Source"]].strip()[:34]))
        prev = e
    ops, sops, wops = collections.Counter(), collections.Counter(), collections.Counter()
    for r, i, s, w in zip(data, inst, samp, wave):
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
        op = m.group(2).split(".")[0] if m else "?"
        ops[op] += i
        sops[op] += s
        wops[op] += w
    print("opcode      inst%  samples%  wavefronts%")
    for op, c in ops.most_common(22):
        print("  %-10s %5.1f   %5.1f   %5.1f" % (op, 100.0 * c / ti, 100.0 * sops[op] / ts, 100.0 * wops[op] / tw))
    if len(insts) != len(data):
        print("cubin does not match the profiled kernel: no line table")
        return
    src = open(cu).read().split("\n")
    tmpl = open(__file__.rsplit("/", 2)[0] + "/histbook_b200/csrc/hb_template.cuh").read().split("\n")
    agg, sagg, wagg = collections.Counter(), collections.Counter(), collections.Counter()
    for (addr, text, loc), i, s, w in zip(insts, inst, samp, wave):
        agg[loc] += i
        sagg[loc] += s
        wagg[loc] += w
    print("line         inst%  samples%  wavefronts%")
    for loc, c in sorted(agg.items(), key=lambda kv: -(kv[1] / ti + sagg[kv[0]] / ts + wagg[kv[0]] / tw))[:top]:
        f, l = loc if loc else ("?", 0)
        text = (src[l - 1] if f.startswith("hb_fill") and l <= len(src) else tmpl[l - 1] if f.startswith("hb_template") and l <= len(tmpl) else "")[:100].strip()
        print("  %5.1f %5.1f %5.1f  %s:%d  %s" % (100.0 * c / ti, 100.0 * sagg[loc] / ts, 100.0 * wagg[loc] / tw, f[:11], l, text))


if __name__ == "__main__":
    main()
