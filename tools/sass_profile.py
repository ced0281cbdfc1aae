#!/usr/bin/env python
"""Executed-instruction profile of one kernel by SOURCE LINE, from an `ncu --set full` report.

ncu's per-instruction counters (`--page source --print-source sass`) are joined, by instruction offset,
with `nvdisasm --print-line-info-inline` of the object the kernel was built from, and aggregated

  * per outermost frame (the line of the __global__ function an instruction was inlined into), and
  * per innermost line of this repository's own sources,

so that "where do the issue slots go" has an answer in terms of the code (profiles/*.lines.txt).

usage: python tools/sass_profile.py <report.ncu-rep> <object.o|.so> <kernel-mangled-substring> [units]
       `units` divides the counts (e.g. the number of warp-rows: W*H/32) to give instructions per unit.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def ncu_sass(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    res = []
    for r in rows[hdr_i + 1:]:
        if len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        res.append((int(d["Address"], 16), d["Source"].strip(), int(d["Instructions Executed"]), int(d["# Samples"] or 0)))
    base = res[0][0]
    return [(a - base, s, n, smp) for a, s, n, smp in res]


def disasm_lines(obj, kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    txt = ""
    for f in sorted(os.listdir(tmp)):
        if f.endswith(".cubin"):
            t = subprocess.run(["nvdisasm", "--print-line-info-inline", "-c", os.path.join(tmp, f)], capture_output=True,
                               text=True).stdout
            if kernel_sub in t:
                txt = t
                break
    sec = None
    frames, table = [], {}
    for line in txt.split("\n"):
        m = re.match(r"^\.text\.(\S+):", line)
        if m:
            sec = m.group(1)
            frames = []
            continue
        if sec is None or kernel_sub not in sec:
            continue
        m = re.match(r'^\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
        if m:
            if not getattr(disasm_lines, "_open", False):
                frames = []
                disasm_lines._open = True
            frames.append((m.group(1), int(m.group(2))))
            continue
        m = re.match(r"^\s+/\*([0-9a-f]+)\*/\s+(.*);", line)
        if m:
            disasm_lines._open = False
            table[int(m.group(1), 16)] = (list(frames), m.group(2).strip())
    return table


def main():
    rep, obj, ksub = sys.argv[1:4]
    units = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
    sass = ncu_sass(rep)
    lines = disasm_lines(obj, ksub)
    outer = collections.Counter()
    inner = collections.Counter()
    outer_smp = collections.Counter()
    total = 0
    missing = 0
    for off, src, n, smp in sass:
        total += n
        fr = lines.get(off)
        if not fr or not fr[0]:
            missing += n
            continue
        frames = fr[0]
        o = frames[-1]
        outer[(os.path.basename(o[0]), o[1])] += n
        outer_smp[(os.path.basename(o[0]), o[1])] += smp
        own = [f for f in frames if "/csrc/" in f[0]]
        i = own[0] if own else frames[0]
        inner[(os.path.basename(i[0]), i[1])] += n
    print("total executed warp-instructions: %d  (%.1f per unit), unattributed %d" % (total, total / units, missing))
    src_cache = {}

    def text(f, l):
        for d in ("image-lens-reproject_b200/csrc",):
            p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d, f)
            if os.path.exists(p):
                if p not in src_cache:
                    src_cache[p] = open(p, errors="ignore").read().split("\n")
                if 0 < l <= len(src_cache[p]):
                    return src_cache[p][l - 1].strip()[:100]
        return ""
    print("\n== by outermost frame (line of the kernel body) ==")
    for (f, l), n in sorted(outer.items(), key=lambda kv: (kv[0][0], kv[0][1])):
        if n / total >= 0.002:
            print("%-18s %4d %8.1f %5.1f%% smp %5d | %s" % (f, l, n / units, 100.0 * n / total, outer_smp[(f, l)], text(f, l)))
    print("\n== by innermost own-source line (top 60) ==")
    for (f, l), n in inner.most_common(int(os.environ.get("TOPN", "60"))):
        print("%-18s %4d %8.1f %5.1f%% | %s" % (f, l, n / units, 100.0 * n / total, text(f, l)))


if __name__ == "__main__":
    main()
