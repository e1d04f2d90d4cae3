#!/usr/bin/env python
"""Attribute the stall samples / executed instructions of an .ncu-rep (source page, needs -lineinfo) to the functions of
pbg_kernels.cuh by line range.  usage: ncu_regions.py rep [path/to/pbg_kernels.cuh]"""
import csv, subprocess, io, collections, re, sys, os
rep = sys.argv[1]
srcf = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pybullet_gym_b200", "csrc", "pbg_kernels.cuh")
marks = []
for i, l in enumerate(open(srcf), 1):
    m = re.match(r"\s*(?:template <[^>]*>\s*)?(?:static\s+)?__(?:device|global)__ .*?\b(\w+)\s*\(", l)
    if m and not l.strip().startswith("//"):
        marks.append((i, m.group(1)))
marks.append((10 ** 9, "end"))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None; lines = []
for r in rows:
    if r and r[0] == "Line No": h = r; continue
    if h and r and r[0].isdigit() and len(r) == len(h): lines.append(r)
iS, iI = h.index("# Samples"), h.index("Instructions Executed")
stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(float(r[iS]) for r in lines); toti = sum(float(r[iI]) for r in lines)
agg = collections.OrderedDict()
for r in lines:
    ln = int(r[0])
    name = "(before)"
    for (a, n), (b, _) in zip(marks, marks[1:]):
        if a <= ln < b: name = n; break
    if ln < marks[0][0]: name = "(helpers)"
    e = agg.setdefault(name, [0.0, 0.0, collections.Counter()])
    e[0] += float(r[iS]); e[1] += float(r[iI])
    for i in stall: e[2][h[i][6:]] += float(r[i])
print("%-22s %8s %7s  top stalls" % ("function", "samples", "inst"))
for name, (s, n, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if s / tot < 0.003: continue
    print("%-22s %7.1f%% %6.1f%%  %s" % (name, 100 * s / tot, 100 * n / toti, ", ".join("%s %.0f%%" % (k, 100 * v / max(s, 1)) for k, v in c.most_common(4))))
