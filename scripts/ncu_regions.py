#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` dump of the decoder by code region (split at BAR.SYNC) and by opcode.
usage: ncu -i rep --page source --csv > src.csv; python scripts/ncu_regions.py src.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iN, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
iW, iWI = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
regions, cur = [], {"n": 0, "ex": 0, "sm": 0, "first": 0}
byop = collections.Counter(); byop_s = collections.Counter()
tot = 0
wf = collections.Counter(); wfi = collections.Counter()
for k, r in enumerate(rows[2:]):
    src = r[iS].strip(); ex = int(r[iN] or 0); sm = int(r[iSm] or 0)
    tok = src.split()
    op = tok[1] if tok and tok[0].startswith("@") else (tok[0] if tok else "?")
    op = op.rstrip(";")
    byop[op] += ex; byop_s[op] += sm; tot += ex
    wf[op] += int(r[iW] or 0); wfi[op] += int(r[iWI] or 0)
    cur["n"] += 1; cur["ex"] += ex; cur["sm"] += sm
    if op.startswith("BAR"):
        regions.append(cur); cur = {"n": 0, "ex": 0, "sm": 0, "first": k + 1}
regions.append(cur)
print("total warp-instructions executed: %d" % tot)
print("region (split at BAR.SYNC): first-line  static-instrs  executed  share  samples")
for i, g in enumerate(regions):
    if g["ex"]:
        print("%3d %6d %5d %12d %5.1f%% %7d" % (i, g["first"], g["n"], g["ex"], 100.0 * g["ex"] / tot, g["sm"]))
print("by opcode: executed share samples [smem wavefronts / ideal]")
for op, n in byop.most_common(40):
    extra = "  wf %d / %d" % (wf[op], wfi[op]) if wf[op] else ""
    print("%-22s %12d %5.1f%% %7d%s" % (op, n, 100.0 * n / tot, byop_s[op], extra))
