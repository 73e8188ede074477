#!/usr/bin/env python
"""Where do the warps of one kernel wait?  Reads an ncu report (--set full --import-source on), prints the SASS instructions
with the most stall samples and, for the kernels with warp roles, the samples per address range.
usage: python scripts/ncu_stalls.py report.ncu-rep [top_n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[start], [r for r in rows[start + 1:] if len(r) == len(rows[start])]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp]) for r in data)
print(f"{rows[0][1] if rows[0] else ''}\ntotal samples {tot}, {len(data)} SASS instructions")
agg = {}
for r in data:
    for c, h in stall:
        agg[h] = agg.get(h, 0) + int(r[c])
print("stall reasons:", ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:top_n]
for i in sorted(top):
    r = data[i]
    st = sorted(((int(r[c]), h[6:]) for c, h in stall if int(r[c]) > 0), reverse=True)[:2]
    print(f"{i:5d} {int(r[isamp]):6d} {100 * int(r[isamp]) / tot:5.1f}%  ex={r[iex]:>9s}  {r[isrc].strip()[:72]:72s} {st}")
