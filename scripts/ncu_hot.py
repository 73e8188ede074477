#!/usr/bin/env python
"""Top stalled SASS instructions of an .ncu-rep (source page). usage: ncu_hot.py file.ncu-rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# first kernel block only
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
body = []
for r in rows[hdr_i + 1:]:
    if not r or r[0] == "Kernel Name":
        break
    body.append(r)
ci = hdr.index("Warp Stall Sampling (All Samples)"); si = hdr.index("Source"); ei = hdr.index("Instructions Executed")
tot = sum(int(r[ci] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
idx = sorted(range(len(body)), key=lambda i: -int(body[i][ci] or 0))[:top]
for i in sorted(idx):
    r = body[i]
    ctxl = body[i - 1][si].strip() if i > 0 else ""
    print(f"{i:5d} {int(r[ci]):7d} {100 * int(r[ci]) / tot:5.1f}%  exec {r[ei]:>8}  {r[si].strip()[:90]}    <- prev: {ctxl[:50]}")
