#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libec_b200.so (evidence of tcgen05 / TMEM / TMA use, B200_PROFILING.md):
    tcgen05.mma -> UTC*MMA | tcgen05.ld / st -> LDTM / STTM | TMA -> UTMALDG / UBLKCP | mma.sync -> HMMA | ldmatrix -> LDSM
usage: python scripts/sass_histogram.py [lib.so] [out.txt]     (default: the in-tree library -> profiles/r02_sass_opcodes.txt)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "HMMA", "LDSM", "LDGSTS",
       "SYNCS", "MUFU", "FFMA2", "REDUX", "SHFL"]


def demangle(names):
    try:
        out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.split("\n")
        return [o if o else n for o, n in zip(out, names)]
    except OSError:
        return names


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "embodied_captioning_b200", "libec_b200.so")
    out_path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "r02_sass_opcodes.txt")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    counts, order, cur = {}, [], None
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for o in OPS:
                if op.startswith(o):
                    counts[cur][o] += 1
            if ".2CTA" in line and op.startswith("UTC"):
                counts[cur]["UTC*.2CTA"] += 1
    names = demangle(order)
    cols = OPS + ["UTC*.2CTA"]
    used = [c for c in cols if any(counts[f][c] for f in order)]
    lines = [f"# SASS opcode histogram per kernel of {os.path.basename(lib)} (cuobjdump -sass; scripts/sass_histogram.py)",
             "# UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UBLKCP = TMA tensor / bulk copies, HMMA = mma.sync",
             "instr  " + "  ".join(f"{c:>9}" for c in used) + "  kernel"]
    tot = collections.Counter()
    for f, nm in sorted(zip(order, names), key=lambda t: t[1]):
        c = counts[f]
        short = re.sub(r"\(.*", "", nm)
        short = re.sub(r"^void ", "", short)
        lines.append(f"{c['_total']:5d}  " + "  ".join(f"{c[k]:9d}" for k in used) + f"  {short[:110]}")
        tot.update(c)
    lines.append(f"{tot['_total']:5d}  " + "  ".join(f"{tot[k]:9d}" for k in used) + "  TOTAL")
    text = "\n".join(lines) + "\n"
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    with open(out_path, "w") as fh:
        fh.write(text)
    print(text if len(sys.argv) > 3 else f"wrote {out_path}: {len(order)} kernels, UTCHMMA {tot['UTCHMMA']}, LDTM {tot['LDTM']}, "
          f"UTMALDG {tot['UTMALDG']}, UBLKCP {tot['UBLKCP']}, HMMA {tot['HMMA']}")


if __name__ == "__main__":
    main()
