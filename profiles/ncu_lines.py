#!/usr/bin/env python
"""Per-source-line summary of an .ncu-rep (needs -lineinfo + --import-source on):
   python profiles/ncu_lines.py gpurun_out/x.ncu-rep [top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur, hdr, lines, func = None, None, {}, None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) == 2 and r[0] == "Function Name":
        func = r[1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if not hdr or len(r) != len(hdr):
        continue
    if r[0] != "":  # a source line row: aggregated over its SASS
        key = (cur, int(r[0]), r[1].strip())
        si = hdr.index("# Samples")
        ii = hdr.index("Instructions Executed")
        try:
            d = lines.setdefault(key, [0, 0, {}])
            d[0] += int(r[si])
            d[1] += int(r[ii])
            for j, h in enumerate(hdr):
                if h.startswith("stall_") and "Not Issued" not in h:
                    d[2][h] = d[2].get(h, 0) + int(r[j])
        except ValueError:
            pass
tot = sum(v[0] for v in lines.values()) or 1
toti = sum(v[1] for v in lines.values()) or 1
print(f"function: {func}\ntotal samples {tot}, warp instructions {toti}")
for (f, ln, src), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    st = sorted(v[2].items(), key=lambda kv: -kv[1])[:2]
    st = " ".join(f"{k[6:]}={c}" for k, c in st if c)
    print(f"{v[0]:7d} {v[0] / tot:6.3f} inst {v[1] / toti:6.3f}  {f}:{ln:<4d} {src[:70]:70s} {st}")
