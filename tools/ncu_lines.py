#!/usr/bin/env python
"""Per-source-line stall samples and executed instructions of one kernel in an .ncu-rep (needs -lineinfo builds).
    python tools/ncu_lines.py rep.ncu-rep <kernel substring> [n_top]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
fpath = func = None; hdr = None; rows = []
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] == "File Path": fpath = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": func = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and func and pat in func and r[0].isdigit():
        extra = len(r) - len(hdr)          # unescaped quotes inside the source text split the field
        if extra < 0: continue
        src = ",".join(r[1:2 + extra]); v = r[2 + extra:]
        try: rows.append((int(v[4] or 0), int(v[5] or 0), fpath, int(r[0]), src.strip()[:110]))
        except ValueError: pass
tot = sum(x[0] for x in rows) or 1; toti = sum(x[1] for x in rows) or 1
print(f"kernel ~ {pat}: {tot} samples, {toti} warp-instructions")
for s, i, f, ln, src in sorted(rows, reverse=True)[:ntop]:
    print(f"{100*s/tot:6.2f}% smp {100*i/toti:6.2f}% ins  {f}:{ln:<4d} {src}")
