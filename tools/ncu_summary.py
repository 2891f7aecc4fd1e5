#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics per kernel and the top stall sites (needs `ncu` on PATH, no GPU).
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [n_top]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    print("##", r[hdr.index("Kernel Name")][:90])
    for w in want:
        if w in hdr:
            print(f"   {w:70s} {r[hdr.index(w)]} {rows[1][hdr.index(w)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kern = None; h = None; data = collections.defaultdict(list)
for r in csv.reader(src.splitlines()):
    if r and r[0] == "Kernel Name": kern = r[1][:60]; continue
    if r and r[0] == "Address": h = r; continue
    if kern and h and len(r) == len(h): data[kern].append(r)
for k, v in data.items():
    si = h.index("# Samples"); so = h.index("Source")
    tot = sum(int(x[si] or 0) for x in v)
    st = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    agg = sorted(((sum(int(x[i] or 0) for x in v), h[i]) for i in st), reverse=True)[:7]
    print("== stalls", k, "samples", tot, [(n, round(100 * c / max(tot, 1), 1)) for c, n in agg])
    seen = set()
    for x in sorted(v, key=lambda x: -int(x[si] or 0)):
        if x[0] in seen: continue
        seen.add(x[0])
        print(f"   {100 * int(x[si] or 0) / max(tot, 1):5.2f}%  {x[so][:100]}")
        if len(seen) >= ntop: break
