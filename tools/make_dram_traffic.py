#!/usr/bin/env python
"""profiles/dram_traffic.json from the ncu summaries (profiles/r02_ncu_<workload>q_<kernel>_summary.txt, captured on a quarter
of each workload's rays by tools/evidence.sh ncu): dram__bytes_read.sum + dram__bytes_write.sum per launch, scaled x4."""
import json, os, re

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
KEYS = {"geo_proposal": "k_geo_ws:proposal", "geo_fine": "k_geo_ws:fine", "tex": None, "bwd_geo": "k_bwd_geo_tc:geo",
        "bwd_tex": "k_bwd_tex_tc"}


def dram_bytes(path):
    tot = 0.0
    for line in open(path):
        m = re.match(r"\s*dram__bytes_(read|write)\.sum\s+([0-9.]+)\s+(\w+)", line)
        if m:
            tot += float(m.group(2)) * UNIT[m.group(3)]
    return tot


out = {}
for wl in ("config3", "config2"):
    d = {}
    for tag, key in KEYS.items():
        p = os.path.join(ROOT, "profiles", f"r02_ncu_{wl}q_{tag}_summary.txt")
        if not os.path.exists(p):
            continue
        b = 4.0 * dram_bytes(p)
        if tag == "tex":      # the colour forward is k_tex_tc1 when 6C <= 192 (config 3), k_tex_tc otherwise
            d["k_tex_tc1" if wl == "config3" else "k_tex_tc"] = b
            d.setdefault("k_tex_tc", b)
        else:
            d[key] = b
    if "k_geo_ws:proposal" in d and "k_geo_ws:fine" in d:
        d["k_geo_ws"] = 0.5 * (d["k_geo_ws:proposal"] + d["k_geo_ws:fine"])      # bench.py averages the two launches of a step
    out[wl] = d
out["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full of the final round-2 build; captured on a "
                "quarter of each workload's rays (config3q / config2q: same planes, samples and kernels, 4x shorter replays) and "
                "scaled x4 by tools/make_dram_traffic.py; raw exports: profiles/r02_ncu_*_summary.txt")
json.dump(out, open(os.path.join(ROOT, "profiles", "dram_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
