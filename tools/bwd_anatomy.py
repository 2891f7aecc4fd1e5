#!/usr/bin/env python
"""Phase anatomy of k_bwd_tex_tc (colour backward).  Needs a timing build:
    TT_LIBNAME=libtt_timing.so bash triplaneturbo_b200/csrc/build.sh -DTT_WS_TIMING
    TT_B200_LIB=triplaneturbo_b200/lib/libtt_timing.so python tools/bwd_anatomy.py [workload=config3q]
Prints the cycles thread 0 of CTA 0 spent per phase and tile (clock64)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from triplaneturbo_b200 import _cabi

wl = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "config3q"])
dev = torch.device("cuda", 0)
rb = bench.RenderBench("x", wl, dev, 0, 1)
_cabi.load()
fn = ctypes.CDLL(_cabi.LIB_PATH).tt_debug_ws_prof
buf = (ctypes.c_ulonglong * 32)()
names = {20: "setup (ids, seeds, masks)", 21: "3 gathers + z1", 22: "layers, dW2/dW3, g1 staged", 23: "scatter (total)",
         24: "  merge: clear", 25: "  merge: hash insert", 26: "  merge: scan + lists", 27: "  merge: sums + reductions"}
for it in range(2):
    fn(buf)
    rb.step_resident(); torch.cuda.synchronize()
    fn(buf)
v = list(buf)
tiles = max(v[16], 1)
print("tiles seen by the profiled thread (forward consumer + backward):", v[16])
tot = sum(v[i] for i in (20, 21, 22, 23))
for i, n in names.items():
    print(f"   {n:32s} {v[i]:14d} cycles  {100.0 * v[i] / max(tot, 1):5.1f} %")
print("   unique texels per tile (merged):", v[28] / max(v[16], 1), "(divide by the backward's share of the tiles)")
