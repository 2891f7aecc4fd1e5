#!/usr/bin/env python
"""Phase anatomy of k_geo_ws (needs a -DTT_WS_TIMING build: `bash triplaneturbo_b200/csrc/build.sh -DTT_WS_TIMING`).
Runs the forward (sampler + fine pass) of a bench workload once and prints the cycles CTA 0's gather warps / consumer group 0
spent per phase.   python tools/ws_anatomy.py [workload=config3q]"""
import ctypes, json, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from triplaneturbo_b200 import _cabi, ops

wl = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "config3q"])
dev = torch.device("cuda", 0)
rb = bench.RenderBench("x", wl, dev, 0, 1)
L = _cabi.load()
fn = ctypes.CDLL(_cabi.LIB_PATH).tt_debug_ws_prof
buf = (ctypes.c_ulonglong * 32)()
names = {0: "M wait fin", 1: "M taps+sync", 2: "M gather", 3: "M wait de_ready", 4: "M normal pass", 7: "M other",
         8: "C wait full", 9: "C layers(+outputs)", 10: "C wait nacc", 15: "C other", 16: "C tiles"}
rb.rend.eval()
with torch.no_grad():
    for it in range(2):
        fn(buf)     # reset
        ro, rd, c2w, cd = rb.rays_d
        o, d = ro.reshape(-1, 3), rd.reshape(-1, 3)
        planes = ops.cached_planes(rb.sc_d.detach())
        w = rb.geom.decoder_weights()
        wpack = ops.cached_wpack(w[:3], w[3:], rb.geom._deformation_weights(), wl["C"])
        s = rb.rend.path_scalars()
        rpc = wl["V"] * wl["H"] * wl["W"]
        tv = ops.importance_sample(planes, wpack, s, o, d, rpc, wl["nimp"], wl["ns"])
        fn(buf); a = list(buf)
        out = ops.render_fwd(planes, wpack, s, o, d, rpc, tv[:, :-1], tv[:, 1:], True, False)
        fn(buf); b = list(buf)
for tag, v in (("proposal pass (k_classify + k_geo_ws<C,0>)", a), ("fine pass (k_geo_ws<C,1> + rest of render_fwd)", b)):
    tiles = max(v[16], 1)
    print(tag, "tiles of consumer group 0 of CTA 0:", v[16])
    for i, n in names.items():
        if i != 16:
            print(f"   {n:22s} {v[i] / tiles:10.0f} cycles/tile")
