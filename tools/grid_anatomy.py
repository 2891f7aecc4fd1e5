#!/usr/bin/env python
"""Phase anatomy of the field query on the 512^3 grid (k_geo_ws<C,false,true>).  Needs the timing build (tools/bwd_anatomy.py).
    TT_B200_LIB=triplaneturbo_b200/lib/libtt_timing.so python tools/grid_anatomy.py [res=512]"""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from triplaneturbo_b200 import _cabi
from triplaneturbo_b200.synthetic import build_plugins, random_decoder, random_triplanes

res = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda", 0)
C, R = 32, 256
sc = random_triplanes(1, C, R, seed=0).to(dev)
fx = {"space_cache": sc, **{k: v.to(dev) for k, v in random_decoder(C, seed=1).items()}}
geom, _ = build_plugins(fx, dev, 64, 128)
_cabi.load()
fn = ctypes.CDLL(_cabi.LIB_PATH).tt_debug_ws_prof
buf = (ctypes.c_ulonglong * 32)()
names = {0: "M wait stage free", 1: "M tables / lines (+sync)", 2: "M gather / blend", 8: "C wait full", 9: "C layers(+outputs)", 15: "C other"}
with torch.no_grad():
    for it in range(2):
        fn(buf)
        geom.forward_field_grid(res, sc); torch.cuda.synchronize()
        fn(buf)
v = list(buf)
tiles = max(v[16], 1)
print("tiles of consumer group 0 of CTA 0:", tiles, "(the gather warps serve 2x as many)")
for i, n in names.items():
    print(f"   {n:26s} {v[i] / tiles:10.0f} cycles per consumer tile")
