#!/usr/bin/env python
"""BASELINE config 5: mesh-export field query, SDF at the res^3 isosurface grid generated in-kernel
(triplaneturbo_executable/utils/mesh_exporter.py:78-141 evaluates the same grid un-chunked through geometry.forward_field).
    python tools/bench_field_query.py [res=512] [C=32]      -> one JSON line (not the bench.py metric)"""
import json, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from tests.helpers import build_plugins
from triplaneturbo_b200.synthetic import random_decoder, random_triplanes

res = int(sys.argv[1]) if len(sys.argv) > 1 else 512
C = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = "cuda:0"
sc = random_triplanes(1, C, 256, seed=0).to(dev)
fx = {"space_cache": sc, **{k: v.to(dev) for k, v in random_decoder(C, seed=1).items()}}
geom, _ = build_plugins(fx, dev, 64, 128)
with torch.no_grad():
    for _ in range(2):
        sdf, _ = geom.forward_field_grid(res, sc)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(3):
        sdf, _ = geom.forward_field_grid(res, sc)
    ev[1].record(); torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / 3
M = res ** 3
f_sdf = 2 * (64 * C + 64 * 64 + 64) + 2 * (64 * C + 64 * 64 + 192)      # SDF + deformation decoders (F_sdf + F_def)
print(json.dumps({"workload": f"config5: {res}^3 SDF + deformation grid query, R=256, C={C}", "points": M, "ms": ms,
                  "points_per_s": M / (ms * 1e-3), "dense_tflops": M * f_sdf / (ms * 1e-3) / 1e12,
                  "logical_gather_tbs": M * 12 * 4 * C / (ms * 1e-3) / 1e12, "sdf_mean": float(sdf.mean())}))
