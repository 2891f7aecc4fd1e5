#!/usr/bin/env python
"""Stand-alone plane sampler (tt_sample_planes_*): achieved algorithmic GB/s against the measured HBM peak.
Algorithmic bytes: forward N*M*(8K + 4*OS) + planes once; backward the same read + 8K*N*M (d/d grid) +
plane gradients once; second derivative: forward + backward streams.
    python tools/bench_sampler.py [C=40] [log2 M=22]   -> JSON lines (not the bench.py metric)"""
import json, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from triplaneturbo_b200.sampler import sample_planes
from triplaneturbo_b200.synthetic import camera_rays

C = int(sys.argv[1]) if len(sys.argv) > 1 else 40
M = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 22)
N, K, R = 4, 3, 256
dev = "cuda:0"
peak = 6458.4
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
gen = torch.Generator(dev).manual_seed(0)
planes = torch.randn(N * K, R, R, C, device=dev, generator=gen)


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        fn()
    ev[1].record(); torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / reps


def grids(kind):
    if kind == "random":
        return torch.rand(N * K, M, 2, device=dev, generator=gen) * 2 - 1
    # samples along camera rays in ray-major order (what the renderer's callers produce), projected on the 3 planes
    S = 64
    ro, rd, _, _ = camera_rays(N, 256, M // (256 * S), seed=2)
    t = torch.linspace(0.6, 2.6, S)
    pts = (ro.reshape(N, -1, 1, 3) + rd.reshape(N, -1, 1, 3) * t.view(1, 1, S, 1)).reshape(N, -1, 3)[:, :M].to(dev)
    from triplaneturbo_b200.sampler import project_onto_planes, PLANES
    return project_onto_planes(PLANES, pts).contiguous()


for kind in ("random", "rays"):
    grid = grids(kind)
    Mk = grid.shape[1]
    for concat in (False, True):
        OS = K * C if concat else C
        pl = planes.clone().requires_grad_(True)
        gr = grid.clone().requires_grad_(True)
        out = sample_planes(pl, gr, K, concat)
        go = torch.randn_like(out)
        stream = N * Mk * (8 * K + 4 * OS)
        pb = planes.numel() * 4
        ms_f = timed(lambda: sample_planes(planes, grid, K, concat))
        ms_b = timed(lambda: torch.autograd.grad(sample_planes(pl, gr, K, concat), [pl, gr], go))
        ms_b -= ms_f                                             # the backward launch alone (forward re-run inside)
        print(json.dumps({"op": "sample_planes", "points": kind, "mode": "concat(v2)" if concat else "sum(v1)", "N": N, "K": K,
                          "C": C, "R": R, "M": Mk, "fwd_ms": round(ms_f, 3), "bwd_ms": round(ms_b, 3),
                          "fwd_gbs": round((stream + pb) / ms_f / 1e6, 1), "fwd_frac_hbm": round((stream + pb) / ms_f / 1e6 / peak, 3),
                          "bwd_gbs": round((stream + 8 * K * N * Mk + 2 * pb) / ms_b / 1e6, 1),
                          "bwd_frac_hbm": round((stream + 8 * K * N * Mk + 2 * pb) / ms_b / 1e6 / peak, 3),
                          "logical_gather_tbs_fwd": round(N * Mk * K * 4 * 4 * C / ms_f / 1e9, 2), "hbm_peak_gbs": peak}))
