import sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from triplaneturbo_b200 import ops
wl = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "config3q"])
dev = torch.device("cuda", 0)
rb = bench.RenderBench("x", wl, dev, 0, 1)
with torch.no_grad():
    ro, rd, c2w, cd = rb.rays_d
    o, d = ro.reshape(-1, 3), rd.reshape(-1, 3)
    planes = ops.cached_planes(rb.sc_d.detach())
    w = rb.geom.decoder_weights()
    wpack = ops.cached_wpack(w[:3], w[3:], rb.geom._deformation_weights(), wl["C"])
    s = rb.rend.path_scalars()
    rpc = wl["V"] * wl["H"] * wl["W"]
    tv = ops.importance_sample(planes, wpack, s, o, d, rpc, wl["nimp"], wl["ns"])
    out = ops.render_fwd(planes, wpack, s, o, d, rpc, tv[:, :-1], tv[:, 1:], True, True)
    T, wt = out["trans"], out["weights"]
    N = T.numel()
    print("samples", N, "T>0:", float((T > 0).float().mean()))
    for eps in (1e-3, 1e-4, 1e-5, 1e-6, 1e-7, 1e-8, 1e-10):
        print(f"eps {eps:g}: T>eps {float((T > eps).float().mean()):.4f}   w>eps {float((wt > eps).float().mean()):.4f}")
    feat = out["features"]
    print("sdf abs mean", float(out["sdf"].abs().mean()))
