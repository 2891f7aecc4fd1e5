import sys, os, torch
sys.path.insert(0, os.getcwd())
from oracle import reference_path as rp
from tests.test_gpu_scale import _scene
from tests.helpers import build_plugins
from triplaneturbo_b200 import ops
DEV="cuda"
P,V,H,W,R,C,ns,nimp = 1,1,128,128,64,32,64,128
sc, wts, rays_o, rays_d, c2w, dist = _scene(P,V,H,W,R,C,ns,nimp)
fx = {"space_cache": sc.to(DEV), **{k: v.to(DEV) for k, v in wts.items()}}
geom, rend = build_plugins(fx, DEV, ns, nimp); rend.train()
w = geom.decoder_weights()
pc = rp.PathConfig(num_samples_per_ray=ns, num_samples_per_ray_importance=nimp)
wc = {n: [wts[f"w_{n}_{i}"].clone() for i in range(3)] for n in ("sdf","feature")}
edges = ops.importance_sample(ops.cached_planes(fx["space_cache"]), ops.cached_wpack(w[:3], w[3:], geom._deformation_weights(), C), rend.path_scalars(), rays_o.to(DEV), rays_d.to(DEV), V*H*W, nimp, ns)
t0,t1 = edges[:,:-1], edges[:,1:]
with torch.no_grad():
    out = rend(rays_o.to(DEV), rays_d.to(DEV), None, torch.ones(3,device=DEV), space_cache=sc.to(DEV), text_embed=torch.zeros(P,4,device=DEV), camera_distances=dist.to(DEV), c2w=c2w.to(DEV), t_starts=t0, t_ends=t1)
ref = rp.render_forward(rays_o, rays_d, sc, wc, pc, torch.ones(3), dist, c2w, t_starts=t0.cpu(), t_ends=t1.cpu())
dn = (out["normal"].cpu()-ref["normal"].detach()).abs().max(-1).values
dg = (out["sdf_grad"].cpu()-ref["sdf_grad"].detach()).abs().max(-1).values
print("samples", dn.numel(), "normal err>1e-4:", int((dn>1e-4).sum()), ">1e-2:", int((dn>1e-2).sum()), "max", float(dn.max()))
print("sdf_grad err > 1e-3:", int((dg>1e-3).sum()), "max", float(dg.max()), "max |g|", float(ref["sdf_grad"].abs().max()))
# oracle pre-activations
pts = ref["points"].detach().reshape(1,-1,3)
enc = rp.interpolate_encodings(rp.rescale_points(pts, 1.0), sc, only_geo=True).reshape(-1, C)
z1 = enc @ wc["sdf"][0].T; h1 = z1.relu(); z2 = h1 @ wc["sdf"][1].T
amb = ((z1.abs() < 2e-5).any(-1) | (z2.abs() < 2e-5).any(-1))
nonempty = (enc.abs().sum(-1) > 0)
amb = amb & nonempty
print("ambiguous (|z|<2e-5) samples:", int(amb.sum()), "of non-empty", int(nonempty.sum()))
bad = dn > 1e-4
print("bad samples flagged ambiguous:", int((bad & amb).sum()), "of", int(bad.sum()))
gn = ref["sdf_grad"].detach().norm(dim=-1)
print("bad not flagged: |g| there", gn[bad & ~amb][:10], "dn", dn[bad & ~amb][:10], "dg", dg[bad & ~amb][:10])
zmin = torch.minimum(z1.abs().min(-1).values, z2.abs().min(-1).values)
print("zmin at bad-not-flagged", zmin[bad & ~amb][:10])
