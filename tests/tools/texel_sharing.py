#!/usr/bin/env python
"""How many bilinear taps of a 128-sample tile fall on the same texel, for different tile shapes, on the REAL config-3
sample positions (oracle sampler on a 32 x 32 pixel window of a 512^2 view).  CPU only (~1 min).  DESIGN 3.5 item 10."""
import sys, torch, numpy as np
import os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from triplaneturbo_b200.synthetic import camera_rays, random_triplanes, random_decoder
from oracle import reference_path as rp
torch.manual_seed(0)
P,V,H,R,C,ns,nimp=1,1,512,256,32,64,128
sc = random_triplanes(P, C, R, seed=0); wts = random_decoder(C, seed=1)
rays_o, rays_d, c2w, dist = camera_rays(1, H, H, seed=2, views_per_prompt=1)
y0,x0=240,240; cr=32
ro = rays_o[:, y0:y0+cr, x0:x0+cr].contiguous(); rd = rays_d[:, y0:y0+cr, x0:x0+cr].contiguous()
pc = rp.PathConfig(num_samples_per_ray=ns, num_samples_per_ray_importance=nimp)
wc = {n: [wts[f"w_{n}_{i}"].clone() for i in range(3)] for n in ("sdf", "feature")}
with torch.no_grad():
    t0, t1 = rp.sample_intervals(ro, rd, sc, wc, pc)
tm = 0.5*(t0+t1)                       # [cr*cr, S]
S = tm.shape[1]
pts = ro.reshape(-1,1,3) + rd.reshape(-1,1,3)*tm[...,None]   # [rays,S,3]
pts = pts.reshape(cr,cr,S,3).numpy()
print(pts.shape, S)
R=256
def cells(p):  # p [...,3] -> three plane cell ids (x0,y0) per plane, -1 if fully outside
    g = (p+1)/2*(R-1)           # align_corners=True style? approx
    i0 = np.floor(g).astype(int)
    inside = ((i0>=-1)&(i0<R)).all(-1)
    out=[]
    for ax,ay in ((0,1),(0,2),(2,1)):
        out.append(np.where(inside, i0[...,ay]*1024+i0[...,ax], -1))
    return out, inside
cs, inside = cells(pts)
print("inside frac", inside.mean())
def texels(c):  # unique texels of 2x2 cells
    c=c[c>=0]
    t=np.concatenate([c, c+1, c+1024, c+1025])
    return len(np.unique(t)), 4*len(c)
def analyse(shape_rays, ns):  # tile = (ry x rx rays) x ns consecutive samples
    ry,rx=shape_rays
    tot_u=0; tot_t=0
    for y in range(0,32,ry):
        for x in range(0,32,rx):
            for s in range(0,192,ns):
                for k in range(3):
                    u,t=texels(cs[k][y:y+ry,x:x+rx,s:s+ns].ravel())
                    tot_u+=u; tot_t+=t
    return tot_t/max(tot_u,1)
for sh,ns in (((1,1),128),((1,2),64),((2,2),32),((2,4),16),((4,4),8),((4,8),4),((8,8),2),((8,16),1),((1,8),16),((1,16),8),((1,32),4),((2,8),8)):
    print(sh,ns, "taps/unique texel = %.2f"%analyse(sh,ns))
