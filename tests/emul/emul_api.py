"""TEST TOOLING: drive the host-emulated build of the kernel sources (tests/emul/libtt_emul.so) with numpy arrays.

This exists so kernel logic can be debugged in the CPU-only build container.  It is not a backend of the
package: ``triplaneturbo_b200`` only ever loads the CUDA library.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from triplaneturbo_b200 import _cabi

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libtt_emul.so")
import glob
SRC = glob.glob(os.path.join(HERE, "..", "..", "triplaneturbo_b200", "csrc", "*.cu*"))      # every kernel source
SRC += [os.path.join(HERE, "cuda_emul.h"), os.path.join(HERE, "..", "..", "include", "triplane_b200.h")]


def lib():
    if not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in SRC):
        subprocess.check_call(["bash", os.path.join(HERE, "build_emul.sh")])
    return _cabi.bind(C.CDLL(SO))


def f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def ptr(a):
    if a is None:
        return None
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def aligned_zeros(shape, dtype=np.float32):
    n = int(np.prod(shape))
    raw = np.zeros(n * 4 + 64, dtype=np.uint8)
    off = (-raw.ctypes.data) % 16
    out = raw[off:off + n * 4].view(dtype).reshape(shape)
    assert out.ctypes.data % 16 == 0
    return out


class Emul:
    def __init__(self):
        self.L = lib()

    def ok(self, code, what):
        assert code == 0, f"{what}: {self.L.tt_last_error().decode()}"

    def config(self, C_, R, P, rays_per_cache=1, radius=1.0, bias=0.5, inv_std=None, car=1.0, near=0.1, far=4.0,
               step=0.05, flags=1):
        if inv_std is None:
            inv_std = float(np.exp(np.float32(0.4605) * np.float32(10.0)))
        return _cabi.TTConfig(C_, R, P, rays_per_cache, radius, bias, inv_std, car, near, far, step, flags)

    def set_option(self, name, value):
        self.ok(self.L.tt_set_option(name.encode(), int(value)), "set_option")

    def set_impl(self, impl):
        self.ok(self.L.tt_set_impl(impl), "set_impl")

    def pack_weights(self, w, C_):
        wp = aligned_zeros(self.L.tt_wpack_floats(C_))
        args = []
        for name in ("sdf", "feature", "deformation"):
            ws = w.get(name)
            args += [ptr(f32(x)) for x in ws] if ws is not None else [None] * 3
        keep = [f32(x) for name in ("sdf", "feature", "deformation") if w.get(name) is not None for x in w[name]]
        args = [ptr(k) for k in keep] + [None] * (9 - len(keep))
        self.ok(self.L.tt_pack_weights(*args, C_, ptr(wp), None), "pack_weights")
        return wp

    def repack(self, sc, C_=None, off_geo=0, off_tex=0):
        sc = f32(sc)
        P, _, Csrc, R, _ = sc.shape
        C_ = C_ or Csrc
        dst = aligned_zeros((P, 6, R, R, C_))
        self.ok(self.L.tt_repack_planes(ptr(sc), P, Csrc, off_geo, off_tex, C_, R, ptr(dst), None), "repack")
        return dst

    def repack_bwd(self, g):
        P, _, R, _, C_ = g.shape
        out = np.zeros((P, 6, C_, R, R), np.float32)
        self.ok(self.L.tt_repack_planes_bwd(ptr(g), P, C_, R, ptr(out), None), "repack_bwd")
        return out

    def repack_bwd_split(self, g, Cdst, off_geo, off_tex):
        P, _, R, _, C_ = g.shape
        out = np.zeros((P, 6, Cdst, R, R), np.float32)
        self.ok(self.L.tt_repack_planes_bwd_split(ptr(g), P, Cdst, off_geo, off_tex, C_, R, ptr(out), None), "repack_bwd_split")
        return out

    def geometry_fwd(self, planes, wp, cfg, points=None, grid_res=0, normal=True, features=True, deform=False):
        M = points.shape[1] if points is not None else grid_res ** 3
        N = cfg.P * M
        o = dict(sdf=np.zeros(N, np.float32), sdf_orig=np.zeros(N, np.float32),
                 features=np.zeros((N, 3), np.float32) if features else None,
                 normal=np.zeros((N, 3), np.float32) if normal else None,
                 sdf_grad=np.zeros((N, 3), np.float32) if normal else None,
                 deformation=np.zeros((N, 3), np.float32) if deform else None)
        pts = f32(points) if points is not None else None
        self.ok(self.L.tt_geometry_fwd(ptr(planes), ptr(wp), C.byref(cfg), ptr(pts), M, grid_res, ptr(o["sdf"]),
                                       ptr(o["sdf_orig"]), ptr(o["features"]), ptr(o["normal"]),
                                       ptr(o["sdf_grad"]), ptr(o["deformation"]), None), "geometry_fwd")
        return o

    def geometry_bwd(self, planes, wp, cfg, points, g_sdf=None, g_features=None, g_normal=None, g_sdf_grad=None):
        pts = f32(points)
        M = pts.shape[1]
        N = cfg.P * M
        scratch = np.zeros(self.L.tt_geometry_bwd_scratch_floats(C.byref(cfg), N), np.float32)
        gplanes = aligned_zeros(planes.shape)
        gw = aligned_zeros(self.L.tt_wgrad_floats(cfg.C))
        a = [None if x is None else f32(x) for x in (g_sdf, g_features, g_normal, g_sdf_grad)]
        self.ok(self.L.tt_geometry_bwd(ptr(planes), ptr(wp), C.byref(cfg), ptr(pts), M, *[ptr(x) for x in a],
                                       ptr(scratch), ptr(gplanes), ptr(gw), None), "geometry_bwd")
        return gplanes, self.split_wgrad(gw, cfg.C)

    def field_bwd(self, planes, wp, cfg, points, g_sdf=None, g_def=None):
        pts = f32(points)
        M = pts.shape[1]
        N = cfg.P * M
        scratch = np.zeros(self.L.tt_geometry_bwd_scratch_floats(C.byref(cfg), N), np.float32)
        gplanes = aligned_zeros(planes.shape)
        gw = aligned_zeros(self.L.tt_wgrad_floats(cfg.C))
        gwd = aligned_zeros(self.L.tt_wgrad_def_floats(cfg.C))
        a = [None if x is None else f32(x) for x in (g_sdf, g_def)]
        self.ok(self.L.tt_field_bwd(ptr(planes), ptr(wp), C.byref(cfg), ptr(pts), M, ptr(a[0]), ptr(a[1]), ptr(scratch),
                                    ptr(gplanes), ptr(gw), ptr(gwd), None), "field_bwd")
        Cc = cfg.C
        gd = [gwd[:64 * Cc].reshape(64, Cc).copy(), gwd[64 * Cc:64 * Cc + 4096].reshape(64, 64).copy(),
              gwd[64 * Cc + 4096:].reshape(3, 64).copy()]
        return gplanes, self.split_wgrad(gw, Cc)[:3], gd

    def split_wgrad(self, gw, C_):
        off = (C.c_int64 * 6)()
        self.ok(self.L.tt_wgrad_offsets(C_, off), "wgrad_offsets")
        shapes = [(64, C_), (64, 64), (1, 64), (64, 3 * C_), (64, 64), (3, 64)]
        return [gw[off[i]:off[i] + s[0] * s[1]].reshape(s).copy() for i, s in enumerate(shapes)]

    def importance_sample(self, planes, wp, cfg, rays_o, rays_d, n_imp, n_fine, jit0=None, jit1=None):
        o, d = f32(rays_o).reshape(-1, 3), f32(rays_d).reshape(-1, 3)
        n = o.shape[0]
        scratch = np.zeros(self.L.tt_sample_scratch_floats(n, n_imp), np.float32)
        t = np.zeros((n, n_imp + n_fine + 2), np.float32)
        j0 = None if jit0 is None else f32(jit0)
        j1 = None if jit1 is None else f32(jit1)
        self.ok(self.L.tt_importance_sample(ptr(planes), ptr(wp), C.byref(cfg), ptr(o), ptr(d), n, n_imp, n_fine,
                                            ptr(j0), ptr(j1), ptr(scratch), ptr(t), None), "importance_sample")
        return t

    def render_fwd(self, planes, wp, cfg, rays_o, rays_d, t_starts, t_ends):
        o, d = f32(rays_o).reshape(-1, 3), f32(rays_d).reshape(-1, 3)
        t0, t1 = f32(t_starts), f32(t_ends)
        n, S = t0.shape
        N = n * S
        out = dict(acc=np.zeros((n, 10), np.float32), sdf=np.zeros(N, np.float32), sdf_orig=np.zeros(N, np.float32),
                   sdf_grad=np.zeros((N, 3), np.float32), normal=np.zeros((N, 3), np.float32),
                   features=np.zeros((N, 3), np.float32), weights=np.zeros(N, np.float32),
                   trans=np.zeros(N, np.float32), tex_masks=np.zeros((N, 4), np.uint64))
        scratch = np.zeros(self.L.tt_render_fwd_scratch_floats(n, S), np.float32)
        self.ok(self.L.tt_render_fwd(ptr(planes), ptr(wp), C.byref(cfg), ptr(o), ptr(d), n, ptr(t0), ptr(t1), S, S,
                                     *[ptr(out[k]) for k in ("acc", "sdf", "sdf_orig", "sdf_grad", "normal",
                                                              "features", "weights", "trans")],
                                     out["tex_masks"].ctypes.data_as(C.c_void_p), ptr(scratch), None),
                "render_fwd")
        return out

    def render_bwd(self, planes, wp, cfg, rays_o, rays_d, t_starts, t_ends, fwd, g_acc, g_sdf=None, g_sdf_grad=None,
                   g_normal=None, g_features=None, g_weights=None, rgb_scale=1.0):
        o, d = f32(rays_o).reshape(-1, 3), f32(rays_d).reshape(-1, 3)
        t0, t1 = f32(t_starts), f32(t_ends)
        n, S = t0.shape
        scratch = np.zeros(self.L.tt_render_bwd_scratch_floats(C.byref(cfg), n, S), np.float32)
        gplanes = aligned_zeros(planes.shape)
        gw = aligned_zeros(self.L.tt_wgrad_floats(cfg.C))
        gis = np.zeros(1, np.float32)
        opt = [None if x is None else f32(x) for x in (g_sdf, g_sdf_grad, g_normal, g_features, g_weights)]
        ga = f32(g_acc)
        self.ok(self.L.tt_render_bwd(ptr(planes), ptr(wp), C.byref(cfg), ptr(o), ptr(d), n, ptr(t0), ptr(t1), S, S,
                                     ptr(fwd["acc"]), ptr(fwd["sdf"]), ptr(fwd["sdf_grad"]), ptr(fwd["features"]),
                                     ptr(fwd["trans"]), fwd["tex_masks"].ctypes.data_as(C.c_void_p), ptr(ga),
                                     *[ptr(x) for x in opt], rgb_scale, ptr(scratch),
                                     ptr(gplanes), ptr(gw), ptr(gis), None), "render_bwd")
        return gplanes, self.split_wgrad(gw, cfg.C), float(gis[0])

    # ---- stand-alone compositor
    def composite_fwd(self, alphas, values):
        a = f32(alphas); n, S = a.shape
        D = 0 if values is None else values.shape[-1]
        v = None if values is None else f32(values)
        w, T, out = np.zeros_like(a), np.zeros_like(a), np.zeros((n, max(D, 1)), np.float32)
        self.ok(self.L.tt_composite_fwd(ptr(a), ptr(v), n, S, D, ptr(w), ptr(T), ptr(out), None), "composite_fwd")
        return w, T, out

    def composite_bwd(self, alphas, values, trans, g_out, g_weights):
        a = f32(alphas); n, S = a.shape
        D = 0 if values is None else values.shape[-1]
        v = None if values is None else f32(values)
        ga = np.zeros_like(a); gv = None if values is None else np.zeros_like(v)
        go = None if g_out is None else f32(g_out); gw = None if g_weights is None else f32(g_weights)
        self.ok(self.L.tt_composite_bwd(ptr(a), ptr(v), ptr(f32(trans)), ptr(go), ptr(gw), n, S, D, ptr(ga), ptr(gv), None), "composite_bwd")
        return ga, gv

    # ---- stand-alone plane sampler (tt_sampler.cuh) ------------------------------------------------------------
    def to_channel_last(self, x):
        x = f32(x); B, C_, HW = x.shape
        y = aligned_zeros((B, HW, C_))
        self.ok(self.L.tt_to_channel_last(ptr(x), B, C_, HW, ptr(y), None), "to_channel_last")
        return y

    def from_channel_last(self, y):
        y = f32(y); B, HW, C_ = y.shape
        x = aligned_zeros((B, C_, HW))
        self.ok(self.L.tt_from_channel_last(ptr(y), B, C_, HW, ptr(x), None), "from_channel_last")
        return x

    @staticmethod
    def _al(a):
        out = aligned_zeros(np.shape(a)); out[...] = a
        return out

    def sample_fwd(self, planes, grid, K, concat):
        planes, grid = self._al(planes), f32(grid)
        NK, H, W, C_ = planes.shape; N, M = NK // K, grid.shape[1]
        out = aligned_zeros((N, M, K * C_ if concat else C_))
        self.ok(self.L.tt_sample_planes_fwd(ptr(planes), N, K, C_, H, W, ptr(grid), M, int(concat), ptr(out), None), "sample_fwd")
        return out

    def sample_bwd(self, planes, grid, K, concat, g_out):
        planes, grid, g_out = self._al(planes), f32(grid), self._al(g_out)
        NK, H, W, C_ = planes.shape; N, M = NK // K, grid.shape[1]
        gp, gg = aligned_zeros(planes.shape), np.zeros_like(grid)
        self.ok(self.L.tt_sample_planes_bwd(ptr(planes), N, K, C_, H, W, ptr(grid), M, int(concat), ptr(g_out), ptr(gp), ptr(gg), None), "sample_bwd")
        return gp, gg

    def sample_bwdbwd(self, planes, grid, K, concat, g_out, gg_planes, gg_grid):
        planes, grid, g_out = self._al(planes), f32(grid), self._al(g_out)
        ggp = None if gg_planes is None else self._al(gg_planes)
        ggg = None if gg_grid is None else f32(gg_grid)
        NK, H, W, C_ = planes.shape; N, M = NK // K, grid.shape[1]
        ggo, gp, gg = aligned_zeros(g_out.shape), aligned_zeros(planes.shape), np.zeros_like(grid)
        self.ok(self.L.tt_sample_planes_bwdbwd(ptr(planes), N, K, C_, H, W, ptr(grid), M, int(concat), ptr(g_out), ptr(ggp), ptr(ggg),
                                               ptr(ggo), ptr(gp), ptr(gg), None), "sample_bwdbwd")
        return ggo, gp, gg
