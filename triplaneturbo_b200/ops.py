"""torch-facing wrappers of the C ABI: device memory, streams and autograd plumbing only.

Every function here takes CUDA fp32 tensors, hands raw device pointers and the current CUDA stream to
``libtriplane_b200.so`` and returns torch tensors that own the outputs.  Nothing in this file computes on the
path: if the library or the GPU is missing, the call raises (``_cabi.TTError``).
"""
import ctypes as C
import math
import weakref
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _cabi

HIDDEN = 64
ACC = 10          # TT_ACC of include/triplane_b200.h
SUPPORTED_C = (8, 16, 32, 40, 64)


@dataclass
class PathScalars:
    """Scalars of the path (configs/TriplaneTurbo_v1.yaml:73-150 of the reference)."""
    radius: float = 1.0
    sdf_bias_radius: float = 0.5
    inv_std: float = 100.0
    cos_anneal_ratio: float = 1.0
    near_plane: float = 0.1
    far_plane: float = 4.0
    render_step_size: float = 1.732 * 2 * 1.0 / 64
    image_h: int = 0        # hint for tt_render_bwd: the rays are [B,image_h,image_w] images (patch-ordered sample lists)
    image_w: int = 0


def _lib():
    return _cabi.load()


def _ptr(t: Optional[Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(dev) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _need(t: Tensor, name: str) -> Tensor:
    if not isinstance(t, Tensor) or not t.is_cuda:
        raise _cabi.TTError(f"{name}: expected a CUDA tensor — libtriplane_b200 has no CPU path")
    if t.dtype != torch.float32:
        raise _cabi.TTError(f"{name}: expected float32, got {t.dtype}")
    return t.contiguous()


def _cfg(C_: int, R: int, P: int, rays_per_cache: int, s: PathScalars, flags: int = 0) -> _cabi.TTConfig:
    return _cabi.TTConfig(C_, R, P, max(int(rays_per_cache), 1), s.radius, s.sdf_bias_radius, s.inv_std,
                          s.cos_anneal_ratio, s.near_plane, s.far_plane, s.render_step_size, flags, int(s.image_h),
                          int(s.image_w))


def set_option(name: str, value: int):
    """Experiment switches of the backward (include/triplane_b200.h tt_set_option): "scatter", "patch_lists"."""
    L = _lib()
    _cabi.check(L, L.tt_set_option(name.encode(), int(value)), "tt_set_option")


def set_impl(impl: int):
    """1 = tcgen05 tensor-core kernels (default), 0 = SIMT fp32 reference kernels.  Both are CUDA."""
    L = _lib()
    _cabi.check(L, L.tt_set_impl(int(impl)), "tt_set_impl")


def get_impl() -> int:
    return int(_lib().tt_get_impl())


def launch_count() -> int:
    return int(_lib().tt_launch_count())


def profile_begin():
    _lib().tt_profile_begin()


def profile_end():
    """[(kernel name, milliseconds)] for every launch since profile_begin()."""
    buf = C.create_string_buffer(1 << 20)
    _lib().tt_profile_end(buf, len(buf))
    recs = []
    for item in buf.value.decode().split(";"):
        if item:
            name, ms = item.rsplit(":", 1)
            recs.append((name.split("<")[0].strip("( "), float(ms)))
    return recs


# ------------------------------------------------------------------------------------------------ raw ops
def repack_planes(space_cache: Tensor, C_: Optional[int] = None, off_geo: int = 0, off_tex: int = 0) -> Tensor:
    """NCHW space cache [P,6,Csrc,R,R] -> channel-last rotated planes [P,6,R,R,C] (tt_repack_planes)."""
    sc = _need(space_cache, "space_cache")
    if sc.dim() != 5 or sc.shape[1] != 6 or sc.shape[3] != sc.shape[4]:
        raise _cabi.TTError(f"space_cache must be [P,6,C,R,R], got {tuple(sc.shape)}")
    P, _, Csrc, R, _ = sc.shape
    C_ = C_ or Csrc
    out = torch.empty((P, 6, R, R, C_), device=sc.device, dtype=torch.float32)
    L = _lib()
    with torch.cuda.device(sc.device):
        _cabi.check(L, L.tt_repack_planes(_ptr(sc), P, Csrc, off_geo, off_tex, C_, R, _ptr(out), _stream(sc.device)),
                    "tt_repack_planes")
    return out


def repack_planes_bwd(gplanes: Tensor, Csrc: Optional[int] = None) -> Tensor:
    """Channel-last rotated gradient [P,6,R,R,C] -> gradient of the NCHW space cache [P,6,Csrc,R,R].  Csrc == 2C: the cache
    was the VAE decoder's raw output and the channel split was folded into the repack (geometry planes own channels [0,C),
    texture planes [C,2C)); the unused halves are zero like the gradient of the reference's masked gather."""
    g = _need(gplanes, "gplanes")
    P, _, R, _, C_ = g.shape
    Csrc = Csrc or C_
    L = _lib()
    with torch.cuda.device(g.device):
        if Csrc == C_:
            out = torch.empty((P, 6, C_, R, R), device=g.device, dtype=torch.float32)
            _cabi.check(L, L.tt_repack_planes_bwd(_ptr(g), P, C_, R, _ptr(out), _stream(g.device)), "tt_repack_planes_bwd")
        else:
            off_geo, off_tex = split_offsets(Csrc, C_)
            out = torch.zeros((P, 6, Csrc, R, R), device=g.device, dtype=torch.float32)
            _cabi.check(L, L.tt_repack_planes_bwd_split(_ptr(g), P, Csrc, off_geo, off_tex, C_, R, _ptr(out), _stream(g.device)),
                        "tt_repack_planes_bwd_split")
    return out


def split_offsets(Csrc: int, C_: int) -> Tuple[int, int]:
    """Channel offsets of the geometry / texture planes inside a space cache with Csrc channels (split_channels v1,
    few_step…diffusion.py:186-196: geometry = first half of planes 0-2, texture = second half of planes 3-5)."""
    if Csrc == C_:
        return 0, 0
    if Csrc == 2 * C_:
        return 0, C_
    raise _cabi.TTError(f"space cache has {Csrc} channels; the decoders expect {C_} (or the un-split {2 * C_})")


def pack_weights(sdf: Sequence[Tensor], feature: Optional[Sequence[Tensor]], deformation: Optional[Sequence[Tensor]],
                 C_: int) -> Tensor:
    """nn.Linear weights of the three VanillaMLPs -> the packed buffer the kernels read (tt_pack_weights)."""
    if C_ not in SUPPORTED_C:
        raise _cabi.TTError(f"unsupported channel count {C_}; compiled for {SUPPORTED_C}")

    def chk(ws, in_dim, out_dim, name):
        ws = [_need(w.detach(), name) for w in ws]
        want = [(HIDDEN, in_dim), (HIDDEN, HIDDEN), (out_dim, HIDDEN)]
        if [tuple(w.shape) for w in ws] != want:
            raise _cabi.TTError(f"{name} MLP must be {want} (n_neurons=64, n_hidden_layers=2, bias-free), got "
                                f"{[tuple(w.shape) for w in ws]}")
        return ws
    s = chk(sdf, C_, 1, "sdf_network")
    f = chk(feature, 3 * C_, 3, "feature_network") if feature is not None else [None] * 3
    d = chk(deformation, C_, 3, "deformation_network") if deformation is not None else [None] * 3
    L = _lib()
    dev = s[0].device
    wp = torch.zeros(L.tt_wpack_floats(C_), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _cabi.check(L, L.tt_pack_weights(*[_ptr(w) for w in s + f + d], C_, _ptr(wp), _stream(dev)), "tt_pack_weights")
    return wp


def split_wgrad(gw: Tensor, C_: int) -> List[Tensor]:
    L = _lib()
    off = (C.c_int64 * 6)()
    _cabi.check(L, L.tt_wgrad_offsets(C_, off), "tt_wgrad_offsets")
    shapes = [(HIDDEN, C_), (HIDDEN, HIDDEN), (1, HIDDEN), (HIDDEN, 3 * C_), (HIDDEN, HIDDEN), (3, HIDDEN)]
    return [gw[off[i]:off[i] + a * b].view(a, b) for i, (a, b) in enumerate(shapes)]


def geometry_fwd(planes: Tensor, wpack: Tensor, s: PathScalars, points: Optional[Tensor], grid_res: int = 0,
                 want=("sdf",)) -> dict:
    """tt_geometry_fwd.  points [P,M,3] or None (+grid_res).  want ⊂ {sdf, sdf_orig, features, normal, sdf_grad,
    deformation}."""
    P, _, R, _, C_ = planes.shape
    dev = planes.device
    if points is not None:
        points = _need(points, "points")
        if points.dim() != 3 or points.shape[0] != P or points.shape[2] != 3:
            raise _cabi.TTError(f"points must be [P={P},M,3], got {tuple(points.shape)}")
        M = points.shape[1]
    else:
        M = grid_res ** 3
    N = P * M
    out = {}
    for k in want:
        out[k] = torch.empty((N,) if k in ("sdf", "sdf_orig") else (N, 3), device=dev, dtype=torch.float32)
    cfg = _cfg(C_, R, P, 1, s)
    L = _lib()
    with torch.cuda.device(dev):
        _cabi.check(L, L.tt_geometry_fwd(_ptr(planes), _ptr(wpack), C.byref(cfg), _ptr(points), M, grid_res,
                                         *[_ptr(out.get(k)) for k in ("sdf", "sdf_orig", "features", "normal",
                                                                      "sdf_grad", "deformation")],
                                         _stream(dev)), "tt_geometry_fwd")
    return out


def geometry_bwd(planes: Tensor, wpack: Tensor, s: PathScalars, points: Tensor, g_sdf, g_features, g_normal,
                 g_sdf_grad, need_planes=True, need_w=True) -> Tuple[Optional[Tensor], Optional[Tensor]]:
    P, _, R, _, C_ = planes.shape
    dev = planes.device
    points = _need(points, "points")
    M = points.shape[1]
    L = _lib()
    cfg = _cfg(C_, R, P, 1, s)
    scratch = torch.empty(L.tt_geometry_bwd_scratch_floats(C.byref(cfg), P * M), device=dev, dtype=torch.float32)
    gplanes = torch.zeros_like(planes) if need_planes else None
    gw = torch.zeros(L.tt_wgrad_floats(C_), device=dev, dtype=torch.float32) if need_w else None
    gs = [None if g is None else _need(g, "grad") for g in (g_sdf, g_features, g_normal, g_sdf_grad)]
    with torch.cuda.device(dev):
        _cabi.check(L, L.tt_geometry_bwd(_ptr(planes), _ptr(wpack), C.byref(cfg), _ptr(points), M,
                                         *[_ptr(g) for g in gs], _ptr(scratch), _ptr(gplanes), _ptr(gw),
                                         _stream(dev)), "tt_geometry_bwd")
    return gplanes, gw


def importance_sample(planes: Tensor, wpack: Tensor, s: PathScalars, rays_o: Tensor, rays_d: Tensor,
                      rays_per_cache: int, n_imp: int, n_fine: int, jitter0: Optional[Tensor] = None,
                      jitter1: Optional[Tensor] = None) -> Tensor:
    """tt_importance_sample -> sorted edges t_vals [n_rays, n_imp+n_fine+2]."""
    P, _, R, _, C_ = planes.shape
    dev = planes.device
    o, d = _need(rays_o, "rays_o").view(-1, 3), _need(rays_d, "rays_d").view(-1, 3)
    n = o.shape[0]
    L = _lib()
    scratch = torch.empty(L.tt_sample_scratch_floats(n, n_imp), device=dev, dtype=torch.float32)
    t_vals = torch.empty((n, n_imp + n_fine + 2), device=dev, dtype=torch.float32)
    j0 = None if jitter0 is None else _need(jitter0, "jitter0")
    j1 = None if jitter1 is None else _need(jitter1, "jitter1")
    cfg = _cfg(C_, R, P, rays_per_cache, s)
    with torch.cuda.device(dev):
        _cabi.check(L, L.tt_importance_sample(_ptr(planes), _ptr(wpack), C.byref(cfg), _ptr(o), _ptr(d), n, n_imp,
                                              n_fine, _ptr(j0), _ptr(j1), _ptr(scratch), _ptr(t_vals),
                                              _stream(dev)), "tt_importance_sample")
    return t_vals


def _intervals(t_starts: Tensor, t_ends: Tensor):
    """Accept [n,S] tensors that are contiguous or row-strided views of one edge buffer."""
    if t_starts.stride(-1) != 1 or t_ends.stride(-1) != 1 or t_starts.stride(0) != t_ends.stride(0) \
            or t_starts.dtype != torch.float32 or not t_starts.is_cuda:
        t_starts, t_ends = _need(t_starts, "t_starts"), _need(t_ends, "t_ends")
    return t_starts, t_ends, t_starts.stride(0), t_starts.shape[1]


def render_fwd(planes, wpack, s: PathScalars, rays_o, rays_d, rays_per_cache, t_starts, t_ends,
               save_for_backward: bool, extras: bool) -> dict:
    P, _, R, _, C_ = planes.shape
    dev = planes.device
    o, d = _need(rays_o, "rays_o").view(-1, 3), _need(rays_d, "rays_d").view(-1, 3)
    n = o.shape[0]
    t0, t1, stride, S = _intervals(t_starts, t_ends)
    if t0.shape[0] != n:
        raise _cabi.TTError(f"t_starts has {t0.shape[0]} rows for {n} rays")
    N = n * S
    out = {"acc": torch.empty((n, ACC), device=dev, dtype=torch.float32)}
    names = []
    if save_for_backward or extras:
        names += ["sdf", "sdf_grad", "features", "trans"]
    if extras:
        names += ["sdf_orig", "normal", "weights"]
    for k in names:
        out[k] = torch.empty((N,) if k in ("sdf", "sdf_orig", "weights", "trans") else (N, 3), device=dev,
                             dtype=torch.float32)
    if save_for_backward:
        out["tex_masks"] = torch.empty((N, 4), device=dev, dtype=torch.int64)
    cfg = _cfg(C_, R, P, rays_per_cache, s, 1 if extras else 0)      # TT_FLAG_ALL_FEATURES
    L = _lib()
    scratch = torch.empty(L.tt_render_fwd_scratch_floats(n, S), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _cabi.check(L, L.tt_render_fwd(_ptr(planes), _ptr(wpack), C.byref(cfg), _ptr(o), _ptr(d), n, _ptr(t0),
                                       _ptr(t1), stride, S, _ptr(out["acc"]),
                                       *[_ptr(out.get(k)) for k in ("sdf", "sdf_orig", "sdf_grad", "normal",
                                                                    "features", "weights", "trans",
                                                                    "tex_masks")],
                                       _ptr(scratch), _stream(dev)), "tt_render_fwd")
    return out


PRECISE_BACKWARD = False      # TT_FLAG_PRECISE_BWD for every backward of this process (see set_precise_backward)


def set_precise_backward(on: bool):
    """Run the colour decoder's backward layers as 3xTF32 (fp32-equivalent gradients of the feature network)
    instead of single-pass TF32.  Slower: one group per CTA fits (config 3: 651 vs 432 ms for that kernel)."""
    global PRECISE_BACKWARD
    PRECISE_BACKWARD = bool(on)


def render_bwd(planes, wpack, s: PathScalars, rays_o, rays_d, rays_per_cache, t_starts, t_ends, saved: dict,
               g_acc, g_sdf=None, g_sdf_grad=None, g_normal=None, g_features=None, g_weights=None,
               rgb_grad_scale: float = 1.0, need_planes=True, need_w=True, need_inv_std=False):
    P, _, R, _, C_ = planes.shape
    dev = planes.device
    o, d = _need(rays_o, "rays_o").view(-1, 3), _need(rays_d, "rays_d").view(-1, 3)
    n = o.shape[0]
    t0, t1, stride, S = _intervals(t_starts, t_ends)
    L = _lib()
    cfg = _cfg(C_, R, P, rays_per_cache, s, 2 if PRECISE_BACKWARD else 0)
    scratch = torch.empty(L.tt_render_bwd_scratch_floats(C.byref(cfg), n, S), device=dev, dtype=torch.float32)
    gplanes = torch.zeros_like(planes) if need_planes else None
    gw = torch.zeros(L.tt_wgrad_floats(C_), device=dev, dtype=torch.float32) if need_w else None
    gis = torch.zeros(1, device=dev, dtype=torch.float32) if need_inv_std else None
    opt = [None if g is None else _need(g, "grad") for g in (g_sdf, g_sdf_grad, g_normal, g_features, g_weights)]
    with torch.cuda.device(dev):
        _cabi.check(L, L.tt_render_bwd(_ptr(planes), _ptr(wpack), C.byref(cfg), _ptr(o), _ptr(d), n, _ptr(t0),
                                       _ptr(t1), stride, S, _ptr(saved["acc"]), _ptr(saved["sdf"]),
                                       _ptr(saved["sdf_grad"]), _ptr(saved["features"]), _ptr(saved["trans"]),
                                       _ptr(saved.get("tex_masks")), _ptr(_need(g_acc, "g_acc")), *[_ptr(g) for g in opt], float(rgb_grad_scale),
                                       _ptr(scratch), _ptr(gplanes), _ptr(gw), _ptr(gis), _stream(dev)),
                    "tt_render_bwd")
    return gplanes, gw, gis


def composite_fwd(alphas: Tensor, values: Optional[Tensor]):
    a = _need(alphas, "alphas")
    n, S = a.shape
    D = 0 if values is None else values.shape[-1]
    v = None if values is None else _need(values, "values")
    w = torch.empty_like(a)
    T = torch.empty_like(a)
    out = torch.empty((n, max(D, 1)), device=a.device, dtype=torch.float32)
    L = _lib()
    with torch.cuda.device(a.device):
        _cabi.check(L, L.tt_composite_fwd(_ptr(a), _ptr(v), n, S, D, _ptr(w), _ptr(T), _ptr(out), _stream(a.device)),
                    "tt_composite_fwd")
    return w, T, out


def composite_bwd(alphas, values, trans, g_out, g_weights):
    a = _need(alphas, "alphas")
    n, S = a.shape
    D = 0 if values is None else values.shape[-1]
    v = None if values is None else _need(values, "values")
    ga = torch.empty_like(a)
    gv = None if values is None else torch.empty_like(v)
    go = None if g_out is None else _need(g_out, "g_out")
    gwt = None if g_weights is None else _need(g_weights, "g_weights")
    L = _lib()
    with torch.cuda.device(a.device):
        _cabi.check(L, L.tt_composite_bwd(_ptr(a), _ptr(v), _ptr(_need(trans, "trans")), _ptr(go), _ptr(gwt), n, S, D,
                                          _ptr(ga), _ptr(gv), _stream(a.device)), "tt_composite_bwd")
    return ga, gv


# ------------------------------------------------------------------------------------------------ autograd
class _PlaneCache:
    """Repacked planes / packed weights are reused while their source tensors are unchanged (the renderer is
    entered twice per step with the same space cache: global + patch views, patch_renderer.py:49-72)."""

    def __init__(self):
        self.refs = None
        self.key = None
        self.val = None

    def get(self, tensors: Sequence[Tensor], build):
        # identity of the live tensor objects + their in-place version counters: a freed tensor whose storage
        # address is recycled can never alias a cache entry
        key = tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in tensors)
        alive = self.refs is not None and len(self.refs) == len(tensors) and \
            all(r() is t for r, t in zip(self.refs, tensors))
        if not alive or key != self.key:
            self.val = build()
            self.key = key
            self.refs = [weakref.ref(t) for t in tensors]
        return self.val


_planes_cache = _PlaneCache()
_weights_caches = {}      # one slot per weight-set shape (with / without the deformation decoder): no thrashing when
                          # the renderer (6 tensors) and the field query (9 tensors) alternate inside a step


def cached_planes(space_cache: Tensor, C_: Optional[int] = None) -> Tensor:
    """Repacked planes of a space cache [P,6,C,R,R] -- or of the VAE decoder's raw output [P,6,2C,R,R] (pass C_): the
    channel split of ``decode`` then happens inside the one repack pass instead of a masked gather + copy."""
    Csrc = space_cache.shape[2]
    C_ = C_ or Csrc
    off_geo, off_tex = split_offsets(Csrc, C_)
    return _planes_cache.get([space_cache], lambda: repack_planes(space_cache.detach(), C_, off_geo, off_tex))


def cached_wpack(sdf_w, feat_w, def_w, C_) -> Tensor:
    ws = list(sdf_w) + list(feat_w or []) + list(def_w or [])
    slot = _weights_caches.setdefault((len(ws), C_), _PlaneCache())
    return slot.get(ws, lambda: pack_weights(sdf_w, feat_w, def_w, C_))


def clear_caches():
    """Drop the cached repacked planes / packed weights (they pin device memory until the next call replaces them)."""
    _planes_cache.refs = _planes_cache.key = _planes_cache.val = None
    _weights_caches.clear()


class _impl_scope:
    """Run a backward with the kernel family that produced its forward's saved state (the tensor-core forward saves
    ReLU masks the SIMT forward never writes): ``set_impl`` between a forward and its backward must not mix them."""

    def __init__(self, impl: int):
        self.want, self.prev = int(impl), None

    def __enter__(self):
        cur = get_impl()
        if cur != self.want:
            self.prev = cur
            set_impl(self.want)

    def __exit__(self, *exc):
        if self.prev is not None:
            set_impl(self.prev)


class RenderFunction(torch.autograd.Function):
    """Fused march (tt_render_fwd / tt_render_bwd) as one autograd node.

    inputs : space_cache [P,6,C,R,R] (NCHW, reference layout), 6 decoder weights, inv_std (0-dim tensor),
             rays_o/d [Nr,3], t_starts/t_ends [Nr,S]
    outputs: acc [Nr,10] and, when ``extras``, sdf, sdf_orig [N,1]; sdf_grad, normal, features [N,3]; weights [N,1]
    """

    @staticmethod
    def forward(ctx, space_cache, ws0, ws1, ws2, wf0, wf1, wf2, inv_std, rays_o, rays_d, t_starts, t_ends,
                scalars: PathScalars, rays_per_cache: int, rgb_grad_scale: float, extras: bool):
        C_ = ws0.shape[1]                   # plane channels = input width of the SDF decoder
        ctx.Csrc = space_cache.shape[2]     # == C_, or 2 C_ for an un-split cache (channel split folded into the repack)
        planes = cached_planes(space_cache, C_)
        wpack = cached_wpack([ws0, ws1, ws2], [wf0, wf1, wf2], None, C_)
        need_grad = any(ctx.needs_input_grad[:8])
        out = render_fwd(planes, wpack, scalars, rays_o, rays_d, rays_per_cache, t_starts, t_ends, need_grad, extras)
        ctx.scalars, ctx.rays_per_cache, ctx.rgb_grad_scale = scalars, rays_per_cache, rgb_grad_scale
        ctx.C, ctx.impl = C_, get_impl()
        if need_grad:
            ctx.save_for_backward(planes, wpack, rays_o, rays_d, t_starts, t_ends, out["acc"], out["sdf"],
                                  out["sdf_grad"], out["features"], out["trans"], out["tex_masks"])
        if not extras:
            return (out["acc"],)
        res = (out["acc"], out["sdf"].view(-1, 1), out["sdf_orig"].view(-1, 1), out["sdf_grad"], out["normal"],
               out["features"], out["weights"].view(-1, 1))
        return res

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_acc, g_sdf=None, g_sdf_orig=None, g_sdf_grad=None, g_normal=None, g_features=None,
                 g_weights=None):
        planes, wpack, rays_o, rays_d, t0, t1, acc, sdf, sdf_grad, features, trans, tex_masks = ctx.saved_tensors
        if g_sdf is not None and g_sdf_orig is not None:
            g_sdf = g_sdf + g_sdf_orig
        elif g_sdf is None:
            g_sdf = g_sdf_orig
        need_planes = ctx.needs_input_grad[0]
        need_w = any(ctx.needs_input_grad[1:7])
        with _impl_scope(ctx.impl):
            gplanes, gw, gis = render_bwd(planes, wpack, ctx.scalars, rays_o, rays_d, ctx.rays_per_cache, t0, t1,
                                          {"acc": acc, "sdf": sdf, "sdf_grad": sdf_grad, "features": features,
                                           "trans": trans, "tex_masks": tex_masks},
                                          g_acc, None if g_sdf is None else g_sdf.reshape(-1), g_sdf_grad, g_normal,
                                          g_features, None if g_weights is None else g_weights.reshape(-1),
                                          ctx.rgb_grad_scale, need_planes, need_w, ctx.needs_input_grad[7])
        g_sc = repack_planes_bwd(gplanes, ctx.Csrc) if need_planes else None
        gws = split_wgrad(gw, ctx.C) if need_w else [None] * 6
        gws = [g if ctx.needs_input_grad[1 + i] else None for i, g in enumerate(gws)]
        g_inv = gis.reshape(()) if gis is not None else None
        return (g_sc, *gws, g_inv, None, None, None, None, None, None, None, None)


class GeometryFunction(torch.autograd.Function):
    """geometry.forward on a point list (tt_geometry_fwd / tt_geometry_bwd).

    outputs: sdf, sdf_orig [N,1], features [N,3], normal, sdf_grad [N,3] (the last two zero-size if not requested)
    """

    @staticmethod
    def forward(ctx, space_cache, ws0, ws1, ws2, wf0, wf1, wf2, points, scalars: PathScalars, output_normal: bool):
        C_ = ws0.shape[1]
        ctx.Csrc = space_cache.shape[2]
        planes = cached_planes(space_cache, C_)
        wpack = cached_wpack([ws0, ws1, ws2], [wf0, wf1, wf2], None, C_)
        want = ["sdf", "sdf_orig", "features"] + (["normal", "sdf_grad"] if output_normal else [])
        out = geometry_fwd(planes, wpack, scalars, points, 0, want)
        ctx.scalars, ctx.C, ctx.output_normal, ctx.impl = scalars, C_, output_normal, get_impl()
        ctx.save_for_backward(planes, wpack, points)
        empty = torch.empty((0, 3), device=planes.device)
        return (out["sdf"].view(-1, 1), out["sdf_orig"].view(-1, 1), out["features"],
                out.get("normal", empty), out.get("sdf_grad", empty))

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_sdf, g_sdf_orig, g_features, g_normal, g_sdf_grad):
        planes, wpack, points = ctx.saved_tensors
        g = None
        for x in (g_sdf, g_sdf_orig):
            if x is not None:
                g = x.reshape(-1) if g is None else g + x.reshape(-1)
        if not ctx.output_normal:
            g_normal = g_sdf_grad = None
        need_planes = ctx.needs_input_grad[0]
        need_w = any(ctx.needs_input_grad[1:7])
        with _impl_scope(ctx.impl):
            gplanes, gw = geometry_bwd(planes, wpack, ctx.scalars, points, g, g_features, g_normal, g_sdf_grad,
                                       need_planes, need_w)
        g_sc = repack_planes_bwd(gplanes, ctx.Csrc) if need_planes else None
        gws = split_wgrad(gw, ctx.C) if need_w else [None] * 6
        gws = [x if ctx.needs_input_grad[1 + i] else None for i, x in enumerate(gws)]
        return (g_sc, *gws, None, None, None)


def field_bwd(planes: Tensor, wpack: Tensor, s: PathScalars, points: Tensor, g_sdf, g_deformation, need_planes=True,
              need_w=True, need_wd=True):
    """tt_field_bwd: backward of forward_field (sdf + deformation decoders) -> (gplanes, gw, gw_def)."""
    P, _, R, _, C_ = planes.shape
    dev = planes.device
    points = _need(points, "points")
    M = points.shape[1]
    L = _lib()
    cfg = _cfg(C_, R, P, 1, s)
    scratch = torch.empty(L.tt_geometry_bwd_scratch_floats(C.byref(cfg), P * M), device=dev, dtype=torch.float32)
    gplanes = torch.zeros_like(planes) if need_planes else None
    gw = torch.zeros(L.tt_wgrad_floats(C_), device=dev, dtype=torch.float32) if need_w else None
    gwd = torch.zeros(L.tt_wgrad_def_floats(C_), device=dev, dtype=torch.float32) if need_wd else None
    gs = None if g_sdf is None else _need(g_sdf, "g_sdf")
    gd = None if g_deformation is None else _need(g_deformation, "g_deformation")
    with torch.cuda.device(dev):
        _cabi.check(L, L.tt_field_bwd(_ptr(planes), _ptr(wpack), C.byref(cfg), _ptr(points), M, _ptr(gs), _ptr(gd),
                                      _ptr(scratch), _ptr(gplanes), _ptr(gw), _ptr(gwd), _stream(dev)), "tt_field_bwd")
    return gplanes, gw, gwd


class FieldFunction(torch.autograd.Function):
    """geometry.forward_field on a point list: sdf [N,1] and deformation [N,3] (tt_geometry_fwd / tt_field_bwd),
    differentiable w.r.t. the space cache, the SDF decoder and the deformation decoder (the mesh renderer trains
    through it, generative_space_mesh_rasterize_renderer.py:449-452)."""

    @staticmethod
    def forward(ctx, space_cache, ws0, ws1, ws2, wf0, wf1, wf2, wd0, wd1, wd2, points, scalars: PathScalars):
        C_ = ws0.shape[1]
        ctx.Csrc = space_cache.shape[2]
        planes = cached_planes(space_cache, C_)
        wpack = cached_wpack([ws0, ws1, ws2], [wf0, wf1, wf2], [wd0, wd1, wd2], C_)
        out = geometry_fwd(planes, wpack, scalars, points, 0, ["sdf", "deformation"])
        ctx.scalars, ctx.C, ctx.impl = scalars, C_, get_impl()
        ctx.save_for_backward(planes, wpack, points)
        return out["sdf"].view(-1, 1), out["deformation"]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_sdf, g_def):
        planes, wpack, points = ctx.saved_tensors
        need_planes = ctx.needs_input_grad[0]
        need_w, need_wd = any(ctx.needs_input_grad[1:4]), any(ctx.needs_input_grad[7:10])
        with _impl_scope(ctx.impl):
            gplanes, gw, gwd = field_bwd(planes, wpack, ctx.scalars, points,
                                         None if g_sdf is None else g_sdf.reshape(-1), g_def, need_planes, need_w,
                                         need_wd)
        g_sc = repack_planes_bwd(gplanes, ctx.Csrc) if need_planes else None
        gws = split_wgrad(gw, ctx.C)[:3] if need_w else [None] * 3
        gws = [x if ctx.needs_input_grad[1 + i] else None for i, x in enumerate(gws)]
        if need_wd:
            C_ = ctx.C
            gds = [gwd[:HIDDEN * C_].view(HIDDEN, C_), gwd[HIDDEN * C_:HIDDEN * C_ + HIDDEN * HIDDEN].view(HIDDEN, HIDDEN),
                   gwd[HIDDEN * C_ + HIDDEN * HIDDEN:].view(3, HIDDEN)]
            gds = [x if ctx.needs_input_grad[7 + i] else None for i, x in enumerate(gds)]
        else:
            gds = [None] * 3
        return (g_sc, *gws, None, None, None, *gds, None, None)


class RepackFunction(torch.autograd.Function):
    """space cache [P,6,C,R,R] (NCHW) -> channel-last rotated planes [P,6,R,R,C]; a permutation, so the gradient is
    tt_repack_planes_bwd.  ``off_geo/off_tex`` fold the channel split of ``decode`` (few_step…diffusion.py:180-196)."""

    @staticmethod
    def forward(ctx, space_cache, C_, off_geo, off_tex):
        ctx.shape, ctx.off = tuple(space_cache.shape), (off_geo, off_tex)
        return repack_planes(space_cache, C_, off_geo, off_tex)

    @staticmethod
    def backward(ctx, g):
        return repack_planes_bwd(g.contiguous(), ctx.shape[2]), None, None, None


class CompositeFunction(torch.autograd.Function):
    """weights = alpha * exclusive_prod(1 - alpha) and Σ weights * values per ray (tt_composite_*)."""

    @staticmethod
    def forward(ctx, alphas, values):
        w, T, out = composite_fwd(alphas, values)
        ctx.save_for_backward(alphas, values if values is not None else torch.empty(0, device=alphas.device), T)
        ctx.has_values = values is not None
        return w, T, out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_w, g_T, g_out):
        alphas, values, T = ctx.saved_tensors
        ga, gv = composite_bwd(alphas, values if ctx.has_values else None, T, g_out, g_w)
        return ga, gv
