"""Pure-PyTorch restatement of the reference's triplane volume-rendering path.

TEST INFRASTRUCTURE (see ``oracle/__init__.py`` for the pinning status).

Every function cites the reference ``file:line`` it follows.  Abbreviations:
  GEO  = custom/triplaneturbo/models/geometry/few_step_triplane_dual_stable_diffusion.py
  GUT  = custom/triplaneturbo/models/geometry/utils.py
  REN  = custom/triplaneturbo/models/renderers/generative_space_sdf_volume_renderer.py
  NEUS = threestudio/models/renderers/neus_volume_renderer.py
  EST  = threestudio/models/estimators.py
  OPS  = threestudio/utils/ops.py
  NET  = threestudio/models/networks.py
  MAT  = threestudio/models/materials/no_material.py
  PATCH= threestudio/models/renderers/patch_renderer.py

The data layout is the reference's: space cache ``[B, 6, C, R, R]`` NCHW fp32
(planes 0-2 geometry, 3-5 texture), decoder weights as ``nn.Linear.weight``
(``[out, in]``, no bias).  Shipped configuration only
(``configs/TriplaneTurbo_v1.yaml:73-150``): ``rotate_planes: v1``,
``geo_interpolate: v1`` (sum), ``tex_interpolate: v2`` (concat), ``sdf_bias:
sphere 0.5``, ``normal_type: analytic``, ``estimator: importance``,
``use_volsdf: false``, ``color_activation: sigmoid-mipnerf``.
"""
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

from . import nerfacc_restated as nerfacc
from .bilinear import grid_sample_2d_manual


def grid_sample(input: Tensor, grid: Tensor) -> Tensor:
    """GUT:21-24: the double-differentiable op iff ``grid.requires_grad`` (the reference's gradfix extension,
    here :func:`oracle.bilinear.grid_sample_2d_manual`), ATen ``F.grid_sample`` otherwise."""
    if grid.requires_grad:
        return grid_sample_2d_manual(input, grid)
    return F.grid_sample(input=input, grid=grid, mode="bilinear", padding_mode="zeros", align_corners=False)


# --------------------------------------------------------------------------- config
@dataclass
class PathConfig:
    radius: float = 1.0                       # yaml :75,131
    sdf_bias_radius: float = 0.5              # yaml :78-79
    num_samples_per_ray: int = 64             # yaml :142
    num_samples_per_ray_importance: int = 128  # yaml :143
    near_plane: float = 0.1                   # yaml :145
    far_plane: float = 4.0                    # yaml :146
    learned_variance_init: float = 0.4605     # yaml :137
    cos_anneal_ratio: float = 1.0             # NEUS:91
    rgb_grad_shrink: float = 1.0              # REN:397-400
    normal_direction: str = "camera"          # REN:71
    rotate_planes: str = "v1"                 # yaml :81

    @property
    def render_step_size(self) -> float:      # NEUS:84-86
        return 1.732 * 2 * self.radius / self.num_samples_per_ray

    @property
    def n_intervals(self) -> int:             # EST:95-99 -> (N_imp+1) + (N+1) edges - 1
        return self.num_samples_per_ray_importance + self.num_samples_per_ray + 1


# --------------------------------------------------------------------------- small ops
def scale_tensor(dat: Tensor, inp_scale, tgt_scale) -> Tensor:
    """OPS:27-38 (operation order kept: subtract, divide, multiply, add)."""
    dat = (dat - inp_scale[0]) / (inp_scale[1] - inp_scale[0])
    dat = dat * (tgt_scale[1] - tgt_scale[0]) + tgt_scale[0]
    return dat


def bbox_of(radius: float, like: Tensor) -> Tensor:
    """threestudio/models/geometry/base.py:71-81: ``[[-r,-r,-r],[r,r,r]]``."""
    return torch.tensor([[-radius] * 3, [radius] * 3], dtype=like.dtype, device=like.device)


def rescale_points(points: Tensor, radius: float) -> Tensor:
    """GEO:261-271 -> GUT:31-43 with ``unbounded=False`` -> ``scale_tensor(x, bbox, (-1, 1))``."""
    return scale_tensor(points, bbox_of(radius, points), (-1, 1))


def sigmoid_mipnerf(x: Tensor) -> Tensor:
    """OPS:118-119."""
    return torch.sigmoid(x) * (1 + 2 * 0.001) - 0.001


def vanilla_mlp(x: Tensor, weights: List[Tensor]) -> Tensor:
    """NET:67-104: Linear(no bias) -> ReLU -> Linear -> ReLU -> Linear, fp32, no output activation."""
    h = x
    for i, w in enumerate(weights):
        h = F.linear(h, w)
        if i + 1 < len(weights):
            h = torch.relu(h)
    return h


def decode_split_channels(triplane: Tensor) -> Tensor:
    """GEO:180-196 (``split_channels == 'v1'``): ``[B,6,2C,H,W]`` -> ``[B,6,C,H,W]``.

    Planes 0-2 keep channels ``[0, C)``, planes 3-5 keep ``[C, 2C)``.
    """
    B, _, C2, H, W = triplane.shape
    C = C2 // 2
    geo = torch.tensor([True] * C + [False] * C)
    tex = torch.tensor([False] * C + [True] * C)
    used = torch.stack([geo] * 3 + [tex] * 3, dim=0).to(triplane.device)
    return triplane[:, used].view(B, 6, C, H, W)


def rotate_planes(space_cache: Tensor, mode: str = "v1") -> Tensor:
    """GEO:212-239."""
    rot = torch.zeros_like(space_cache)
    if mode == "v1":
        rot[:, 0::3] = torch.transpose(space_cache[:, 0::3], 3, 4)
    elif mode == "v2":
        rot[:, 0::3] = torch.flip(space_cache[:, 0::3], dims=(4,))
    else:
        raise NotImplementedError(mode)
    rot[:, 1::3] = torch.rot90(space_cache[:, 1::3], k=2, dims=(3, 4))
    rot[:, 2::3] = torch.rot90(space_cache[:, 2::3], k=-1, dims=(3, 4))
    return rot


_PLANES = torch.tensor(
    [[[1, 0, 0], [0, 1, 0], [0, 0, 1]],
     [[1, 0, 0], [0, 0, 1], [0, 1, 0]],
     [[0, 0, 1], [0, 1, 0], [1, 0, 0]]], dtype=torch.float32)  # GUT:46-63


def project_onto_planes(planes: Tensor, coordinates: Tensor) -> Tensor:
    """GUT:111-125."""
    N, M, _ = coordinates.shape
    n_planes = planes.shape[0]
    coordinates = coordinates.unsqueeze(1).expand(-1, n_planes, -1, -1).reshape(N * n_planes, M, 3)
    inv_planes = torch.linalg.inv(planes).unsqueeze(0).expand(N, -1, -1, -1).reshape(N * n_planes, 3, 3)
    return torch.bmm(coordinates, inv_planes)[..., :2]


def sample_from_planes(plane_features: Tensor, coordinates: Tensor, interpolate_feat: str = "v1",
                       box_warp: float = 2.0, sampler: Callable = grid_sample) -> Tensor:
    """GUT:127-145. plane_features [N,3,C,H,W], coordinates [N,M,3] -> [N,M,C] (v1) / [N,M,3C] (v2)."""
    N, n_planes, C, H, W = plane_features.shape
    _, M, _ = coordinates.shape
    plane_features = plane_features.reshape(N * n_planes, C, H, W)
    coordinates = (2 / box_warp) * coordinates
    projected = project_onto_planes(_PLANES.to(coordinates), coordinates).unsqueeze(1)
    out = sampler(plane_features, projected)
    out = out.permute(0, 3, 2, 1).reshape(N, n_planes, M, C)
    if interpolate_feat in (None, "v1"):
        return out.sum(dim=1, keepdim=True).reshape(N, M, C)
    if interpolate_feat == "v2":
        return out.permute(0, 2, 1, 3).reshape(N, M, n_planes * C)
    raise NotImplementedError(interpolate_feat)


def interpolate_encodings(points: Tensor, space_cache: Tensor, only_geo: bool = False,
                          rotate: str = "v1", sampler: Callable = grid_sample):
    """GEO:198-258 with ``geo_interpolate=v1``, ``tex_interpolate=v2``."""
    rot = rotate_planes(space_cache, rotate)
    geo = sample_from_planes(rot[:, 0:3].contiguous(), points, "v1", sampler=sampler)
    if only_geo:
        return geo
    tex = sample_from_planes(rot[:, 3:6].contiguous(), points, "v2", sampler=sampler)
    return geo, tex


def get_shifted_sdf(points_unscaled: Tensor, sdf: Tensor, sdf_bias_radius: float) -> Tensor:
    """GEO:131-154, ``sdf_bias == 'sphere'``."""
    return sdf + ((points_unscaled ** 2).sum(dim=-1, keepdim=True).sqrt() - sdf_bias_radius)


# --------------------------------------------------------------------------- geometry
def geometry_forward(points: Tensor, space_cache: Tensor, weights: Dict[str, List[Tensor]], cfg: PathConfig,
                     output_normal: bool = False, create_graph: Optional[bool] = None) -> Dict[str, Tensor]:
    """GEO:273-351. points [B,N,3] (world), space_cache [B,6,C,R,R]."""
    B, N, _ = points.shape
    grad_enabled = torch.is_grad_enabled()
    if create_graph is None:
        create_graph = grad_enabled
    if output_normal:
        torch.set_grad_enabled(True)
        points = points.detach().requires_grad_(True) if not points.requires_grad else points
    try:
        points_unscaled = points
        p = rescale_points(points, cfg.radius)
        enc_geo, enc_tex = interpolate_encodings(p, space_cache, rotate=cfg.rotate_planes)
        sdf_orig = vanilla_mlp(enc_geo, weights["sdf"]).view(B, N, 1)
        sdf = get_shifted_sdf(points_unscaled, sdf_orig, cfg.sdf_bias_radius)
        out = {"sdf": sdf.view(B * N, 1), "sdf_orig": sdf_orig.view(B * N, 1)}
        out["features"] = vanilla_mlp(enc_tex, weights["feature"]).view(B * N, -1)
        if output_normal:
            sdf_grad = torch.autograd.grad(sdf, points_unscaled, grad_outputs=torch.ones_like(sdf),
                                           create_graph=create_graph)[0]
            normal = F.normalize(sdf_grad, dim=-1)
            if not create_graph:
                sdf_grad, normal = sdf_grad.detach(), normal.detach()
            out.update({"normal": normal.view(B * N, 3), "shading_normal": normal.view(B * N, 3),
                        "sdf_grad": sdf_grad.view(B * N, 3)})
    finally:
        torch.set_grad_enabled(grad_enabled)
    if not grad_enabled:
        out = {k: v.detach() for k, v in out.items()}
    return out


def forward_sdf(points: Tensor, space_cache: Tensor, weights, cfg: PathConfig) -> Tensor:
    """GEO:353-373."""
    B = points.shape[0]
    p = rescale_points(points, cfg.radius)
    enc = interpolate_encodings(p.reshape(B, -1, 3), space_cache, only_geo=True, rotate=cfg.rotate_planes)
    enc = enc.reshape(*points.shape[:-1], -1)
    sdf = vanilla_mlp(enc, weights["sdf"]).reshape(*points.shape[:-1], 1)
    return get_shifted_sdf(points, sdf, cfg.sdf_bias_radius)


def forward_field(points: Tensor, space_cache: Tensor, weights, cfg: PathConfig) -> Tuple[Tensor, Optional[Tensor]]:
    """GEO:375-394. points [B,M,3] -> sdf [B,M,1], deformation [B,M,3] or None."""
    p = rescale_points(points, cfg.radius)
    enc = interpolate_encodings(p, space_cache, only_geo=True, rotate=cfg.rotate_planes)
    sdf = vanilla_mlp(enc, weights["sdf"]).reshape(*points.shape[:-1], 1)
    sdf = get_shifted_sdf(points, sdf, cfg.sdf_bias_radius)
    deformation = None
    if weights.get("deformation") is not None:
        deformation = vanilla_mlp(enc, weights["deformation"]).reshape(*points.shape[:-1], 3)
    return sdf, deformation


def export_features(points: Tensor, space_cache: Tensor, weights, cfg: PathConfig) -> Tensor:
    """GEO:402-430 (batch 1)."""
    orig = points.shape
    pts = rescale_points(points.reshape(1, -1, 3), cfg.radius)
    _, enc_tex = interpolate_encodings(pts, space_cache, rotate=cfg.rotate_planes)
    return vanilla_mlp(enc_tex, weights["feature"]).view(orig[:-1] + (-1,))


def isosurface_grid_points(resolution: int, device="cpu") -> Tensor:
    """threestudio/models/isosurface.py:37-51 + triplaneturbo_executable/utils/mesh_exporter.py:95-100.

    ``linspace(0,1,R)`` meshgrid('ij') -> index ``(ix*R + iy)*R + iz`` -> scale_tensor((0,1) -> (-1,1)).
    """
    lin = torch.linspace(0, 1, resolution, device=device)
    x, y, z = torch.meshgrid(lin, lin, lin, indexing="ij")
    verts = torch.stack([x, y, z], dim=-1).reshape(-1, 3)
    verts = verts * (1 - 0) + 0
    return scale_tensor(verts, (0, 1), (-1, 1))


# --------------------------------------------------------------------------- NeuS alpha
def inv_std_of(learned_variance_init: float, like: Tensor) -> Tensor:
    """REN:24-35: ``exp(10 * p).clamp(1e-6, 1e6)``."""
    return torch.exp(torch.as_tensor(learned_variance_init, dtype=like.dtype, device=like.device) * 10.0).clamp(1e-6, 1e6)


def get_alpha(sdf: Tensor, normal: Tensor, dirs: Tensor, dists: Tensor, inv_std: Tensor,
              cos_anneal_ratio: float = 1.0) -> Tensor:
    """NEUS:93-117 (``use_volsdf=False``)."""
    true_cos = (dirs * normal).sum(-1, keepdim=True)
    iter_cos = -(F.relu(-true_cos * 0.5 + 0.5) * (1.0 - cos_anneal_ratio) + F.relu(-true_cos) * cos_anneal_ratio)
    estimated_next_sdf = sdf + iter_cos * dists * 0.5
    estimated_prev_sdf = sdf - iter_cos * dists * 0.5
    prev_cdf = torch.sigmoid(estimated_prev_sdf * inv_std)
    next_cdf = torch.sigmoid(estimated_next_sdf * inv_std)
    p = prev_cdf - next_cdf
    c = prev_cdf
    return ((p + 1e-5) / (c + 1e-5)).clip(0.0, 1.0)


def proposal_density(sdf: Tensor, inv_std: Tensor, render_step_size: float) -> Tensor:
    """REN:289-297."""
    estimated_next_sdf = sdf - render_step_size * 0.5
    estimated_prev_sdf = sdf + render_step_size * 0.5
    prev_cdf = torch.sigmoid(estimated_prev_sdf * inv_std)
    next_cdf = torch.sigmoid(estimated_next_sdf * inv_std)
    p = prev_cdf - next_cdf
    c = prev_cdf
    alpha = ((p + 1e-5) / (c + 1e-5)).clip(0.0, 1.0)
    return alpha / render_step_size


# --------------------------------------------------------------------------- sampler
def transform_stot(s_vals: Tensor, t_min: float, t_max: float) -> Tensor:
    """EST:104-118, ``sampling_type='uniform'``."""
    return s_vals * t_max + (1 - s_vals) * t_min


@torch.no_grad()
def importance_estimator_sampling(prop_sigma_fns: List[Callable], prop_samples: List[int], num_samples: int,
                                  n_rays: int, near_plane: float, far_plane: float, stratified: bool = False,
                                  jitters: Optional[List[Tensor]] = None, device="cpu",
                                  dtype=torch.float32) -> Tuple[Tensor, Tensor]:
    """EST:22-101. ``jitters``: one ``[n_rays]`` tensor per importance_sampling call when stratified."""
    cdfs = torch.cat([torch.zeros((n_rays, 1), device=device, dtype=dtype),
                      torch.ones((n_rays, 1), device=device, dtype=dtype)], dim=-1)
    vals = cdfs
    k = 0
    t_vals = None
    for level_fn, level_samples in zip(prop_sigma_fns, prop_samples):
        vals = nerfacc.importance_sampling(vals, cdfs, level_samples, stratified,
                                           jitters[k] if stratified else None)
        k += 1
        t_vals = transform_stot(vals, near_plane, far_plane)
        t_starts, t_ends = t_vals[..., :-1], t_vals[..., 1:]
        sigmas = level_fn(t_starts, t_ends)
        assert sigmas.shape == t_starts.shape
        trans, _ = nerfacc.render_transmittance_from_density(t_starts, t_ends, sigmas)
        cdfs = 1.0 - torch.cat([trans, torch.zeros_like(trans[:, :1])], dim=-1)
    vals_fine = nerfacc.importance_sampling(vals, cdfs, num_samples, stratified,
                                            jitters[k] if stratified else None)
    t_vals_fine = transform_stot(vals_fine, near_plane, far_plane)
    t_vals = torch.cat([t_vals, t_vals_fine], dim=-1)
    t_vals, _ = torch.sort(t_vals, dim=-1)
    return t_vals[..., :-1], t_vals[..., 1:]


def sample_intervals(rays_o: Tensor, rays_d: Tensor, space_cache: Tensor, weights, cfg: PathConfig,
                     stratified: bool = False, jitters=None) -> Tuple[Tensor, Tensor]:
    """REN:243-316: proposal closure + estimator. rays [B,H,W,3]; returns t_starts, t_ends ``[Nr, S]``."""
    B = space_cache.shape[0]
    o = rays_o.reshape(-1, 3)
    d = rays_d.reshape(-1, 3)
    n_rays = o.shape[0]

    def prop_sigma_fn(t_starts, t_ends):
        positions = o.unsqueeze(-2) + d.unsqueeze(-2) * (t_starts + t_ends)[..., None] / 2.0
        with torch.no_grad():
            geo_out = geometry_forward(positions.reshape(B, -1, 3), space_cache, weights, cfg, output_normal=False)
            inv_std = inv_std_of(cfg.learned_variance_init, geo_out["sdf"])
            return proposal_density(geo_out["sdf"], inv_std, cfg.render_step_size).reshape(positions.shape[:2])

    return importance_estimator_sampling([prop_sigma_fn], [cfg.num_samples_per_ray_importance],
                                         cfg.num_samples_per_ray, n_rays, cfg.near_plane, cfg.far_plane,
                                         stratified, jitters, device=o.device, dtype=o.dtype)


# --------------------------------------------------------------------------- renderer
def render_forward(rays_o: Tensor, rays_d: Tensor, space_cache: Tensor, weights, cfg: PathConfig,
                   bg_color: Tensor, camera_distances: Tensor, c2w: Tensor, training: bool = True,
                   t_starts: Optional[Tensor] = None, t_ends: Optional[Tensor] = None,
                   views_per_prompt: Optional[int] = None) -> Dict[str, Tensor]:
    """REN:98-546 (`forward` + `_forward`) for tensor space caches.

    rays [B,H,W,3]; space_cache [P,6,C,R,R] with ``B = P * V``; bg_color broadcastable to [Nr,3]
    (the background module is out of scope: its output enters as this tensor, REN:356-362,433-437).
    ``t_starts/t_ends [Nr,S]`` may be injected so that both sides of a parity test march identical
    intervals; otherwise they come from :func:`sample_intervals` (non-stratified).
    """
    B, H, W = rays_o.shape[:3]
    P = space_cache.shape[0]
    if P != B:  # REN:119-127
        assert B % P == 0
        space_cache = space_cache.repeat_interleave(B // P, dim=0)
    V = views_per_prompt if views_per_prompt is not None else B // P
    o = rays_o.reshape(-1, 3)
    d = rays_d.reshape(-1, 3)
    n_rays = o.shape[0]

    if t_starts is None:
        t_starts, t_ends = sample_intervals(rays_o, rays_d, space_cache, weights, cfg)
    S = t_starts.shape[1]
    ray_indices = torch.arange(n_rays, device=o.device).unsqueeze(-1).expand(-1, S).flatten().long()  # REN:317-332
    t_starts_ = t_starts.flatten()[..., None]
    t_ends_ = t_ends.flatten()[..., None]
    t_origins = o[ray_indices]
    t_dirs = d[ray_indices]
    t_positions = (t_starts_ + t_ends_) / 2.0                      # REN:337
    positions = t_origins + t_dirs * t_positions                   # REN:338
    t_intervals = t_ends_ - t_starts_                              # REN:339

    geo_out = geometry_forward(positions.reshape(B, -1, 3), space_cache, weights, cfg, output_normal=True)  # REN:342-346
    rgb_fg_all = sigmoid_mipnerf(geo_out["features"])              # MAT:41-54, REN:347-353
    if cfg.rgb_grad_shrink != 1.0:                                 # REN:397-400
        rgb_fg_all = cfg.rgb_grad_shrink * rgb_fg_all + (1.0 - cfg.rgb_grad_shrink) * rgb_fg_all.detach()

    inv_std = inv_std_of(cfg.learned_variance_init, geo_out["sdf"])
    alpha = get_alpha(geo_out["sdf"], geo_out["normal"], t_dirs, t_intervals, inv_std, cfg.cos_anneal_ratio)  # REN:403
    weights_, _ = nerfacc.render_weight_from_alpha(alpha[..., 0], ray_indices, n_rays)   # REN:408
    w = weights_[..., None]
    opacity = nerfacc.accumulate_along_rays(w[..., 0], None, ray_indices, n_rays)        # REN:414
    depth = nerfacc.accumulate_along_rays(w[..., 0], t_positions, ray_indices, n_rays)   # REN:417
    comp_rgb_fg = nerfacc.accumulate_along_rays(w[..., 0], rgb_fg_all, ray_indices, n_rays)  # REN:420
    t_depth = depth[ray_indices]
    z_variance = nerfacc.accumulate_along_rays(w[..., 0], (t_positions - t_depth) ** 2, ray_indices, n_rays)  # REN:426

    comp_rgb_bg = bg_color
    if bg_color.shape[:-1] == (B, H, W):                           # REN:436-437
        bg_color = bg_color.reshape(B * H * W, -1)
    comp_rgb = comp_rgb_fg + bg_color * (1.0 - opacity)            # REN:439

    out = {
        "comp_rgb": comp_rgb.view(B, H, W, -1),
        "comp_rgb_fg": comp_rgb_fg.view(B, H, W, -1),
        "opacity": opacity.view(B, H, W, 1),
        "depth": depth.view(B, H, W, 1),
        "z_variance": z_variance.view(B, H, W, 1),
    }
    # REN:451-462 (RichDreamer disparity)
    sqrt3 = torch.sqrt(3 * torch.ones(1, 1, 1, 1, device=o.device, dtype=o.dtype))
    far = camera_distances.reshape(-1, 1, 1, 1) + sqrt3
    near = camera_distances.reshape(-1, 1, 1, 1) - sqrt3
    disparity_tmp = out["depth"] * out["opacity"] + (1.0 - out["opacity"]) * far
    out["disparity"] = torch.clamp((far - disparity_tmp) / (far - near), 0.0, 1.0).view(B, H, W, 1)

    # REN:466-530
    comp_normal = nerfacc.accumulate_along_rays(w[..., 0], geo_out["normal"], ray_indices, n_rays)
    comp_normal = F.normalize(comp_normal, dim=-1)
    out["comp_normal"] = comp_normal.view(B, H, W, 3)
    if cfg.normal_direction == "camera":
        bg_normal = 0.5 * torch.ones_like(comp_normal)
        bg_normal[:, 2] = 1.0
        bg_normal_white = torch.ones_like(comp_normal)
        w2c = torch.inverse(c2w)
        rot = w2c[:, :3, :3]
        comp_normal_cam = comp_normal.view(B, -1, 3) @ rot.permute(0, 2, 1)
        flip_x = torch.eye(3, device=o.device, dtype=o.dtype)
        flip_x[0, 0] = -1
        comp_normal_cam = (comp_normal_cam @ flip_x[None, :, :]).view(-1, 3)
        out["comp_normal_cam_vis"] = ((comp_normal_cam + 1.0) / 2.0 * opacity + (1 - opacity) * bg_normal).view(B, H, W, 3)
        out["comp_normal_cam_vis_white"] = ((comp_normal_cam + 1.0) / 2.0 * opacity + (1 - opacity) * bg_normal_white).view(B, H, W, 3)
    elif cfg.normal_direction == "front":
        bg_normal_white = torch.ones_like(comp_normal)
        c2w_front = c2w[0::V].repeat_interleave(V, dim=0)
        rot = torch.inverse(c2w_front)[:, :3, :3]
        comp_normal_front = (comp_normal.view(B, -1, 3) @ rot.permute(0, 2, 1)).view(-1, 3)
        out["comp_normal_cam_vis_white"] = ((comp_normal_front + 1.0) / 2.0 * opacity + (1 - opacity) * bg_normal_white).view(B, H, W, 3)

    if training:  # REN:532-545
        out.update({"weights": w, "t_points": t_positions, "t_intervals": t_intervals, "t_dirs": t_dirs,
                    "ray_indices": ray_indices, "points": positions, **geo_out, "inv_std": inv_std})
    return out


def patch_render(render_fn: Callable, rays_o: Tensor, rays_d: Tensor, patch_size: int, global_downsample: int,
                 patch_xy: Tuple[int, int], global_detach: bool = False) -> Dict[str, Tensor]:
    """PATCH:39-89 (training branch). ``render_fn(rays_o, rays_d) -> dict``; ``patch_xy`` replaces the
    reference's ``torch.randint`` draw (PATCH:66-67)."""
    B, H, W, _ = rays_o.shape
    ds = global_downsample
    g_o = F.interpolate(rays_o.permute(0, 3, 1, 2), (H // ds, W // ds), mode="bilinear").permute(0, 2, 3, 1)
    g_d = F.interpolate(rays_d.permute(0, 3, 1, 2), (H // ds, W // ds), mode="bilinear").permute(0, 2, 3, 1)
    out_global = render_fn(g_o, g_d)
    PS = patch_size
    px, py = patch_xy
    out = render_fn(rays_o[:, py:py + PS, px:px + PS], rays_d[:, py:py + PS, px:px + PS])
    keys = [k for k in out if torch.is_tensor(out[k]) and out[k].dim() == out["comp_rgb"].dim()
            and out[k][..., 0].shape == out["comp_rgb"][..., 0].shape]
    for k in keys:
        out_global[k] = F.interpolate(out_global[k].permute(0, 3, 1, 2), (H, W), mode="bilinear").permute(0, 2, 3, 1)
        if global_detach:
            out_global[k] = out_global[k].detach()
        out_global[k][:, py:py + PS, px:px + PS] = out[k]
    return out_global
