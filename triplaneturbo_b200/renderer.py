"""Renderer plugin: NeuS-style SDF volume renderer over per-prompt triplane space caches, on sm_100a kernels.

Registered under the reference's names and keeping its call surface:
  * ``generative-space-sdf-volume-renderer``
    (custom/triplaneturbo/models/renderers/generative_space_sdf_volume_renderer.py:38-565)
  * ``patch-renderer`` (threestudio/models/renderers/patch_renderer.py:14-106)
  * ``ImportanceEstimator.sampling`` (threestudio/models/estimators.py:22-101)
One call = tt_importance_sample + tt_render_fwd (+ tt_render_bwd under autograd); nothing of size
[rays x samples x channels] is ever materialised.
"""
import dataclasses
import math
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import ops
from .compat import BaseModule, C, find, register
from .image_ops import compose_images


class LearnedVariance(nn.Module):
    """…sdf_volume_renderer.py:24-35; state-dict key ``_inv_std``."""

    def __init__(self, init_val, requires_grad=True):
        super().__init__()
        self.register_parameter("_inv_std", nn.Parameter(torch.tensor(float(init_val)), requires_grad=requires_grad))
        self._cached = None

    @property
    def inv_std(self):
        return torch.exp(self._inv_std * 10.0)

    def value(self) -> float:
        """Python float of clamp(exp(10 p), 1e-6, 1e6) without a device sync per call."""
        key = (self._inv_std.data_ptr(), self._inv_std._version)
        if self._cached is None or self._cached[0] != key:
            self._cached = (key, float(self.inv_std.detach().clamp(1.0e-6, 1.0e6).item()))
        return self._cached[1]

    def forward(self, x):
        return torch.ones_like(x) * self.inv_std.clamp(1.0e-6, 1.0e6)


class ImportanceEstimator(nn.Module):
    """Drop-in for threestudio/models/estimators.py:15-118, same ``sampling`` signature.

    ``prop_sigma_fns`` is a list of callables ``(t_starts, t_ends) -> densities`` like the reference's.  The one closure
    that exists on this path is the renderer's SDF density (…sdf_volume_renderer.py:243-299); the renderer passes it as
    a :class:`ProposalSpec` (a callable that also carries the data the closure captured), and ``sampling`` then runs the
    whole estimator -- both nerfacc.importance_sampling calls, the density evaluation, the transmittance scan and the
    final merge-sort -- as ONE fused call (``tt_importance_sample``).  Any other callable takes the general route:
    the estimator's steps as written in the reference, on torch device tensors, with the closure called as is.
    """

    @dataclass
    class ProposalSpec:
        planes: Tensor
        wpack: Tensor
        scalars: ops.PathScalars
        rays_o: Tensor
        rays_d: Tensor
        rays_per_cache: int

        def __call__(self, t_starts: Tensor, t_ends: Tensor) -> Tensor:
            """The proposal closure itself (…sdf_volume_renderer.py:243-299): NeuS density of the SDF at the interval
            midpoints.  Used when a caller evaluates the closure directly; ``sampling`` fuses it instead."""
            n, m = t_starts.shape
            t_mid = (t_starts + t_ends) / 2.0
            pos = self.rays_o.view(-1, 1, 3) + self.rays_d.view(-1, 1, 3) * t_mid[..., None]
            P = self.planes.shape[0]
            sdf = ops.geometry_fwd(self.planes, self.wpack, self.scalars, pos.reshape(P, -1, 3).contiguous(), 0,
                                   ["sdf"])["sdf"].view(n, m)
            s = self.scalars
            prev_cdf = torch.sigmoid((sdf + s.render_step_size * 0.5) * s.inv_std)
            next_cdf = torch.sigmoid((sdf - s.render_step_size * 0.5) * s.inv_std)
            alpha = ((prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)).clip(0.0, 1.0)
            return alpha / s.render_step_size

    @staticmethod
    def _quantiles(n: int, n_rays: int, jitter: Optional[Tensor], device) -> Tensor:
        j = torch.arange(n + 1, dtype=torch.float32, device=device)
        if jitter is None:
            return (j / float(n))[None, :].expand(n_rays, -1).contiguous()
        return ((j[None, :] + jitter[:, None]) / float(n + 1)).contiguous()

    @staticmethod
    def _importance_sampling(vals: Tensor, cdfs: Tensor, n: int, jitter: Optional[Tensor]) -> Tensor:
        """nerfacc.pdf.importance_sampling on dense rays: n+1 edge quantiles inverted through the piece-wise linear
        CDF (same conventions as the fused kernel: u_j = j/n, or (j+b)/(n+1) with one jitter b per ray)."""
        n_rays, n_in = cdfs.shape
        u = ImportanceEstimator._quantiles(n, n_rays, jitter, cdfs.device)
        p = torch.searchsorted(cdfs.contiguous(), u, right=True).clamp(1, n_in - 1)
        c0, c1 = torch.gather(cdfs, 1, p - 1), torch.gather(cdfs, 1, p)
        v0, v1 = torch.gather(vals, 1, p - 1), torch.gather(vals, 1, p)
        den = c1 - c0
        frac = torch.where(den > 0, (u - c0) / den, torch.zeros_like(u)).clamp(0.0, 1.0)
        return v0 + frac * (v1 - v0)

    @torch.no_grad()
    def sampling(self, prop_sigma_fns: List[Any], prop_samples: List[int], num_samples: int, n_rays: int,
                 near_plane: float, far_plane: float, sampling_type: str = "uniform", stratified: bool = False,
                 requires_grad: bool = False, jitters: Optional[Tuple[Tensor, ...]] = None
                 ) -> Tuple[Tensor, Tensor]:
        assert len(prop_sigma_fns) == len(prop_samples), \
            "The number of proposal networks and the number of samples should be the same."
        if sampling_type not in ("uniform", "lindisp"):
            raise ValueError(f"Unknown transform_type: {sampling_type}")
        fused = (len(prop_sigma_fns) == 1 and isinstance(prop_sigma_fns[0], ImportanceEstimator.ProposalSpec)
                 and sampling_type == "uniform")
        if fused:
            spec = prop_sigma_fns[0]
            s = ops.PathScalars(**{**spec.scalars.__dict__, "near_plane": float(near_plane),
                                   "far_plane": float(far_plane)})
            j0 = j1 = None
            if stratified:
                if jitters is None:   # one offset per ray and per importance_sampling call (nerfacc draws its own)
                    j0 = torch.rand(n_rays, device=spec.rays_o.device)
                    j1 = torch.rand(n_rays, device=spec.rays_o.device)
                else:
                    j0, j1 = jitters
            t_vals = ops.importance_sample(spec.planes, spec.wpack, s, spec.rays_o, spec.rays_d, spec.rays_per_cache,
                                           int(prop_samples[0]), int(num_samples), j0, j1)
            assert t_vals.shape[0] == n_rays
            return t_vals[:, :-1], t_vals[:, 1:]
        # ---- general route: arbitrary closures, the estimator's steps as in estimators.py:63-101 ----------------
        dev = next((getattr(f, "rays_o", None) for f in prop_sigma_fns if hasattr(f, "rays_o")), None)
        dev = dev.device if dev is not None else torch.device("cuda", torch.cuda.current_device())

        def stot(sv):
            if sampling_type == "uniform":
                return sv * far_plane + (1 - sv) * near_plane
            return 1.0 / (sv * (1.0 / far_plane) + (1 - sv) * (1.0 / near_plane))
        n_lv = len(prop_sigma_fns)
        if stratified and jitters is None:
            jitters = tuple(torch.rand(n_rays, device=dev) for _ in range(n_lv + 1))
        vals = torch.cat([torch.zeros((n_rays, 1), device=dev), torch.ones((n_rays, 1), device=dev)], dim=-1)
        cdfs = vals.clone()
        t_vals = None
        for lv, (level_fn, level_samples) in enumerate(zip(prop_sigma_fns, prop_samples)):
            vals = self._importance_sampling(vals, cdfs, int(level_samples), jitters[lv] if stratified else None)
            t_vals = stot(vals)
            t0, t1 = t_vals[..., :-1], t_vals[..., 1:]
            with torch.set_grad_enabled(requires_grad):
                sigmas = level_fn(t0, t1)
                assert sigmas.shape == t0.shape
                sdt = sigmas * (t1 - t0)
                excl = torch.cumsum(torch.cat([torch.zeros_like(sdt[:, :1]), sdt[:, :-1]], dim=-1), dim=-1)
                trans = torch.exp(-excl)
                cdfs = 1.0 - torch.cat([trans, torch.zeros_like(trans[:, :1])], dim=-1)
        vals_f = self._importance_sampling(vals, cdfs, int(num_samples), jitters[n_lv] if stratified else None)
        t_fine = stot(vals_f)
        t_all = t_fine if t_vals is None else torch.cat([t_vals, t_fine], dim=-1)
        t_all, _ = torch.sort(t_all, dim=-1)
        return t_all[..., :-1], t_all[..., 1:]


@register("generative-space-sdf-volume-renderer")
class GenerativeSpaceSDFVolumeRenderer(BaseModule):
    @dataclass
    class Config(BaseModule.Config):
        radius: float = 1.0
        num_samples_per_ray: int = 512
        randomized: bool = True
        eval_chunk_size: int = 320000      # accepted for config compatibility; the fused kernel needs no chunking
        learned_variance_init: float = 0.3
        cos_anneal_end_steps: int = 0
        use_volsdf: bool = False
        near_plane: float = 0.0
        far_plane: float = 1e10
        trainable_variance: bool = True
        estimator: str = "occgrid"
        grid_prune: bool = True
        prune_alpha_threshold: bool = True
        num_samples_per_ray_importance: int = 64
        train_chunk_size: int = 0          # ditto
        rgb_grad_shrink: Any = 1.0
        normal_direction: str = "camera"
        return_samples: bool = True        # training extras (weights, sdf, normal, … per sample), REN:532-545
        anneal_cos_in_update_step: bool = False   # the reference's GenerativeSpaceSDFVolumeRenderer.update_step
                                                  # (…sdf_volume_renderer.py:548-553) never touches cos_anneal_ratio,
                                                  # so it stays 1.0; True applies neus_volume_renderer.py:87-91

    cfg: Config

    def configure(self, geometry, material, background) -> None:
        @dataclass
        class SubModules:
            geometry: Any
            material: Any
            background: Any
        self.sub_modules = SubModules(geometry, material, background)
        r = self.cfg.radius
        self.register_buffer("bbox", torch.as_tensor([[-r, -r, -r], [r, r, r]], dtype=torch.float32))
        self.variance = LearnedVariance(self.cfg.learned_variance_init, requires_grad=self.cfg.trainable_variance)
        if self.cfg.estimator == "occgrid":
            raise NotImplementedError("Occgrid estimator not supported for generative-space-volsdf-volume-renderer")
        if self.cfg.estimator != "importance":
            raise NotImplementedError(f"Estimator {self.cfg.estimator} not implemented")
        if self.cfg.use_volsdf:
            raise NotImplementedError("use_volsdf=True is not on the shipped path (configs/TriplaneTurbo_v1.yaml:136)")
        assert self.cfg.normal_direction in ["front", "camera", "world"], \
            "normal_direction must be in ['front', 'camera', 'world']"
        self.estimator = ImportanceEstimator()
        self.render_step_size = 1.732 * 2 * self.cfg.radius / self.cfg.num_samples_per_ray
        self.cos_anneal_ratio = 1.0
        self.randomized = self.cfg.randomized
        self.rgb_grad_shrink = C(self.cfg.rgb_grad_shrink, 0, 0)
        material_ok = getattr(material, "is_sigmoid_mipnerf_no_material", False)
        if material is not None and not material_ok:
            raise NotImplementedError("the fused kernel implements NoMaterial with color_activation sigmoid-mipnerf")

    @property
    def geometry(self):
        return self.sub_modules.geometry

    @property
    def material(self):
        return self.sub_modules.material

    @property
    def background(self):
        return self.sub_modules.background

    def path_scalars(self) -> ops.PathScalars:
        return self.geometry.path_scalars(inv_std=self.variance.value(), cos_anneal_ratio=float(self.cos_anneal_ratio),
                                          near_plane=float(self.cfg.near_plane), far_plane=float(self.cfg.far_plane),
                                          render_step_size=float(self.render_step_size))

    # ------------------------------------------------------------------------------------------------
    def forward(self, rays_o: Tensor, rays_d: Tensor, light_positions: Optional[Tensor] = None,
                bg_color: Optional[Tensor] = None, noise: Optional[Tensor] = None,
                space_cache: Optional[Tensor] = None, text_embed: Optional[Tensor] = None, **kwargs
                ) -> Dict[str, Tensor]:
        """REN:98-199.  ``space_cache`` [P,6,C,R,R]; rays [B,H,W,3] with B = P * views.  The reference copies the
        cache per view (training) or loops view by view (eval); here every ray indexes its prompt's planes."""
        batch_size = rays_o.shape[0]
        P = text_embed.shape[0] if text_embed is not None else batch_size
        if space_cache is None:
            space_cache = self.geometry.generate_space_cache(styles=noise, text_embed=text_embed)
        if not torch.is_tensor(space_cache):
            raise NotImplementedError("space_cache must be a tensor [P,6,C,R,R]")
        if space_cache.shape[0] not in (P, batch_size):
            raise AssertionError("space_cache must have the batch size of text_embed or of rays_o")
        if batch_size % space_cache.shape[0] != 0:
            raise AssertionError("batch size of rays_o must be a multiple of the space_cache batch")
        if not self.training and space_cache.shape[0] != batch_size:
            assert space_cache.shape[0] == 1, "batch_size of space_cache must be 1 or equal to batch_size of rays_o"
        return self._forward(rays_o, rays_d, light_positions, bg_color, space_cache=space_cache,
                             text_embed=text_embed, **kwargs)

    def _forward(self, rays_o, rays_d, light_positions=None, bg_color=None, space_cache=None, text_embed=None,
                 camera_distances=None, c2w=None, t_starts=None, t_ends=None, jitters=None, **kwargs):
        B, H, W = rays_o.shape[:3]
        Pc = space_cache.shape[0]
        views_per_cache = B // Pc
        P_text = text_embed.shape[0] if text_embed is not None else B
        num_views_per_batch = B // P_text
        o = rays_o.reshape(-1, 3).contiguous()
        d = rays_d.reshape(-1, 3).contiguous()
        n_rays = o.shape[0]
        rays_per_cache = views_per_cache * H * W
        geom = self.geometry
        scalars = dataclasses.replace(self.path_scalars(), image_h=int(H), image_w=int(W))
        weights = geom.decoder_weights()
        C_ = geom.plane_channels

        if t_starts is None:       # REN:243-316
            spec = ImportanceEstimator.ProposalSpec(
                ops.cached_planes(space_cache, C_),
                ops.cached_wpack(weights[:3], weights[3:], geom._deformation_weights(), C_), scalars, o, d,
                rays_per_cache)
            t_starts, t_ends = self.estimator.sampling(
                prop_sigma_fns=[spec], prop_samples=[self.cfg.num_samples_per_ray_importance],
                num_samples=self.cfg.num_samples_per_ray, n_rays=n_rays, near_plane=self.cfg.near_plane,
                far_plane=self.cfg.far_plane, sampling_type="uniform", stratified=self.randomized, jitters=jitters)
        S = t_starts.shape[1]

        extras = bool(self.training and self.cfg.return_samples)
        inv_std_t = self.variance.inv_std.clamp(1.0e-6, 1.0e6)
        res = ops.RenderFunction.apply(space_cache, *weights, inv_std_t, o, d, t_starts, t_ends, scalars,
                                       rays_per_cache, float(self.rgb_grad_shrink), extras)
        acc = res[0]

        # background (REN:356-362,433-437): an out-of-scope module whose output enters as a tensor
        if bg_color is None:
            bgm = self.background
            if getattr(bgm, "enabling_hypernet", False):
                bg_color = bgm(dirs=rays_d, text_embed=kwargs.get("text_embed_bg", text_embed))
            else:
                bg_color = bgm(dirs=rays_d)
        if not self.training and Pc != B:
            num_views_per_batch = 1      # REN:150-176: eval renders view by view, each view is its own "front"
        out = compose_images(acc, bg_color, camera_distances, c2w, B, H, W, self.cfg.normal_direction,
                             num_views_per_batch)
        if out["comp_rgb_bg"].dim() < 4:
            out["comp_rgb_bg"] = out["comp_rgb_bg"].expand(B, H, W, -1) if out["comp_rgb_bg"].dim() == 1 \
                else out["comp_rgb_bg"].view(B, H, W, -1)

        if self.training:          # REN:532-545
            if extras:
                _, sdf, sdf_orig, sdf_grad, normal, features, wts = res
                ray_indices = torch.arange(n_rays, device=o.device).unsqueeze(-1).expand(-1, S).flatten().long()
                t0 = t_starts.reshape(-1, 1)
                t1 = t_ends.reshape(-1, 1)
                t_positions = (t0 + t1) / 2.0
                t_dirs = d[ray_indices]
                out.update({"weights": wts, "t_points": t_positions, "t_intervals": t1 - t0, "t_dirs": t_dirs,
                            "ray_indices": ray_indices, "points": o[ray_indices] + t_dirs * t_positions,
                            "sdf": sdf, "sdf_orig": sdf_orig, "features": features, "normal": normal,
                            "shading_normal": normal, "sdf_grad": sdf_grad})
            out.update({"inv_std": self.variance.inv_std})
        return out

    def update_step(self, epoch: int, global_step: int, on_load_weights: bool = False) -> None:
        self.rgb_grad_shrink = C(self.cfg.rgb_grad_shrink, epoch, global_step)
        if self.cfg.anneal_cos_in_update_step:      # threestudio/models/renderers/neus_volume_renderer.py:87-91
            self.cos_anneal_ratio = 1.0 if self.cfg.cos_anneal_end_steps == 0 else \
                min(1.0, global_step / self.cfg.cos_anneal_end_steps)

    def train(self, mode=True):                     # …sdf_volume_renderer.py:555-565
        self.randomized = mode and self.cfg.randomized
        if self.geometry is not None and hasattr(self.geometry, "train"):
            self.geometry.train(mode)
        return super().train(mode=mode)

    def eval(self):
        self.randomized = False
        if self.geometry is not None and hasattr(self.geometry, "eval"):
            self.geometry.eval()
        return super().eval()


@register("patch-renderer")
class PatchRenderer(BaseModule):
    """threestudio/models/renderers/patch_renderer.py:14-106 (image-space wrapper, stays in torch)."""

    @dataclass
    class Config(BaseModule.Config):
        radius: float = 1.0
        patch_size: int = 128
        base_renderer_type: str = ""
        base_renderer: Optional[Any] = None
        global_detach: bool = False
        global_downsample: int = 4

    cfg: Config

    def configure(self, geometry, material, background) -> None:
        self.base_renderer = find(self.cfg.base_renderer_type)(self.cfg.base_renderer, geometry=geometry,
                                                               material=material, background=background)

    def forward(self, rays_o, rays_d, light_positions=None, bg_color=None, **kwargs):
        B, H, W, _ = rays_o.shape
        if not self.base_renderer.training:
            return self.base_renderer(rays_o, rays_d, light_positions, bg_color, **kwargs)
        ds = self.cfg.global_downsample

        def shrink(x):
            return F.interpolate(x.permute(0, 3, 1, 2), (H // ds, W // ds), mode="bilinear").permute(0, 2, 3, 1)
        out_global = self.base_renderer(shrink(rays_o), shrink(rays_d), light_positions, bg_color, **kwargs)
        PS = self.cfg.patch_size
        patch_x = torch.randint(0, W - PS, (1,)).item()
        patch_y = torch.randint(0, H - PS, (1,)).item()
        out = self.base_renderer(rays_o[:, patch_y:patch_y + PS, patch_x:patch_x + PS],
                                 rays_d[:, patch_y:patch_y + PS, patch_x:patch_x + PS], light_positions, bg_color,
                                 **kwargs)
        valid = [k for k in out if torch.is_tensor(out[k]) and out[k].dim() == out["comp_rgb"].dim()
                 and out[k][..., 0].shape == out["comp_rgb"][..., 0].shape]
        for key in valid:
            g = F.interpolate(out_global[key].permute(0, 3, 1, 2), (H, W), mode="bilinear").permute(0, 2, 3, 1)
            if self.cfg.global_detach:
                g = g.detach()
            g = g.clone() if not g.is_contiguous() else g
            g[:, patch_y:patch_y + PS, patch_x:patch_x + PS] = out[key]
            out_global[key] = g
        out_global.update({"patch_x": patch_x, "patch_y": patch_y})
        return out_global

    def update_step(self, epoch: int, global_step: int, on_load_weights: bool = False) -> None:
        self.base_renderer.update_step(epoch, global_step, on_load_weights)

    def train(self, mode=True):
        return self.base_renderer.train(mode)

    def eval(self):
        return self.base_renderer.eval()


@register("no-material")
class NoMaterial(BaseModule):
    """threestudio/models/materials/no_material.py:14-63 with ``color_activation: sigmoid-mipnerf`` and no MLP: the
    fused march applies it in-kernel; ``forward`` is kept for callers that shade feature tensors directly (mesh
    export)."""

    @dataclass
    class Config(BaseModule.Config):
        n_output_dims: int = 3
        color_activation: str = "sigmoid"
        input_feature_dims: Optional[int] = None
        mlp_network_config: Optional[dict] = None
        requires_normal: bool = False

    cfg: Config
    is_sigmoid_mipnerf_no_material = True

    def configure(self) -> None:
        if self.cfg.color_activation != "sigmoid-mipnerf" or self.cfg.mlp_network_config is not None:
            raise NotImplementedError("kernels implement NoMaterial(color_activation='sigmoid-mipnerf') without an MLP")
        self.requires_normal = self.cfg.requires_normal

    def forward(self, features: Tensor, **kwargs) -> Tensor:
        return torch.sigmoid(features.view(-1, self.cfg.n_output_dims)).view(features.shape) * (1 + 2 * 0.001) - 0.001

    def export(self, features: Tensor, **kwargs) -> Dict[str, Any]:
        color = self(features, **kwargs).clamp(0, 1)
        return {"albedo": color}


@register("solid-color-background")
class SolidColorBackground(BaseModule):
    """Constant background (threestudio/models/background/solid_color_background.py); the reference's hash-grid
    background is outside the path and enters the renderer only as a tensor."""

    @dataclass
    class Config(BaseModule.Config):
        n_output_dims: int = 3
        color: Tuple = (1.0, 1.0, 1.0)

    cfg: Config

    def configure(self) -> None:
        self.register_buffer("env_color", torch.as_tensor(self.cfg.color, dtype=torch.float32))

    def forward(self, dirs: Tensor, **kwargs) -> Tensor:
        return torch.ones(*dirs.shape[:-1], self.cfg.n_output_dims).to(dirs) * self.env_color.to(dirs)
