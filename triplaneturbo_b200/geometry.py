"""Geometry plugin: triplane sampler + SDF / feature / deformation decoders on hand-written sm_100a kernels.

Registered under the reference's name ``few-step-triplane-dual-stable-diffusion`` and keeping its call surface
(custom/triplaneturbo/models/geometry/few_step_triplane_dual_stable_diffusion.py:20-447; inference twin
triplaneturbo_executable/models/geometry/sd_dual_triplanes.py:66-394): ``forward``, ``forward_sdf``,
``forward_field``, ``forward_level``, ``export``, ``decode``, ``interpolate_encodings``, ``rescale_points``,
attributes ``bbox``, ``unbounded``, ``sdf_network`` / ``feature_network`` / ``deformation_network`` with the
state-dict keys ``layers.{0,2,4}.weight``.

The SD-UNet/VAE generator that emits the triplanes is outside the path; pass one in as
``space_generator`` (anything with ``forward_denoise`` / ``forward_decode``) to get ``denoise`` / ``decode``.
"""
from dataclasses import dataclass, field
from typing import Any, Dict, Optional, Tuple, Union

import torch
import torch.nn as nn
from torch import Tensor

from . import ops
from .compat import BaseModule, register


class VanillaMLP(nn.Module):
    """Parameter container with the reference's layout (threestudio/models/networks.py:67-104): Linear(no bias),
    ReLU, Linear, ReLU, Linear inside ``layers`` so checkpoints load unchanged.  On the path the evaluation happens in
    the fused kernels; ``forward`` exists for callers that evaluate a decoder on their own tensors."""

    def __init__(self, dim_in: int, dim_out: int, config: dict):
        super().__init__()
        self.n_neurons, self.n_hidden_layers = config["n_neurons"], config["n_hidden_layers"]
        if self.n_neurons != ops.HIDDEN or self.n_hidden_layers != 2:
            raise NotImplementedError("the sm_100a kernels are built for n_neurons=64, n_hidden_layers=2 "
                                      "(configs/TriplaneTurbo_v1.yaml:86-91)")
        if config.get("activation", "ReLU") != "ReLU" or config.get("output_activation", "none") not in (None, "none"):
            raise NotImplementedError("only ReLU hidden activations and no output activation are supported")
        self.layers = nn.Sequential(nn.Linear(dim_in, 64, bias=False), nn.ReLU(inplace=False),
                                    nn.Linear(64, 64, bias=False), nn.ReLU(inplace=False),
                                    nn.Linear(64, dim_out, bias=False))

    def weights(self):
        return [self.layers[0].weight, self.layers[2].weight, self.layers[4].weight]

    def forward(self, x):
        """Direct evaluation for callers outside the fused path (networks.py:90-95, autocast off): plain library
        linears.  The renderer / geometry never come here: they hand ``weights()`` to the CUDA kernels."""
        with torch.autocast(device_type=x.device.type, enabled=False):
            return self.layers(x.float())


@register("few-step-triplane-dual-stable-diffusion")
class StableDiffusionTriplaneDualAttention(BaseModule):
    @dataclass
    class Config(BaseModule.Config):
        radius: float = 1.0
        isosurface: bool = True
        isosurface_method: str = "mt"
        isosurface_resolution: int = 128
        isosurface_threshold: Union[float, str] = 0.0
        isosurface_chunk: int = 0
        isosurface_coarse_to_fine: bool = True
        isosurface_deformable_grid: bool = False
        isosurface_remove_outliers: bool = False
        isosurface_outlier_n_faces_threshold: Union[int, float] = 0.01
        n_feature_dims: int = 3
        space_generator_config: dict = field(default_factory=lambda: {"output_dim": 32})
        mlp_network_config: dict = field(default_factory=lambda: {
            "otype": "VanillaMLP", "activation": "ReLU", "output_activation": "none", "n_neurons": 64,
            "n_hidden_layers": 2})
        backbone: str = "few_step_triplane_dual_stable_diffusion"
        normal_type: Optional[str] = "analytic"
        finite_difference_normal_eps: Union[float, str] = 0.01
        sdf_bias: Union[float, str] = 0.0
        sdf_bias_params: Optional[Any] = None
        rotate_planes: Optional[str] = None
        split_channels: Optional[str] = None
        geo_interpolate: str = "v1"
        tex_interpolate: str = "v1"
        # hand the VAE decoder's raw output [B,6,2C,H,W] to the renderer untouched; the channel split of
        # `decode` (few_step…diffusion.py:186-196: a boolean-mask gather, 100 MB read + 50 MB written per prompt) then
        # happens inside the one repack pass (tt_repack_planes with c_off_tex = C).  Off by default: `decode` then returns
        # the reference's [B,6,C,H,W] tensor.
        fuse_channel_split: bool = False

    cfg: Config

    def configure(self, space_generator: Optional[nn.Module] = None) -> None:
        r = self.cfg.radius
        self.register_buffer("bbox", torch.as_tensor([[-r, -r, -r], [r, r, r]], dtype=torch.float32))
        self.unbounded = False
        self.space_generator = space_generator
        # the kernels implement the shipped configuration (configs/TriplaneTurbo_v1.yaml:73-92)
        if self.cfg.rotate_planes != "v1" or self.cfg.geo_interpolate != "v1" or self.cfg.tex_interpolate != "v2":
            raise NotImplementedError("kernels implement rotate_planes=v1, geo_interpolate=v1, tex_interpolate=v2")
        if self.cfg.sdf_bias != "sphere" or self.cfg.normal_type != "analytic":
            raise NotImplementedError("kernels implement sdf_bias=sphere and normal_type=analytic")
        if self.cfg.split_channels not in (None, "v1"):
            raise NotImplementedError("split_channels must be None or 'v1'")
        if self.cfg.n_feature_dims != 3:
            raise NotImplementedError("n_feature_dims must be 3")
        dim = int(self.cfg.space_generator_config["output_dim"])
        if self.cfg.split_channels == "v1":
            dim //= 2
        self.plane_channels = dim
        self.sdf_network = VanillaMLP(dim, 1, self.cfg.mlp_network_config)
        self.feature_network = VanillaMLP(3 * dim, 3, self.cfg.mlp_network_config)
        if self.cfg.isosurface_deformable_grid:
            self.deformation_network = VanillaMLP(dim, 3, self.cfg.mlp_network_config)

    # ------------------------------------------------------------------ helpers
    def path_scalars(self, **over) -> ops.PathScalars:
        return ops.PathScalars(radius=float(self.cfg.radius), sdf_bias_radius=float(self.cfg.sdf_bias_params), **over)

    def decoder_weights(self):
        return self.sdf_network.weights() + self.feature_network.weights()

    def _deformation_weights(self):
        return self.deformation_network.weights() if hasattr(self, "deformation_network") else None

    def _check_cache(self, space_cache: Tensor, batch: int) -> Tensor:
        if not torch.is_tensor(space_cache):
            raise NotImplementedError("space_cache must be a tensor [B,6,C,R,R]")
        if space_cache.shape[0] != batch:
            raise AssertionError("space_cache must have the same batch size as points")
        return space_cache

    # ------------------------------------------------------------------ generator hand-off
    def denoise(self, *args, **kwargs):
        if self.space_generator is None:
            raise RuntimeError("no space_generator attached (the SD generator is outside this package)")
        return self.space_generator.forward_denoise(*args, **kwargs)

    def decode(self, latents: Tensor) -> Tensor:
        """few_step…diffusion.py:180-196.  With a generator attached: VAE decode then channel split."""
        triplane = self.space_generator.forward_decode(latents) if self.space_generator is not None else latents
        if self.cfg.split_channels is None or self.cfg.fuse_channel_split:
            return triplane         # un-split: every consumer of the space cache in this package accepts [B,6,2C,H,W]
        B, _, C2, H, W = triplane.shape
        C_ = C2 // 2
        return torch.cat([triplane[:, 0:3, :C_], triplane[:, 3:6, C_:]], dim=1).contiguous()

    def generate_space_cache(self, styles: Optional[Tensor] = None, text_embed: Optional[Tensor] = None):
        """few_step…diffusion.py:156-165: ``space_generator(text_embed=, styles=)``.  The generator is outside this
        package; with none attached this raises."""
        if self.space_generator is None:
            raise RuntimeError("no space_generator attached: pass space_cache= to the renderer "
                               "(the SD generator is outside this package)")
        return self.space_generator(text_embed=text_embed, styles=styles)

    # ------------------------------------------------------------------ the plugin surface
    def rescale_points(self, points: Tensor) -> Tensor:
        lo, hi = self.bbox[0], self.bbox[1]
        return (points - lo) / (hi - lo) * 2.0 - 1.0

    def interpolate_encodings(self, points: Tensor, space_cache: Tensor, only_geo: bool = False):
        """few_step…diffusion.py:198-258.  points [B,N,3] ALREADY rescaled to [-1,1]^3 (as the reference calls it);
        returns the geometry encoding [B,N,C] (planes summed, ``v1``) and the texture encoding [B,N,3C] (planes
        concatenated, ``v2``).  The reference rotates the cache into a zero tensor and makes two contiguous copies per
        call; here the rotation is the one-time repack and the two samplers are ``tt_sample_planes_fwd`` launches.
        Differentiable w.r.t. ``space_cache`` and ``points`` (to second order in ``points``, like the reference's
        grid_sample_gradfix op)."""
        from . import sampler
        B = points.shape[0]
        sc = self._check_cache(space_cache, B)
        if sc.requires_grad and torch.is_grad_enabled():
            planes = ops.RepackFunction.apply(sc, self.plane_channels, *ops.split_offsets(sc.shape[2], self.plane_channels))
        else:
            planes = ops.cached_planes(sc, self.plane_channels)
        _, _, R, _, C_ = planes.shape
        grid = sampler.project_onto_planes(sampler.PLANES, points.reshape(B, -1, 3)).float().contiguous()   # [3B,N,2]
        geo = sampler.sample_planes(planes[:, 0:3].reshape(B * 3, R, R, C_), grid, 3, False)
        geo = geo.view(*points.shape[:-1], -1)
        if only_geo:
            return geo
        tex = sampler.sample_planes(planes[:, 3:6].reshape(B * 3, R, R, C_), grid, 3, True)
        return geo, tex.view(*points.shape[:-1], -1)

    def forward(self, points: Tensor, space_cache: Tensor, output_normal: bool = False) -> Dict[str, Tensor]:
        """few_step…diffusion.py:273-351.  points [B,N,3] world coordinates."""
        B = points.shape[0]
        sc = self._check_cache(space_cache, B)
        grad = torch.is_grad_enabled()
        pts = points.detach().reshape(B, -1, 3)
        sdf, sdf_orig, features, normal, sdf_grad = ops.GeometryFunction.apply(
            sc, *self.decoder_weights(), pts, self.path_scalars(), bool(output_normal))
        out = {"sdf": sdf, "sdf_orig": sdf_orig, "features": features}
        if output_normal:
            out.update({"normal": normal, "shading_normal": normal, "sdf_grad": sdf_grad})
        if not grad:
            out = {k: v.detach() for k, v in out.items()}
        return out

    def forward_sdf(self, points: Tensor, space_cache: Tensor) -> Tensor:
        """few_step…diffusion.py:353-373."""
        sdf, _ = self._field(points, space_cache, with_deformation=False)
        return sdf.view(*points.shape[:-1], 1)

    def forward_field(self, points: Tensor, space_cache: Tensor) -> Tuple[Tensor, Optional[Tensor]]:
        """few_step…diffusion.py:375-394: sdf [B,M,1] and deformation [B,M,3] (None without a deformable grid)."""
        sdf, deform = self._field(points, space_cache, with_deformation=self.cfg.isosurface_deformable_grid)
        sdf = sdf.view(*points.shape[:-1], 1)
        if deform is not None:
            deform = deform.view(*points.shape[:-1], 3)
        return sdf, deform

    def forward_field_grid(self, resolution: int, space_cache: Tensor) -> Tuple[Tensor, Optional[Tensor]]:
        """Same as ``forward_field`` on the isosurface helper's vertex grid (threestudio/models/isosurface.py:37-51)
        without materialising the points: the kernel generates vertex (ix*res+iy)*res+iz in place."""
        planes = ops.cached_planes(space_cache, self.plane_channels)
        wpack = ops.cached_wpack(self.sdf_network.weights(), self.feature_network.weights(),
                                 self._deformation_weights(), self.plane_channels)
        want = ["sdf"] + (["deformation"] if self.cfg.isosurface_deformable_grid else [])
        out = ops.geometry_fwd(planes, wpack, self.path_scalars(), None, int(resolution), want)
        P = space_cache.shape[0]
        d = out.get("deformation")
        return out["sdf"].view(P, -1, 1), None if d is None else d.view(P, -1, 3)

    def _field(self, points, space_cache, with_deformation):
        B = points.shape[0]
        sc = self._check_cache(space_cache, B)
        pts = points.detach().reshape(B, -1, 3)
        dw = self._deformation_weights() if with_deformation else None
        need_grad = torch.is_grad_enabled() and (sc.requires_grad or any(w.requires_grad for w in self.decoder_weights())
                                                 or any(w.requires_grad for w in (dw or [])))
        if need_grad and not with_deformation:
            sdf = ops.GeometryFunction.apply(sc, *self.decoder_weights(), pts, self.path_scalars(), False)[0]
            return sdf, None
        if need_grad:       # mesh renderer: gradients reach the planes, the SDF decoder and the deformation decoder
            return ops.FieldFunction.apply(sc, *self.decoder_weights(), *dw, pts, self.path_scalars())
        planes = ops.cached_planes(sc, self.plane_channels)
        wpack = ops.cached_wpack(self.sdf_network.weights(), self.feature_network.weights(),
                                 self._deformation_weights(), self.plane_channels)
        want = ["sdf"] + (["deformation"] if with_deformation else [])
        out = ops.geometry_fwd(planes, wpack, self.path_scalars(), pts, 0, want)
        return out["sdf"], out.get("deformation")

    def forward_level(self, field: Tensor, threshold: float) -> Tensor:
        return field - threshold

    def export(self, points: Tensor, space_cache: Tensor, **kwargs) -> Dict[str, Any]:
        """few_step…diffusion.py:402-430: features at surface points (vertex colours)."""
        orig = points.shape
        sc = self._check_cache(space_cache, 1)
        planes = ops.cached_planes(sc, self.plane_channels)
        wpack = ops.cached_wpack(self.sdf_network.weights(), self.feature_network.weights(),
                                 self._deformation_weights(), self.plane_channels)
        out = ops.geometry_fwd(planes, wpack, self.path_scalars(), points.detach().reshape(1, -1, 3), 0, ["features"])
        return {"features": out["features"].view(*orig[:-1], 3)}
