"""Synthetic inputs of the path: random cameras/rays drawn like the reference's data module
(custom/triplaneturbo/data/…multistep_v2.py:250-337,841-874; ray helpers threestudio/utils/ops.py:194-290),
random triplanes and decoder weights.  Used by bench.py, smoke() and the tests; deterministic per seed."""
import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F


def camera_rays(B: int, H: int, W: int, seed: int = 2, fovy_deg: float = 60.0, views_per_prompt: int = None
                ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """B cameras looking at the origin (up +z): elevation U[0,30] deg, azimuths spread over the views of a prompt,
    distance U[0.8,1.0]/tan(fovy/2).  Returns rays_o, rays_d [B,H,W,3] (normalised), c2w [B,4,4], distances [B]."""
    g = torch.Generator().manual_seed(seed)
    V = views_per_prompt or B
    elev = torch.rand(B, generator=g) * 30.0
    k = torch.arange(B) % V
    azim = (torch.rand(B, generator=g) + k) / V * 360.0 - 180.0
    fovy = torch.full((B,), fovy_deg) * math.pi / 180
    dist = (torch.rand(B, generator=g) * 0.2 + 0.8) / torch.tan(0.5 * fovy)
    el, az = elev * math.pi / 180, azim * math.pi / 180
    pos = torch.stack([dist * torch.cos(el) * torch.cos(az), dist * torch.cos(el) * torch.sin(az),
                       dist * torch.sin(el)], -1).float()
    up = torch.tensor([0.0, 0.0, 1.0])[None].repeat(B, 1)
    lookat = F.normalize(-pos, dim=-1)
    right = F.normalize(torch.linalg.cross(lookat, up), dim=-1)
    up = F.normalize(torch.linalg.cross(right, lookat), dim=-1)
    c2w = torch.zeros(B, 4, 4)
    c2w[:, :3, :3] = torch.stack([right, up, -lookat], dim=-1)
    c2w[:, :3, 3] = pos
    c2w[:, 3, 3] = 1.0
    i, j = torch.meshgrid(torch.arange(W, dtype=torch.float32) + 0.5, torch.arange(H, dtype=torch.float32) + 0.5,
                          indexing="xy")
    dirs = torch.stack([(i - W / 2), -(j - H / 2), -torch.ones_like(i)], -1)[None].repeat(B, 1, 1, 1)
    focal = 0.5 * H / torch.tan(0.5 * fovy)
    dirs[..., :2] = dirs[..., :2] / focal[:, None, None, None]
    rays_d = (dirs[:, :, :, None, :] * c2w[:, None, None, :3, :3]).sum(-1)
    rays_o = c2w[:, None, None, :3, 3].expand(rays_d.shape)
    rays_d = F.normalize(rays_d, dim=-1)
    return rays_o.contiguous(), rays_d.contiguous(), c2w, dist


def random_triplanes(P: int, C: int, R: int, seed: int = 0, device="cpu") -> torch.Tensor:
    """randn * 0.5 (VAE-like scale), NCHW [P,6,C,R,R]."""
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(P, 6, C, R, R, generator=g) * 0.5).to(device)


def random_decoder(C: int, seed: int = 1, gain: float = 1.5) -> Dict[str, torch.Tensor]:
    """nn.Linear default init (kaiming_uniform, a=sqrt(5)) times ``gain`` so the SDF crosses zero inside the volume;
    keys ``w_{sdf,feature,deformation}_{0,1,2}``."""
    g = torch.Generator().manual_seed(seed)

    def lin(o, i):
        bound = 1.0 / math.sqrt(i)
        return (torch.rand(o, i, generator=g) * 2 - 1) * bound * gain
    out = {}
    for name, din, dout in (("sdf", C, 1), ("feature", 3 * C, 3), ("deformation", C, 3)):
        out[f"w_{name}_0"], out[f"w_{name}_1"], out[f"w_{name}_2"] = lin(64, din), lin(64, 64), lin(dout, 64)
    return out


def build_plugins(fx, device, n_samples=64, n_imp=128, normal_direction="camera", rgb_grad_shrink=1.0, **renderer_over):
    """Geometry + renderer plugins, created by registry name like the reference's systems do
    (threestudio/systems/base.py:292-303), holding the weights of ``fx`` (keys ``w_{sdf,feature,deformation}_{0,1,2}``,
    ``space_cache`` only for its channel count).  Shipped configuration: configs/TriplaneTurbo_v1.yaml:73-150."""
    import triplaneturbo_b200 as tt
    C_ = fx["space_cache"].shape[2]
    geom = tt.find("few-step-triplane-dual-stable-diffusion")(dict(
        radius=1.0, normal_type="analytic", sdf_bias="sphere", sdf_bias_params=0.5, rotate_planes="v1",
        split_channels="v1", geo_interpolate="v1", tex_interpolate="v2",
        space_generator_config={"output_dim": 2 * C_}, isosurface_deformable_grid=True)).to(device)
    sd = {}
    for name in ("sdf", "feature", "deformation"):
        for i, idx in enumerate((0, 2, 4)):
            if f"w_{name}_{i}" in fx:
                sd[f"{name}_network.layers.{idx}.weight"] = fx[f"w_{name}_{i}"]
    missing = geom.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys
    material = tt.find("no-material")(dict(n_output_dims=3, color_activation="sigmoid-mipnerf", requires_normal=True))
    background = tt.find("solid-color-background")({}).to(device)
    cfg = dict(radius=1.0, use_volsdf=False, trainable_variance=False, learned_variance_init=0.4605,
               rgb_grad_shrink=rgb_grad_shrink, estimator="importance", num_samples_per_ray=n_samples,
               num_samples_per_ray_importance=n_imp, near_plane=0.1, far_plane=4.0, train_chunk_size=0, randomized=False,
               normal_direction=normal_direction, eval_chunk_size=500)
    cfg.update(renderer_over)
    rend = tt.find("generative-space-sdf-volume-renderer")(cfg, geometry=geom, material=material,
                                                           background=background).to(device)
    rend.update_step(0, 0)
    return geom, rend
