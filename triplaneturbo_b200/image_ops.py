"""Image-space tail of the renderer: Nr-sized element-wise maps applied to the per-ray accumulators that
``tt_render_fwd`` produces.  Plain differentiable torch ops (a few dozen bytes per ray; the marched samples
never reach this code).

Follows custom/triplaneturbo/models/renderers/generative_space_sdf_volume_renderer.py:433-530 of the reference.
"""
from typing import Dict, Optional

import torch
import torch.nn.functional as F
from torch import Tensor


def compose_images(acc: Tensor, bg_color: Tensor, camera_distances: Optional[Tensor], c2w: Optional[Tensor],
                   B: int, H: int, W: int, normal_direction: str = "camera", views_per_cache: int = 1
                   ) -> Dict[str, Tensor]:
    """acc [Nr,10] = opacity, depth, rgb(3), z_variance, normal_sum(3), eikonal_sum -> the renderer's image
    dictionary (plus ``eikonal_sum`` [B,H,W,1]: Σ_samples (|sdf_grad|-1)^2 per ray, an extension that lets the
    eikonal loss be taken without per-sample tensors)."""
    opacity, depth = acc[:, 0:1], acc[:, 1:2]
    comp_rgb_fg, z_variance, nsum = acc[:, 2:5], acc[:, 5:6], acc[:, 6:9]
    comp_rgb_bg = bg_color
    if bg_color.shape[:-1] == (B, H, W):                                   # :436-437
        bg_color = bg_color.reshape(B * H * W, -1)
    comp_rgb = comp_rgb_fg + bg_color * (1.0 - opacity)                    # :439
    out = {
        "comp_rgb": comp_rgb.view(B, H, W, -1),
        "comp_rgb_fg": comp_rgb_fg.reshape(B, H, W, -1),
        "comp_rgb_bg": comp_rgb_bg,
        "opacity": opacity.reshape(B, H, W, 1),
        "depth": depth.reshape(B, H, W, 1),
        "z_variance": z_variance.reshape(B, H, W, 1),
        "eikonal_sum": acc[:, 9:10].reshape(B, H, W, 1),
    }
    if camera_distances is not None:                                       # :451-462, RichDreamer disparity
        sqrt3 = torch.sqrt(3 * torch.ones(1, 1, 1, 1, device=acc.device, dtype=acc.dtype))
        far = camera_distances.reshape(-1, 1, 1, 1) + sqrt3
        near = camera_distances.reshape(-1, 1, 1, 1) - sqrt3
        disparity_tmp = out["depth"] * out["opacity"] + (1.0 - out["opacity"]) * far
        out["disparity"] = torch.clamp((far - disparity_tmp) / (far - near), 0.0, 1.0).view(B, H, W, 1)
    comp_normal = F.normalize(nsum, dim=-1)                                # :467-473
    out["comp_normal"] = comp_normal.view(B, H, W, 3)
    if normal_direction == "camera" and c2w is not None:                   # :475-511
        bg_normal = 0.5 * torch.ones_like(comp_normal)
        bg_normal[:, 2] = 1.0
        bg_normal_white = torch.ones_like(comp_normal)
        rot = torch.inverse(c2w)[:, :3, :3]
        comp_normal_cam = comp_normal.view(B, -1, 3) @ rot.permute(0, 2, 1)
        flip_x = torch.eye(3, device=acc.device, dtype=acc.dtype)
        flip_x[0, 0] = -1
        comp_normal_cam = (comp_normal_cam @ flip_x[None, :, :]).view(-1, 3)
        out["comp_normal_cam_vis"] = ((comp_normal_cam + 1.0) / 2.0 * opacity
                                      + (1 - opacity) * bg_normal).view(B, H, W, 3)
        out["comp_normal_cam_vis_white"] = ((comp_normal_cam + 1.0) / 2.0 * opacity
                                            + (1 - opacity) * bg_normal_white).view(B, H, W, 3)
    elif normal_direction == "front" and c2w is not None:                  # :512-527
        V = views_per_cache
        bg_normal_white = torch.ones_like(comp_normal)
        c2w_front = c2w[0::V].repeat_interleave(V, dim=0)
        rot = torch.inverse(c2w_front)[:, :3, :3]
        comp_normal_front = (comp_normal.view(B, -1, 3) @ rot.permute(0, 2, 1)).view(-1, 3)
        out["comp_normal_cam_vis_white"] = ((comp_normal_front + 1.0) / 2.0 * opacity
                                            + (1 - opacity) * bg_normal_white).view(B, H, W, 3)
    return out
