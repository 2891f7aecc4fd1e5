"""The two nerfacc v0.5.2 entry points the renderer calls, on tt_composite_* kernels
(custom/triplaneturbo/models/renderers/generative_space_sdf_volume_renderer.py:408-431,467).

Only the path's dense sample layout is supported: ``ray_indices`` is ``arange(n_rays)`` repeated S times
(…sdf_volume_renderer.py:317-322), so packed == ``[n_rays, S]``.
"""
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import ops


def _dense(n: int, n_rays: int, ray_indices: Optional[Tensor]) -> int:
    if n_rays <= 0 or n % n_rays != 0:
        raise ValueError("samples must be dense: len(alphas) == n_rays * S")
    return n // n_rays


def render_weight_from_alpha(alphas: Tensor, packed_info=None, ray_indices: Optional[Tensor] = None,
                             n_rays: Optional[int] = None, prefix_trans=None) -> Tuple[Tensor, Tensor]:
    if packed_info is not None or prefix_trans is not None:
        raise NotImplementedError("only dense ray_indices are on the path")
    if alphas.dim() == 2:
        n_rays, S = alphas.shape
    else:
        S = _dense(alphas.numel(), int(n_rays), ray_indices)
    w, T, _ = ops.CompositeFunction.apply(alphas.reshape(n_rays, S), None)
    return w.reshape(alphas.shape), T.reshape(alphas.shape)


def accumulate_along_rays(weights: Tensor, values: Optional[Tensor] = None, ray_indices: Optional[Tensor] = None,
                          n_rays: Optional[int] = None) -> Tensor:
    """Σ_samples weights * values per ray.  Differentiable w.r.t. weights and values (torch reduction over the
    dense [n_rays, S] view; the fused renderer never calls this)."""
    S = _dense(weights.numel(), int(n_rays), ray_indices)
    w = weights.reshape(n_rays, S, 1)
    if values is None:
        return w.sum(1)
    return (w * values.reshape(n_rays, S, -1)).sum(1)
