"""Restatement of the nerfacc v0.5.2 entry points the path calls.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  **Parity unpinned**: nerfacc
(``requirements.txt:5`` of the reference, pinned
``git+https://github.com/KAIR-BAIR/nerfacc.git@v0.5.2``) is an un-vendored
dependency, its source is not in this image and the reference holds no tests
or golden vectors at this boundary.  What follows restates the documented
v0.5.x Python API.

Call sites in the reference:
  * ``threestudio/models/estimators.py:9-12,72-90``  RayIntervals,
    importance_sampling, render_transmittance_from_density
  * ``custom/triplaneturbo/models/renderers/generative_space_sdf_volume_renderer.py:408-431,467``
    render_weight_from_alpha, accumulate_along_rays

Fixed conventions (the part nerfacc's docs do not pin down):
  * ``importance_sampling`` draws ``n + 1`` interval *edges* at quantiles
    ``u_j = j / n`` (``stratified=False``) or ``u_j = (j + b) / (n + 1)`` with
    one jitter ``b in [0, 1)`` per ray (``stratified=True``; nerfacc takes ``b``
    from its own Philox stream, here it is an explicit input), and inverts the
    piece-wise-linear CDF ``(vals_k, cdf_k)`` with a right-sided search
    (``cdf[p-1] <= u < cdf[p]``, ``p`` clamped to ``[1, n_in-1]``) and linear
    interpolation ``v = v0 + clamp((u - c0) / (c1 - c0), 0, 1) * (v1 - v0)``
    (fraction 0 when ``c1 == c0``).
  With these, level 0 of ``ImportanceEstimator.sampling`` (cdf ``[0, 1]``)
  yields the uniform edges ``j / n`` exactly.
Every arithmetic step is a separate fp32 torch op (no fused multiply-add), so
the CUDA kernel can reproduce the result bit for bit.
"""
from typing import Optional, Tuple

import torch
from torch import Tensor


def exclusive_sum(x: Tensor) -> Tensor:
    """Dense ``[n_rays, S]`` exclusive prefix sum along the last dim."""
    return torch.cumsum(torch.cat([torch.zeros_like(x[..., :1]), x[..., :-1]], dim=-1), dim=-1)


def exclusive_prod(x: Tensor) -> Tensor:
    """Dense ``[n_rays, S]`` exclusive prefix product along the last dim."""
    return torch.cumprod(torch.cat([torch.ones_like(x[..., :1]), x[..., :-1]], dim=-1), dim=-1)


def quantiles(n: int, n_rays: int, stratified: bool, jitter: Optional[Tensor], like: Tensor) -> Tensor:
    j = torch.arange(n + 1, dtype=like.dtype, device=like.device)
    if not stratified:
        return (j / float(n))[None, :].expand(n_rays, -1)
    assert jitter is not None and jitter.shape == (n_rays,), "stratified sampling needs one jitter per ray"
    return (j[None, :] + jitter[:, None].to(like.dtype)) / float(n + 1)


def importance_sampling(
    vals: Tensor, cdfs: Tensor, n_intervals: int, stratified: bool = False, jitter: Optional[Tensor] = None
) -> Tensor:
    """``nerfacc.pdf.importance_sampling`` for batched rays.

    vals, cdfs: [n_rays, n_in] (edges and CDF at the edges) → [n_rays, n_intervals + 1] new edges.
    """
    n_rays, n_in = cdfs.shape
    u = quantiles(n_intervals, n_rays, stratified, jitter, cdfs).contiguous()
    p = torch.searchsorted(cdfs.contiguous(), u, right=True).clamp(1, n_in - 1)
    c0 = torch.gather(cdfs, 1, p - 1)
    c1 = torch.gather(cdfs, 1, p)
    v0 = torch.gather(vals, 1, p - 1)
    v1 = torch.gather(vals, 1, p)
    denom = c1 - c0
    frac = torch.where(denom > 0, (u - c0) / denom, torch.zeros_like(u)).clamp(0.0, 1.0)
    return v0 + frac * (v1 - v0)


def render_transmittance_from_density(t_starts: Tensor, t_ends: Tensor, sigmas: Tensor) -> Tuple[Tensor, Tensor]:
    """Dense ``[n_rays, S]``: ``T = exp(-exclusive_sum(sigma * dt))``, ``alpha = 1 - exp(-sigma * dt)``."""
    sigmas_dt = sigmas * (t_ends - t_starts)
    alphas = 1.0 - torch.exp(-sigmas_dt)
    trans = torch.exp(-exclusive_sum(sigmas_dt))
    return trans, alphas


def _dense_view(x: Tensor, ray_indices: Tensor, n_rays: int):
    """The path's ray_indices are ``arange(n_rays)`` repeated S times (reference
    ``…sdf_volume_renderer.py:317-322``): packed ≡ dense ``[n_rays, S]``."""
    n = ray_indices.numel()
    if n_rays > 0 and n % n_rays == 0:
        S = n // n_rays
        expect = torch.arange(n_rays, device=ray_indices.device).repeat_interleave(S)
        if torch.equal(ray_indices.long(), expect):
            return x.reshape(n_rays, S)
    return None


def render_weight_from_alpha(alphas: Tensor, ray_indices: Tensor, n_rays: int) -> Tuple[Tensor, Tensor]:
    """``w_i = alpha_i * prod_{j<i, same ray} (1 - alpha_j)`` on flattened samples."""
    dense = _dense_view(alphas, ray_indices, n_rays)
    if dense is not None:
        trans = exclusive_prod(1.0 - dense).reshape(-1)
    else:  # general packed case: samples sorted by ray, ragged counts
        trans = torch.ones_like(alphas)
        start = 0
        _, counts = torch.unique_consecutive(ray_indices, return_counts=True)
        pieces = []
        for c in counts.tolist():
            pieces.append(exclusive_prod(1.0 - alphas[start:start + c][None])[0])
            start += c
        trans = torch.cat(pieces) if pieces else trans
    return trans * alphas, trans


def accumulate_along_rays(weights: Tensor, values: Optional[Tensor], ray_indices: Tensor, n_rays: int) -> Tensor:
    """``out[n_rays, D].index_add_(0, ray_indices, weights[:, None] * values)``."""
    src = weights[:, None] if values is None else weights[:, None] * values
    out = torch.zeros((n_rays, src.shape[-1]), dtype=src.dtype, device=src.device)
    return out.index_add(0, ray_indices.long(), src)
