"""Gather-based bilinear ``grid_sample`` that autograd can differentiate twice.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

Restates ATen ``grid_sampler_2d`` for ``mode='bilinear'``,
``padding_mode='zeros'``, ``align_corners=False`` — the only mode the path uses
(``custom/triplaneturbo/models/geometry/utils.py:21-24``) — with the operation
order of the CUDA kernel the reference runs on
(``aten/src/ATen/native/cuda/GridSampler.cuh``: unnormalise
``((g + 1) * size - 1) / 2``, corners from ``floor``, weights
``nw = (ix_se - ix) * (iy_se - iy)`` …, accumulation order nw, ne, sw, se).

Because every step is an ordinary differentiable torch op, double backward
works on any device.  That replaces, for the oracle, the reference's
``grid_sample_gradfix`` CUDA extension
(``custom/triplaneturbo/extern/grid_sample_gradfix/cuda_gridsample.py:31-79``,
``gridsample_cuda.cu:27-210``), which exists only because ATen has no
derivative for ``grid_sampler_2d_backward``.
"""
import torch


def unnormalize(coord: torch.Tensor, size: int) -> torch.Tensor:
    """ATen ``grid_sampler_unnormalize`` with ``align_corners=False``."""
    return ((coord + 1.0) * size - 1.0) / 2.0


def corner_indices(grid: torch.Tensor, H: int, W: int):
    """Integer corner indices (ix_nw, iy_nw) of every sample; bit-exact contract.

    Returns int64 tensors shaped like ``grid[..., 0]``.
    """
    ix = unnormalize(grid[..., 0], W)
    iy = unnormalize(grid[..., 1], H)
    return torch.floor(ix).long(), torch.floor(iy).long()


def grid_sample_2d_manual(input: torch.Tensor, grid: torch.Tensor) -> torch.Tensor:
    """input [N,C,H,W], grid [N,Ho,Wo,2] (x→W, y→H) → [N,C,Ho,Wo]."""
    N, C, H, W = input.shape
    Ho, Wo = grid.shape[1], grid.shape[2]
    gx = grid[..., 0].reshape(N, -1)
    gy = grid[..., 1].reshape(N, -1)
    ix = unnormalize(gx, W)
    iy = unnormalize(gy, H)
    ix_nw = torch.floor(ix)
    iy_nw = torch.floor(iy)
    ix_ne, iy_ne = ix_nw + 1, iy_nw
    ix_sw, iy_sw = ix_nw, iy_nw + 1
    ix_se, iy_se = ix_nw + 1, iy_nw + 1
    nw = (ix_se - ix) * (iy_se - iy)
    ne = (ix - ix_sw) * (iy_sw - iy)
    sw = (ix_ne - ix) * (iy - iy_ne)
    se = (ix - ix_nw) * (iy - iy_nw)

    flat = input.reshape(N, C, H * W)

    def tap(xi, yi, w):
        inb = (xi >= 0) & (xi <= W - 1) & (yi >= 0) & (yi <= H - 1)
        idx = (yi.clamp(0, H - 1) * W + xi.clamp(0, W - 1)).long()
        vals = torch.gather(flat, 2, idx[:, None, :].expand(-1, C, -1))
        return vals * (w * inb.to(w.dtype))[:, None, :]

    out = tap(ix_nw, iy_nw, nw)
    out = out + tap(ix_ne, iy_ne, ne)
    out = out + tap(ix_sw, iy_sw, sw)
    out = out + tap(ix_se, iy_se, se)
    return out.reshape(N, C, Ho, Wo)
