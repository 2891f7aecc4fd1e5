"""Stand-alone plane sampler: the reference's functional API on the ``tt_sample_planes_*`` kernels.

    grid_sample_2d(input, grid, padding_mode, align_corners)   custom/triplaneturbo/extern/grid_sample_gradfix/cuda_gridsample.py:22-24
    grid_sample(input, grid)                                   custom/triplaneturbo/models/geometry/utils.py:21-24
    sample_from_planes(plane_features, coordinates, ...)       custom/triplaneturbo/models/geometry/utils.py:127-161
    project_onto_planes(planes, coordinates)                   custom/triplaneturbo/models/geometry/utils.py:111-125

The autograd structure mirrors the reference's (forward -> backward -> backward-of-backward,
``cuda_gridsample.py:31-79``), each stage one kernel (``csrc/tt_sampler.cuh``); the second derivative is the function of
``gridsample_cuda.cu:87-209``.  Inside the renderer the same arithmetic is fused into the decoder kernels; this module is
for callers that use the operator on its own.  Only the path's mode exists: bilinear, zeros padding,
``align_corners=False``.  CUDA fp32 tensors only (``_cabi.TTError`` otherwise): there is no CPU path.
"""
import ctypes as C
from typing import Optional

import torch
from torch import Tensor

from . import _cabi
from .ops import _lib, _need, _ptr, _stream

# plane axes of geometry/utils.py:46-63 (eg3d issue 67 fix); projection = coordinates @ inv(axes), first two columns
PLANES = torch.tensor([[[1, 0, 0], [0, 1, 0], [0, 0, 1]],
                       [[1, 0, 0], [0, 0, 1], [0, 1, 0]],
                       [[0, 0, 1], [0, 1, 0], [1, 0, 0]]], dtype=torch.float32)


def _call(name, *args):
    L = _lib()
    _cabi.check(L, getattr(L, name)(*args), name)


class _ToChannelLast(torch.autograd.Function):
    """[B, C, HW] -> [B, HW, C] (tt_to_channel_last); a permutation, so every derivative is the other transpose."""

    @staticmethod
    def forward(ctx, x):
        x = _need(x, "input")
        B, C_, HW = x.shape
        y = torch.empty((B, HW, C_), device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _call("tt_to_channel_last", _ptr(x), B, C_, HW, _ptr(y), _stream(x.device))
        return y

    @staticmethod
    def backward(ctx, g):
        return _FromChannelLast.apply(g)


class _FromChannelLast(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y):
        y = _need(y, "input")
        B, HW, C_ = y.shape
        x = torch.empty((B, C_, HW), device=y.device, dtype=torch.float32)
        with torch.cuda.device(y.device):
            _call("tt_from_channel_last", _ptr(y), B, C_, HW, _ptr(x), _stream(y.device))
        return x

    @staticmethod
    def backward(ctx, g):
        return _ToChannelLast.apply(g)


def _dims(planes, grid, K):
    NK, H, W, C_ = planes.shape
    M = grid.shape[1]
    if NK % K or grid.shape[0] != NK or grid.shape[2] != 2:
        raise _cabi.TTError("sample_planes: planes [N*K,H,W,C] and grid [N*K,M,2] do not match")
    return NK // K, H, W, C_, M


class _SampleForward(torch.autograd.Function):
    """planes [N*K,H,W,C] channel-last, grid [N*K,M,2] -> [N,M,C] (sum over K) or [N,M,K*C] (concat)."""

    @staticmethod
    def forward(ctx, planes, grid, K, concat):
        planes, grid = _need(planes, "planes"), _need(grid, "grid")
        N, H, W, C_, M = _dims(planes, grid, K)
        out = torch.empty((N, M, K * C_ if concat else C_), device=planes.device, dtype=torch.float32)
        with torch.cuda.device(planes.device):
            _call("tt_sample_planes_fwd", _ptr(planes), N, K, C_, H, W, _ptr(grid), M, int(concat), _ptr(out),
                  _stream(planes.device))
        ctx.save_for_backward(planes, grid)
        ctx.K, ctx.concat = K, concat
        return out

    @staticmethod
    def backward(ctx, g_out):
        planes, grid = ctx.saved_tensors
        g_planes, g_grid = _SampleBackward.apply(g_out, planes, grid, ctx.K, ctx.concat,
                                                 ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return g_planes, g_grid, None, None


class _SampleBackward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, g_out, planes, grid, K, concat, need_planes, need_grid):
        g_out = _need(g_out, "grad_output")
        N, H, W, C_, M = _dims(planes, grid, K)
        g_planes = torch.zeros_like(planes) if need_planes else None
        g_grid = torch.empty_like(grid) if need_grid else None
        with torch.cuda.device(planes.device):
            _call("tt_sample_planes_bwd", _ptr(planes), N, K, C_, H, W, _ptr(grid), M, int(concat), _ptr(g_out),
                  _ptr(g_planes), _ptr(g_grid), _stream(planes.device))
        ctx.save_for_backward(g_out, planes, grid)
        ctx.K, ctx.concat = K, concat
        return g_planes, g_grid

    @staticmethod
    def backward(ctx, gg_planes, gg_grid):
        g_out, planes, grid = ctx.saved_tensors
        K, concat = ctx.K, ctx.concat
        N, H, W, C_, M = _dims(planes, grid, K)
        need_go, need_planes, need_grid = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        gg_planes = None if gg_planes is None else _need(gg_planes, "grad2_grad_input")
        gg_grid = None if gg_grid is None else _need(gg_grid, "grad2_grad_grid")
        gg_out = torch.empty_like(g_out) if need_go else None
        # d/d planes of the backward comes only through d/d grid (it is linear in planes with weights ẇ)
        g_planes = torch.zeros_like(planes) if (need_planes and gg_grid is not None) else None
        g_grid = torch.empty_like(grid) if need_grid else None
        with torch.cuda.device(planes.device):
            _call("tt_sample_planes_bwdbwd", _ptr(planes), N, K, C_, H, W, _ptr(grid), M, int(concat), _ptr(g_out),
                  _ptr(gg_planes), _ptr(gg_grid), _ptr(gg_out), _ptr(g_planes), _ptr(g_grid), _stream(planes.device))
        return gg_out, g_planes, g_grid, None, None, None, None


def sample_planes(planes_cl: Tensor, grid: Tensor, n_planes: int = 1, concat: bool = False) -> Tensor:
    """The operator on its native layout: channel-last planes [N*K,H,W,C], grid [N*K,M,2] -> [N,M,C] / [N,M,K*C]."""
    return _SampleForward.apply(planes_cl, grid, int(n_planes), bool(concat))


def to_channel_last(x: Tensor) -> Tensor:
    """[B,C,H,W] -> [B,H,W,C] through the library's transpose kernel (differentiable to any order)."""
    B, C_, H, W = x.shape
    return _ToChannelLast.apply(x.reshape(B, C_, H * W)).view(B, H, W, C_)


def grid_sample_2d(input: Tensor, grid: Tensor, padding_mode: str = "zeros", align_corners: bool = False) -> Tensor:
    """input [N,C,H,W], grid [N,Ho,Wo,2] -> [N,C,Ho,Wo] (a permuted view of the point-major result)."""
    if padding_mode != "zeros" or align_corners:
        raise NotImplementedError("the path samples with padding_mode='zeros', align_corners=False "
                                  "(custom/triplaneturbo/models/geometry/utils.py:21-24)")
    if input.dim() != 4 or grid.dim() != 4 or input.shape[0] != grid.shape[0] or grid.shape[3] != 2:
        raise _cabi.TTError("grid_sample_2d: expected input [N,C,H,W] and grid [N,Ho,Wo,2]")
    N, C_, H, W = input.shape
    Ho, Wo = grid.shape[1], grid.shape[2]
    out = sample_planes(to_channel_last(input), grid.reshape(N, Ho * Wo, 2), 1, False)       # [N, Ho*Wo, C]
    return out.permute(0, 2, 1).unflatten(2, (Ho, Wo))


def grid_sample(input: Tensor, grid: Tensor) -> Tensor:
    """geometry/utils.py:21-24 (the reference switches to ATen when grid needs no gradient: same values)."""
    return grid_sample_2d(input, grid, padding_mode="zeros", align_corners=False)


def project_onto_planes(planes: Tensor, coordinates: Tensor) -> Tensor:
    """geometry/utils.py:111-125: [N,M,3] -> [N*n_planes,M,2].  The plane axes are permutation matrices, so the
    projection is a column selection of the coordinates (bit-identical to the reference's bmm with the inverse)."""
    N, M, _ = coordinates.shape
    inv = torch.linalg.inv(planes.to(torch.float32).cpu())
    cols = []
    for k in range(planes.shape[0]):
        sel = inv[k][:, :2]                                   # [3, 2]
        if not bool(((sel == 0) | (sel == 1)).all()) or not bool((sel.sum(0) == 1).all()):
            return torch.bmm(coordinates.unsqueeze(1).expand(-1, planes.shape[0], -1, -1).reshape(-1, M, 3),
                             inv.to(coordinates).unsqueeze(0).expand(N, -1, -1, -1).reshape(-1, 3, 3))[..., :2]
        cols.append([int(sel[:, 0].argmax()), int(sel[:, 1].argmax())])
    idx = torch.tensor(cols, device=coordinates.device)       # [n_planes, 2]
    return coordinates[:, :, idx].permute(0, 2, 1, 3).reshape(N * planes.shape[0], M, 2)


def sample_from_planes(plane_features: Tensor, coordinates: Tensor, mode: str = "bilinear", padding_mode: str = "zeros",
                       box_warp: float = 2, interpolate_feat: Optional[str] = "None") -> Tensor:
    """geometry/utils.py:127-161.  plane_features [N,n_planes,C,H,W], coordinates [N,M,3] ->
    [N,M,C] (None / "v1": planes summed), [N,M,n_planes*C] ("v2"), [N,M,C-1] ("v3": last channel gates), [N,M,C] ("v4")."""
    assert padding_mode == "zeros" and mode == "bilinear"
    N, n_planes, C_, H, W = plane_features.shape
    _, M, _ = coordinates.shape
    feats = plane_features.reshape(N * n_planes, C_, H, W)
    coordinates = (2 / box_warp) * coordinates
    grid = project_onto_planes(PLANES, coordinates).float()
    if interpolate_feat in (None, "None", "v1"):
        concat = False
    elif interpolate_feat == "v2":
        concat = True
    elif interpolate_feat == "v3":
        feats, concat = torch.sigmoid(feats[:, -1:, ...]) * feats[:, :-1, ...], False
    elif interpolate_feat == "v4":
        feats, concat = torch.tanh(feats), False
    else:
        raise NotImplementedError(interpolate_feat)
    if feats.shape[1] % 4:
        raise _cabi.TTError("sample_from_planes: the channel count must be a multiple of 4")
    return sample_planes(to_channel_last(feats), grid, n_planes, concat)
