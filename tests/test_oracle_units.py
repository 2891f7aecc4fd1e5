"""Known-answer tests for the oracle's building blocks (CPU)."""
import torch
import torch.nn.functional as F

from oracle import nerfacc_restated as nf
from oracle import reference_path as rp
from oracle.bilinear import grid_sample_2d_manual, corner_indices


def test_bilinear_matches_aten_value_and_first_derivative():
    g = torch.Generator().manual_seed(0)
    inp = torch.randn(3, 5, 8, 16, generator=g, requires_grad=True)
    grid = (torch.rand(3, 1, 200, 2, generator=g) * 2.4 - 1.2).requires_grad_(True)  # incl. out of bounds
    a = grid_sample_2d_manual(inp, grid)
    b = F.grid_sample(inp, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    assert (a - b).abs().max() < 1e-6
    cot = torch.randn(a.shape, generator=g)
    ga = torch.autograd.grad((a * cot).sum(), [inp, grid])
    gb = torch.autograd.grad((b * cot).sum(), [inp, grid])
    assert (ga[0] - gb[0]).abs().max() < 1e-5 and (ga[1] - gb[1]).abs().max() < 1e-4


def test_bilinear_second_derivative_fp64():
    g = torch.Generator().manual_seed(1)
    inp = torch.randn(2, 3, 4, 5, generator=g, dtype=torch.float64, requires_grad=True)
    # keep samples away from texel boundaries where the function is not differentiable
    base = torch.tensor([-0.9, -0.55, -0.1, 0.33, 0.71])
    grid = torch.stack(torch.meshgrid(base, base, indexing="ij"), -1).reshape(1, 1, 25, 2).repeat(2, 1, 1, 1)
    grid = (grid.double() + 0.013).requires_grad_(True)
    assert torch.autograd.gradcheck(grid_sample_2d_manual, (inp, grid), atol=1e-6)
    assert torch.autograd.gradgradcheck(grid_sample_2d_manual, (inp, grid), atol=1e-6)


def test_corner_indices_known_answers():
    # align_corners=False: g=-1 -> ix=-0.5 -> floor -1 ; g=+1 -> ix=W-0.5 -> floor W-1 ; g=0 -> (W-1)/2
    grid = torch.tensor([[[[-1.0, -1.0], [1.0, 1.0], [0.0, 0.0], [-1.0 + 1.0 / 8, 1.0 - 1.0 / 8]]]])
    ix, iy = corner_indices(grid, H=8, W=8)
    assert ix.flatten().tolist() == [-1, 7, 3, 0] and iy.flatten().tolist() == [-1, 7, 3, 7]


def test_rotation_folding_identity():
    """Sampling rotated planes == sampling the un-rotated planes at remapped coordinates (SURVEY probe table)."""
    g = torch.Generator().manual_seed(2)
    sc = torch.randn(1, 6, 4, 8, 8, generator=g)
    pts = torch.rand(1, 50, 3, generator=g) * 2 - 1
    geo = rp.interpolate_encodings(pts, sc, only_geo=True)
    x, y, z = pts[..., 0], pts[..., 1], pts[..., 2]

    def samp(plane, gx, gy):
        return F.grid_sample(plane, torch.stack([gx, gy], -1)[:, None], align_corners=False)[:, :, 0].permute(0, 2, 1)
    alt = samp(sc[:, 0], y, x) + samp(sc[:, 1], -x, -z) + samp(sc[:, 2], y, -z)
    assert (geo - alt).abs().max() < 1e-5


def test_weights_from_alpha_closed_forms():
    n_rays, S = 3, 5
    ray_idx = torch.arange(n_rays).repeat_interleave(S)
    w, T = nf.render_weight_from_alpha(torch.zeros(n_rays * S), ray_idx, n_rays)
    assert w.abs().max() == 0 and (T == 1).all()
    w, T = nf.render_weight_from_alpha(torch.ones(n_rays * S), ray_idx, n_rays)
    assert w.view(n_rays, S)[:, 0].eq(1).all() and w.view(n_rays, S)[:, 1:].eq(0).all()
    a = torch.full((n_rays * S,), 0.5)
    w, _ = nf.render_weight_from_alpha(a, ray_idx, n_rays)
    assert torch.allclose(w.view(n_rays, S)[0], torch.tensor([0.5, 0.25, 0.125, 0.0625, 0.03125]))
    acc = nf.accumulate_along_rays(w, None, ray_idx, n_rays)
    assert torch.allclose(acc[:, 0], torch.full((n_rays,), 1 - 0.5 ** S))
    # ragged packed rays take the general branch
    ridx = torch.tensor([0, 0, 0, 2, 2])
    w, T = nf.render_weight_from_alpha(torch.full((5,), 0.5), ridx, 3)
    assert torch.allclose(w, torch.tensor([0.5, 0.25, 0.125, 0.5, 0.25]))
    assert nf.accumulate_along_rays(w, None, ridx, 3)[:, 0].tolist() == [0.875, 0.0, 0.75]


def test_importance_sampling_known_answers():
    n_rays = 2
    cdf01 = torch.tensor([[0.0, 1.0]]).repeat(n_rays, 1)
    edges = nf.importance_sampling(cdf01, cdf01, 4)
    assert torch.equal(edges, torch.tensor([[0.0, 0.25, 0.5, 0.75, 1.0]]).repeat(n_rays, 1))
    # all mass in the second of three bins: every interior quantile lands inside that bin
    vals = torch.tensor([[0.0, 1.0, 2.0, 3.0]])
    cdfs = torch.tensor([[0.0, 0.0, 1.0, 1.0]])
    e = nf.importance_sampling(vals, cdfs, 4)
    assert torch.allclose(e, torch.tensor([[1.0, 1.25, 1.5, 1.75, 2.0]]))
    # uniform density reproduces the identity
    vals = torch.linspace(0, 1, 9)[None]
    e = nf.importance_sampling(vals, vals.clone(), 16)
    assert torch.allclose(e, torch.linspace(0, 1, 17)[None], atol=1e-7)
    # stratified: one jitter per ray, quantiles (j + b) / (n + 1) stay in [0, 1)
    e = nf.importance_sampling(cdf01[:1], cdf01[:1], 3, stratified=True, jitter=torch.tensor([0.5]))
    assert torch.allclose(e, torch.tensor([[0.125, 0.375, 0.625, 0.875]]))


def test_transmittance_closed_form():
    t0 = torch.tensor([[0.0, 1.0, 2.0]])
    t1 = torch.tensor([[1.0, 2.0, 3.0]])
    trans, alphas = nf.render_transmittance_from_density(t0, t1, torch.full((1, 3), 0.5))
    assert torch.allclose(trans, torch.exp(-torch.tensor([[0.0, 0.5, 1.0]])))
    assert torch.allclose(alphas, 1 - torch.exp(torch.tensor(-0.5)).expand(1, 3))


def test_sampler_shapes_sorted_and_dense_indices():
    g = torch.Generator().manual_seed(3)
    cfg = rp.PathConfig(num_samples_per_ray=8, num_samples_per_ray_importance=16)
    C = 4
    sc = torch.randn(1, 6, C, 8, 8, generator=g) * 0.5
    w = {"sdf": [torch.randn(64, C, generator=g) * 0.3, torch.randn(64, 64, generator=g) * 0.2,
                 torch.randn(1, 64, generator=g) * 0.2],
         "feature": [torch.randn(64, 3 * C, generator=g) * 0.3, torch.randn(64, 64, generator=g) * 0.2,
                     torch.randn(3, 64, generator=g) * 0.2]}
    o = torch.tensor([0.0, -2.0, 0.0]).expand(1, 2, 3, 3).contiguous()
    d = F.normalize(torch.tensor([0.0, 1.0, 0.0]) + 0.1 * torch.randn(1, 2, 3, 3, generator=g), dim=-1)
    t0, t1 = rp.sample_intervals(o, d, sc, w, cfg)
    assert t0.shape == (6, cfg.n_intervals) == t1.shape
    assert (t1 >= t0).all() and torch.equal(t0[:, 1:], t1[:, :-1])
    assert abs(t0[:, 0].min().item() - cfg.near_plane) < 1e-6 and abs(t1[:, -1].max().item() - cfg.far_plane) < 1e-5
