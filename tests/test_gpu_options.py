"""Switchable paths (include/triplane_b200.h tt_set_option) against each other on the GPU: they compute the same sums in
another order, so gradients agree to the accuracy of fp32 reductions and forward outputs to 1e-6.
  * "patch_lists" + "scatter" = 2 (the default when the image shape is known): patch-ordered sample lists (k_patch_lists)
    + tile-merged hidden-gradient scatter
  * "scatter" = 0 / 1: plain / run-length merged scatter, "patch_lists" = 0: ray-ordered lists
  * "grid_lines": z-line gather of the regular isosurface grid (ws_grid_segment)"""
import pytest
import torch

from tests.helpers import build_plugins, rel_err
from triplaneturbo_b200 import ops
from triplaneturbo_b200.synthetic import camera_rays, random_decoder, random_triplanes

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _defaults():
    yield
    for k, v in (("scatter", -1), ("patch_lists", 1), ("grid_lines", 0)):       # the library defaults
        ops.set_option(k, v)
    ops.set_impl(2)


def _train_grads(C, R, H, W, ns, nimp, seed=0):
    sc = random_triplanes(2, C, R, seed=seed).to(DEV)
    fx = {"space_cache": sc, **{k: v.to(DEV) for k, v in random_decoder(C, seed=1).items()}}
    geom, rend = build_plugins(fx, DEV, ns, nimp, rgb_grad_shrink=0.3)
    rend.train()
    w = geom.decoder_weights()
    for p_ in w:
        p_.requires_grad_(True)
    rays_o, rays_d, c2w, dist = [t.to(DEV) for t in camera_rays(4, H, W, seed=5, views_per_prompt=2)]
    g = torch.Generator().manual_seed(3)
    cot = torch.randn(4, H, W, 3, generator=g).to(DEV)

    def run():
        sc_g = sc.clone().requires_grad_(True)
        out = rend(rays_o, rays_d, None, torch.ones(3, device=DEV), space_cache=sc_g, text_embed=torch.zeros(2, 4, device=DEV),
                   camera_distances=dist, c2w=c2w)
        loss = (out["comp_rgb"] * cot).sum() + out["opacity"].sum() + 0.1 * ((out["sdf_grad"].norm(dim=-1) - 1) ** 2).mean()
        return [x.clone() for x in torch.autograd.grad(loss, [sc_g] + w)], out["comp_rgb"].detach().clone()
    return run


@pytest.mark.parametrize("impl", [2, 1], ids=["tcgen05-ws", "tcgen05-r1"])
@pytest.mark.parametrize("C,R,S,HW", [(32, 64, (32, 64), 24), (40, 32, (96, 192), 24), (8, 16, (8, 16), 8)],
                         ids=["c32", "c40-long-rays", "tiny"])          # tiny: the case tools/evidence.sh runs under the sanitizer
def test_scatter_variants_and_patch_lists_agree(impl, C, R, S, HW):
    ops.set_impl(impl)
    run = _train_grads(C, R, HW, HW, *S)
    ref, img = run()
    for scatter, patch in ((2, 1), (2, 0), (0, 1), (0, 0), (1, 0)):
        ops.set_option("scatter", scatter)
        ops.set_option("patch_lists", patch)
        got, img2 = run()
        assert torch.equal(img, img2)
        for a, b in zip(got, ref):
            assert rel_err(a, b) < 1e-5, (scatter, patch)


def test_patch_lists_need_image_dimensions_divisible_by_four():
    """22 x 22 views: the hint is ignored (ray-ordered lists), results unchanged."""
    run = _train_grads(32, 32, 22, 22, 16, 32)
    ref, _ = run()
    ops.set_option("scatter", 2)
    ops.set_option("patch_lists", 1)
    got, _ = run()
    for a, b in zip(got, ref):
        assert rel_err(a, b) < 1e-5


@pytest.mark.parametrize("C,R,res", [(32, 64, 64), (8, 16, 16), (32, 32, 128)])
def test_grid_line_gather_matches_the_tap_gather(C, R, res):
    sc = random_triplanes(2, C, R, seed=4).to(DEV)
    fx = {"space_cache": sc, **{k: v.to(DEV) for k, v in random_decoder(C, seed=1).items()}}
    geom, _ = build_plugins(fx, DEV, 32, 64)
    with torch.no_grad():
        s0, d0 = geom.forward_field_grid(res, sc)
        ops.set_option("grid_lines", 1)
        s1, d1 = geom.forward_field_grid(res, sc)
        s2, d2 = geom.forward_field_grid(res, sc)
    assert torch.equal(s1, s2) and torch.equal(d1, d2)                       # deterministic
    assert float((s0 - s1).abs().max()) < 2e-6 * max(1.0, float(s0.abs().max()))
    assert float((d0 - d1).abs().max()) < 2e-6 * max(1.0, float(d0.abs().max()))
    assert not torch.equal(s0, s1) or res < 16                              # (it IS another summation order)
