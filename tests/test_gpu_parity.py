"""Parity of the sm_100a kernels (through the plugin API and the C ABI) against the golden vectors produced by the
reference's own source and against the CPU oracle.  Needs a B200: run with ``pytest -m gpu``.

Tolerances: 1e-4 absolute on every forward output (north_star: 1e-4 on RGB / sigma); ray/sample indices and
plane repacking bit-exact; gradients 2e-3 relative to the tensor's max magnitude (they are sums of up to 1e5
atomically accumulated fp32 terms)."""
import pytest
import torch

import triplaneturbo_b200 as tt
from triplaneturbo_b200 import ops
from oracle import reference_path as rp
from tests.helpers import (load_golden, weights_from, max_abs, rel_err, proposal_cdf, assert_intervals_close,
                           build_plugins)

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=[2, 1, 0], ids=["tcgen05-ws", "tcgen05-r1", "simt"])
def kernel_family(request):
    """Every parity test runs on both kernel families: tcgen05 tensor-core kernels and the SIMT reference kernels."""
    ops.set_impl(request.param)
    yield request.param
    ops.set_impl(2)


TOL = 1e-4
GTOL = 2e-3
DEV = "cuda"


def test_library_is_the_cuda_build():
    from triplaneturbo_b200 import _cabi
    L = _cabi.load()
    assert L.tt_version() == 100 and L.tt_device_ok() == 1


def test_repack_bit_exact():
    g = torch.Generator().manual_seed(0)
    sc = torch.randn(2, 6, 8, 20, 20, generator=g)
    planes = ops.repack_planes(sc.to(DEV))
    assert torch.equal(planes.cpu(), rp.rotate_planes(sc).permute(0, 1, 3, 4, 2).contiguous())
    assert torch.equal(ops.repack_planes_bwd(planes).cpu(), sc)
    tri = torch.randn(1, 6, 16, 20, 20, generator=g)
    p2 = ops.repack_planes(tri.to(DEV), 8, 0, 8)
    assert torch.equal(p2.cpu(), rp.rotate_planes(rp.decode_split_channels(tri)).permute(0, 1, 3, 4, 2).contiguous())


@pytest.mark.parametrize("name", ["geometry_c8_r16", "geometry_c32_r16"])
def test_geometry_plugin_matches_reference(name):
    fx = load_golden(name, DEV)
    geom, _ = build_plugins(fx, DEV)
    sc = fx["space_cache"].clone().requires_grad_(True)
    out = geom(fx["points"], sc, output_normal=True)
    for k in ("sdf", "sdf_orig", "features", "normal", "shading_normal", "sdf_grad"):
        assert max_abs(out[k], fx["out_" + k]) < TOL, k
    loss = sum((out[k] * fx["cot_" + k]).sum() for k in ("sdf", "features", "normal", "sdf_grad"))
    params = geom.decoder_weights()
    grads = torch.autograd.grad(loss, [sc] + params)
    assert rel_err(grads[0], fx["grad_space_cache"]) < GTOL
    for i in range(3):
        assert rel_err(grads[1 + i], fx[f"grad_w_sdf_{i}"]) < GTOL, i
        assert rel_err(grads[4 + i], fx[f"grad_w_feature_{i}"]) < GTOL, i
    with torch.no_grad():
        sdf, deform = geom.forward_field(fx["points"], fx["space_cache"])
        assert max_abs(sdf, fx["field_sdf"]) < TOL and max_abs(deform, fx["field_deformation"]) < TOL
        assert max_abs(geom.forward_sdf(fx["points"], fx["space_cache"]), fx["forward_sdf"]) < TOL
        feats = geom.export(fx["points"][:1], fx["space_cache"][:1])["features"]
        assert max_abs(feats, fx["export_features"]) < TOL
        assert torch.equal(geom.decode(fx["triplane"]), fx["decoded"])


def test_forward_field_grid_matches_oracle_vertex_order():
    fx = load_golden("geometry_c8_r16", DEV)
    geom, _ = build_plugins(fx, DEV)
    res = 17
    with torch.no_grad():
        sdf, deform = geom.forward_field_grid(res, fx["space_cache"][:1])
    cpu = load_golden("geometry_c8_r16")
    s_ref, d_ref = rp.forward_field(rp.isosurface_grid_points(res)[None], cpu["space_cache"][:1], weights_from(cpu),
                                    rp.PathConfig())
    assert max_abs(sdf.cpu(), s_ref) < TOL and max_abs(deform.cpu(), d_ref) < TOL


def _renderer_for(fx):
    P, V, H, W, ns, nimp = [int(v) for v in fx["meta"][:6]]
    geom, rend = build_plugins(fx, DEV, ns, nimp, normal_direction=fx["normal_direction"],
                               rgb_grad_shrink=float(fx["rgb_grad_shrink"]),
                               trainable_variance=bool(int(fx.get("trainable_variance", 0))))
    rend.cos_anneal_ratio = float(fx.get("cos_anneal_ratio", 1.0))      # the attribute get_alpha reads (NEUS:91,100-103)
    return geom, rend, (P, V, H, W, ns, nimp)


def _kw(fx, P):
    explicit_bg = bool(fx["meta"][7])
    return dict(rays_o=fx["rays_o"], rays_d=fx["rays_d"], light_positions=torch.zeros(fx["rays_o"].shape[0], 3, device=DEV),
                bg_color=torch.ones(3, device=DEV) if explicit_bg else fx["bg"], text_embed=torch.zeros(P, 4, device=DEV),
                camera_distances=fx["camera_distances"], c2w=fx["c2w"])


@pytest.mark.parametrize("name", ["render_train_c8", "render_train_c32", "render_train_front", "render_train_shrink",
                                  "render_train_cos_anneal", "render_train_variance"])
def test_renderer_training_matches_reference(name):
    fx = load_golden(name, DEV)
    geom, rend, (P, V, H, W, ns, nimp) = _renderer_for(fx)
    rend.train()
    sc = fx["space_cache"].clone().requires_grad_(True)
    # (1) the sampler reproduces the reference estimator's intervals (raw sorted edges from tt_importance_sample)
    cpu = load_golden(name)
    pc = rp.PathConfig(num_samples_per_ray=ns, num_samples_per_ray_importance=nimp)
    tv, cdf = proposal_cdf(cpu, pc)
    w = geom.decoder_weights()
    edges = ops.importance_sample(ops.cached_planes(fx["space_cache"]),
                                  ops.cached_wpack(w[:3], w[3:], geom._deformation_weights(), fx["space_cache"].shape[2]),
                                  rend.path_scalars(), fx["rays_o"], fx["rays_d"], V * H * W, nimp, ns).cpu()
    assert_intervals_close(edges, torch.cat([cpu["t_starts"], cpu["t_ends"][:, -1:]], 1), tv.double(), cdf.double(),
                           tol=3e-5)
    # (2) marching the reference's own intervals reproduces every output of the reference renderer
    out = rend(space_cache=sc, t_starts=fx["t_starts"], t_ends=fx["t_ends"], **_kw(fx, P))
    keys = [k[4:] for k in fx if k.startswith("out_") and k != "out_comp_rgb_bg"]
    assert "weights" in keys and "sdf_grad" in keys and "comp_normal" in keys
    for k in keys:
        if k == "ray_indices":
            assert torch.equal(out[k], fx["out_" + k])          # bit-exact contract
        else:
            assert max_abs(out[k], fx["out_" + k]) < TOL, k
    # (3) gradients of the reference's loss (image cotangents + eikonal)
    cot_keys = [k[4:] for k in fx if k.startswith("cot_")]
    loss = sum((out[k] * fx["cot_" + k]).sum() for k in cot_keys)
    loss = loss + 0.1 * ((torch.linalg.norm(out["sdf_grad"], ord=2, dim=-1) - 1.0) ** 2).sum()
    var = [rend.variance._inv_std] if name == "render_train_variance" else []
    grads = torch.autograd.grad(loss, [sc] + geom.decoder_weights() + var)
    assert rel_err(grads[0], fx["grad_space_cache"]) < GTOL
    for i in range(3):
        assert rel_err(grads[1 + i], fx[f"grad_w_sdf_{i}"]) < GTOL, i
        assert rel_err(grads[4 + i], fx[f"grad_w_feature_{i}"]) < GTOL, i
    if var:         # trainable NeuS variance (the class default, REN:53): d loss / d p through inv_std = exp(10 p)
        assert rel_err(grads[7], fx["grad_inv_std_param"]) < GTOL


def test_renderer_stratified_matches_reference():
    fx = load_golden("render_train_stratified", DEV)
    geom, rend, (P, V, H, W, ns, nimp) = _renderer_for(fx)
    rend.train()
    rend.randomized = True
    with torch.no_grad():
        out = rend(space_cache=fx["space_cache"], jitters=(fx["jitter0"], fx["jitter1"]), **_kw(fx, P))
    # stratified fixture: the CDF of this scene has no flat segments where samples land -> direct comparison
    for k in ("t_points", "t_intervals"):
        assert max_abs(out[k], fx["out_" + k]) < 5e-5, k
    for k in ("comp_rgb", "opacity", "weights"):
        assert max_abs(out[k], fx["out_" + k]) < 5e-4, k        # intervals differ by ~1e-5 -> alpha by ~1e-4


def test_renderer_eval_matches_reference():
    fx = load_golden("render_eval_c8", DEV)
    geom, rend, (P, V, H, W, ns, nimp) = _renderer_for(fx)
    rend.eval()
    with torch.no_grad():
        out = rend(space_cache=fx["space_cache"], **_kw(fx, P))
    assert "weights" not in out
    for k in ("comp_rgb", "comp_rgb_fg", "opacity", "depth", "z_variance", "disparity", "comp_normal",
              "comp_normal_cam_vis", "comp_normal_cam_vis_white"):
        assert max_abs(out[k], fx["out_" + k]) < 5e-4, k        # own sampler: see test above


def test_patch_renderer_matches_reference():
    fx = load_golden("patch_c8", DEV)
    P, V, H, W, ns, nimp, PS, ds = [int(v) for v in fx["meta"]]
    geom, rend = build_plugins(fx, DEV, ns, nimp)
    patch = tt.find("patch-renderer")(dict(patch_size=PS, global_downsample=ds, base_renderer_type="generative-space-sdf-volume-renderer",
                                           base_renderer=rend.cfg), geometry=geom, material=rend.material,
                                      background=rend.background).to(DEV)
    patch.base_renderer.variance.load_state_dict(rend.variance.state_dict())
    patch.train()
    torch.manual_seed(26)   # the seed make_golden.py drew the patch position with (PATCH:66-67)
    with torch.no_grad():
        out = patch(fx["rays_o"], fx["rays_d"], torch.zeros(P * V, 3, device=DEV), torch.ones(3, device=DEV),
                    space_cache=fx["space_cache"], text_embed=torch.zeros(P, 4, device=DEV),
                    camera_distances=fx["camera_distances"], c2w=fx["c2w"])
    assert [out["patch_x"], out["patch_y"]] == [int(v) for v in fx["patch_xy"]]
    for k in ("comp_rgb", "opacity", "depth", "disparity", "comp_normal", "comp_normal_cam_vis"):
        assert max_abs(out[k], fx["out_" + k]) < 5e-4, k


@pytest.mark.parametrize("C_", [16, 40, 64])
def test_other_channel_counts_against_oracle(C_):
    """Channel counts without a golden fixture (config 2 uses C=40): seeded inputs, oracle on the CPU."""
    g = torch.Generator().manual_seed(40 + C_)
    R, P, V, H, W, ns, nimp = 24, 1, 2, 3, 4, 12, 24
    sc = torch.randn(P, 6, C_, R, R, generator=g) * 0.5
    w = {"sdf": [torch.randn(64, C_, generator=g) * 0.25, torch.randn(64, 64, generator=g) * 0.2,
                 torch.randn(1, 64, generator=g) * 0.2],
         "feature": [torch.randn(64, 3 * C_, generator=g) * 0.15, torch.randn(64, 64, generator=g) * 0.2,
                     torch.randn(3, 64, generator=g) * 0.2]}
    from tests.helpers import camera_rays
    rays_o, rays_d, c2w, dist = camera_rays(P * V, H, W, seed=7)
    pc = rp.PathConfig(num_samples_per_ray=ns, num_samples_per_ray_importance=nimp)
    sc_ref = sc.clone().requires_grad_(True)
    w_ref = {k: [t.clone().requires_grad_(True) for t in v] for k, v in w.items()}
    t0, t1 = rp.sample_intervals(rays_o, rays_d, sc.repeat_interleave(V, 0), w, pc)
    ref = rp.render_forward(rays_o, rays_d, sc_ref, w_ref, pc, torch.ones(3), dist, c2w, t_starts=t0, t_ends=t1)
    cot = {k: torch.randn(ref[k].shape, generator=g) for k in ("comp_rgb", "opacity", "depth", "comp_normal_cam_vis")}

    def loss_of(o, dev):
        l = sum((o[k] * cot[k].to(dev)).sum() for k in cot)
        return l + 0.1 * ((torch.linalg.norm(o["sdf_grad"], dim=-1) - 1.0) ** 2).sum()
    g_ref = torch.autograd.grad(loss_of(ref, "cpu"), [sc_ref] + w_ref["sdf"] + w_ref["feature"])

    fx = {"space_cache": sc.to(DEV)}
    fx.update({f"w_{n}_{i}": t.to(DEV) for n, ws in w.items() for i, t in enumerate(ws)})
    geom, rend = build_plugins(fx, DEV, ns, nimp)
    rend.train()
    sc_g = sc.to(DEV).requires_grad_(True)
    out = rend(rays_o.to(DEV), rays_d.to(DEV), None, torch.ones(3, device=DEV), space_cache=sc_g,
               text_embed=torch.zeros(P, 4, device=DEV), camera_distances=dist.to(DEV), c2w=c2w.to(DEV),
               t_starts=t0.to(DEV), t_ends=t1.to(DEV))
    for k in ("comp_rgb", "opacity", "depth", "z_variance", "disparity", "comp_normal", "comp_normal_cam_vis",
              "weights", "sdf", "features", "normal", "sdf_grad"):
        # (this synthetic decoder has |sdf_grad| up to ~20: scale the absolute tolerance with the magnitude)
        assert max_abs(out[k].cpu(), ref[k]) < TOL * max(1.0, float(ref[k].abs().max())), k
    g_gpu = torch.autograd.grad(loss_of(out, DEV), [sc_g] + geom.decoder_weights())
    names = ["space_cache", "sdf.0", "sdf.1", "sdf.2", "feature.0", "feature.1", "feature.2"]
    errs = {n: rel_err(a.cpu(), b) for n, a, b in zip(names, g_gpu, g_ref)}
    assert all(e < GTOL for e in errs.values()), errs


def test_full_size_properties():
    """BASELINE config 2 shapes (R=256, C=40, 96+192 samples) on a 64x64 crop of one view: size-independent
    properties — sorted intervals, Σ weights = opacity <= 1, ray-batch independence, linearity of the backward."""
    from tests.helpers import camera_rays, random_decoder
    g = torch.Generator().manual_seed(3)
    C_, R, ns, nimp = 40, 256, 96, 192
    sc = (torch.randn(1, 6, C_, R, R, generator=g) * 0.5).to(DEV)
    fx = {"space_cache": sc}
    fx.update({k: v.to(DEV) for k, v in random_decoder(C_, seed=1).items()})
    geom, rend = build_plugins(fx, DEV, ns, nimp)
    rend.train()
    rays_o, rays_d, c2w, dist = [t.to(DEV) for t in camera_rays(1, 64, 64, seed=2)]
    kw = dict(light_positions=None, bg_color=torch.ones(3, device=DEV), text_embed=torch.zeros(1, 4, device=DEV),
              camera_distances=dist, c2w=c2w)
    sc1 = sc.clone().requires_grad_(True)
    out = rend(rays_o, rays_d, space_cache=sc1, **kw)
    S = ns + nimp + 1
    assert out["weights"].shape == (64 * 64 * S, 1)
    assert bool((out["t_intervals"] >= 0).all())
    wsum = out["weights"].view(-1, S).sum(1)
    assert max_abs(wsum, out["opacity"].view(-1)) < 1e-4 and float(out["opacity"].detach().max()) <= 1.0 + 1e-5
    assert torch.equal(out["ray_indices"], torch.arange(64 * 64, device=DEV).repeat_interleave(S))
    assert float(out["opacity"].mean()) > 0.02, "scene should not be empty"
    # ray-batch independence: the top half rendered alone equals the top half of the full render
    with torch.no_grad():
        half = rend(rays_o[:, :32], rays_d[:, :32], space_cache=sc, **kw)
    for k in ("comp_rgb", "opacity", "depth", "comp_normal"):
        assert max_abs(half[k], out[k][:, :32]) < 1e-5, k
    # backward is linear in the cotangent: grad(a*L1 + b*L2) = a*grad(L1) + b*grad(L2)
    c1, c2 = torch.randn_like(out["comp_rgb"]), torch.randn_like(out["opacity"])
    g1, = torch.autograd.grad((out["comp_rgb"] * c1).sum(), sc1, retain_graph=True)
    g2, = torch.autograd.grad((out["opacity"] * c2).sum(), sc1, retain_graph=True)
    g12, = torch.autograd.grad((out["comp_rgb"] * c1).sum() * 0.5 + (out["opacity"] * c2).sum() * 2.0, sc1)
    assert rel_err(g12, 0.5 * g1 + 2.0 * g2) < 1e-3
    assert float(g1.abs().max()) > 0


def test_compositor_ops_match_nerfacc_semantics():
    from triplaneturbo_b200.nerfacc_compat import render_weight_from_alpha, accumulate_along_rays
    from oracle import nerfacc_restated as nf
    g = torch.Generator().manual_seed(9)
    n, S = 37, 21
    alphas = torch.rand(n * S, generator=g)
    alphas[5 * S + 3] = 1.0          # an opaque sample: everything behind it has zero transmittance
    alphas[7 * S:8 * S] = 0.0        # an empty ray
    vals = torch.randn(n * S, 3, generator=g)
    ridx = torch.arange(n).repeat_interleave(S)
    a_ref = alphas.clone().requires_grad_(True)
    w_ref, T_ref = nf.render_weight_from_alpha(a_ref, ridx, n)
    c_ref = nf.accumulate_along_rays(w_ref, vals, ridx, n)
    cot = torch.randn(n, 3, generator=g)
    ga_ref, = torch.autograd.grad((c_ref * cot).sum() + w_ref.sum(), a_ref)
    a = alphas.to(DEV).requires_grad_(True)
    w, T = render_weight_from_alpha(a, ray_indices=ridx.to(DEV), n_rays=n)
    c = accumulate_along_rays(w, vals.to(DEV), ridx.to(DEV), n)
    assert max_abs(w.cpu(), w_ref) < 1e-6 and max_abs(T.cpu(), T_ref) < 1e-6 and max_abs(c.cpu(), c_ref) < 1e-5
    ga, = torch.autograd.grad((c * cot.to(DEV)).sum() + w.sum(), a)
    assert max_abs(ga.cpu(), ga_ref) < 1e-4


def test_empty_and_error_paths():
    fx = load_golden("render_train_c8", DEV)
    geom, rend, (P, V, H, W, ns, nimp) = _renderer_for(fx)
    with pytest.raises(Exception):
        geom(fx["rays_o"].cpu().reshape(1, -1, 3), fx["space_cache"].cpu())      # CPU tensors: no fallback
    with pytest.raises(Exception):
        ops.repack_planes(torch.zeros(1, 5, 8, 4, 4, device=DEV))
    with pytest.raises(Exception):
        ops.pack_weights([torch.zeros(64, 12, device=DEV), torch.zeros(64, 64, device=DEV),
                          torch.zeros(1, 64, device=DEV)], None, None, 12)        # unsupported channel count
    # rays that miss the volume entirely: opacity ~ 0, colour = background
    rend.eval()
    o = torch.tensor([[[[3.0, 3.0, 3.0]]]], device=DEV)
    d = torch.nn.functional.normalize(torch.tensor([[[[1.0, 1.0, 1.0]]]], device=DEV), dim=-1)
    with torch.no_grad():
        out = rend(o, d, None, torch.ones(3, device=DEV), space_cache=fx["space_cache"][:1],
                   camera_distances=torch.tensor([5.0], device=DEV), c2w=torch.eye(4, device=DEV)[None])
    assert float(out["opacity"].max()) < 0.05 and max_abs(out["comp_rgb"], torch.ones_like(out["comp_rgb"])) < 0.05
