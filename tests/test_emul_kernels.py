"""Kernel-logic checks on the CPU: the CUDA kernel sources compiled for the host with tests/emul/cuda_emul.h
(one OS thread per CUDA thread) are run through the C ABI on the golden fixtures and compared with the oracle /
the reference's golden outputs.  This validates the kernels' arithmetic before any GPU time is spent; the
authoritative parity tests are the ``-m gpu`` ones in tests/test_gpu_parity.py, which run the real sm_100a build.
"""
import numpy as np
import pytest
import torch

from oracle import reference_path as rp
from tests.emul.emul_api import Emul
from tests.helpers import load_golden, weights_from, max_abs, rel_err, proposal_cdf, assert_intervals_close
from triplaneturbo_b200.image_ops import compose_images

TOL = 1e-4      # north_star: 1e-4 on RGB / sigma
GTOL = 2e-3     # gradients, relative to the tensor's max magnitude


@pytest.fixture(scope="module", params=[2, 1, 0], ids=["tcgen05-ws", "tcgen05-r1", "simt"])
def em(request):
    e = Emul()
    e.set_impl(request.param)     # both kernel families run through the same checks
    return e


def np_w(w):
    return {k: [t.numpy() for t in v] for k, v in w.items()}


def test_repack_is_exact_rotation_and_adjoint(em):
    g = torch.Generator().manual_seed(0)
    sc = torch.randn(2, 6, 8, 12, 12, generator=g)
    planes = em.repack(sc.numpy())
    want = rp.rotate_planes(sc).permute(0, 1, 3, 4, 2).contiguous().numpy()
    assert np.array_equal(planes, want)
    # channel split folded in (decode, few_step…:180-196)
    tri = torch.randn(1, 6, 16, 12, 12, generator=g)
    planes2 = em.repack(tri.numpy(), C_=8, off_geo=0, off_tex=8)
    want2 = rp.rotate_planes(rp.decode_split_channels(tri)).permute(0, 1, 3, 4, 2).contiguous().numpy()
    assert np.array_equal(planes2, want2)
    # adjoint: <repack(x), y> == <x, repack_bwd(y)>  (pure permutation -> exact inverse)
    back = em.repack_bwd(planes)
    assert np.array_equal(back, sc.numpy())
    # adjoint with the channel split folded in: the used halves come back, the unused halves stay zero
    back2 = em.repack_bwd_split(planes2, 16, 0, 8)
    want_back = tri.numpy().copy()
    want_back[:, 0:3, 8:] = 0.0
    want_back[:, 3:6, :8] = 0.0
    assert np.array_equal(back2, want_back)


@pytest.mark.parametrize("name", ["geometry_c8_r16", "geometry_c32_r16"])
def test_geometry_forward(em, name):
    fx = load_golden(name)
    w = weights_from(fx)
    P, _, C_, R, _ = fx["space_cache"].shape
    cfg = em.config(C_, R, P)
    planes = em.repack(fx["space_cache"].numpy())
    wp = em.pack_weights(np_w(w), C_)
    out = em.geometry_fwd(planes, wp, cfg, fx["points"].numpy(), deform=True)
    for k in ("sdf", "sdf_orig", "features", "normal", "sdf_grad"):
        assert max_abs(torch.from_numpy(out[k]).reshape(fx["out_" + k].shape), fx["out_" + k]) < TOL, k
    assert max_abs(torch.from_numpy(out["deformation"]).reshape(fx["field_deformation"].shape),
                   fx["field_deformation"]) < TOL
    assert max_abs(torch.from_numpy(out["sdf"]).reshape(fx["field_sdf"].shape), fx["field_sdf"]) < TOL


@pytest.mark.parametrize("name", ["geometry_c8_r16", "geometry_c32_r16"])
def test_geometry_backward(em, name):
    fx = load_golden(name)
    w = weights_from(fx)
    P, _, C_, R, _ = fx["space_cache"].shape
    cfg = em.config(C_, R, P)
    planes = em.repack(fx["space_cache"].numpy())
    wp = em.pack_weights(np_w(w), C_)
    gplanes, gw = em.geometry_bwd(planes, wp, cfg, fx["points"].numpy(), g_sdf=fx["cot_sdf"].numpy().ravel(),
                                  g_features=fx["cot_features"].numpy(), g_normal=fx["cot_normal"].numpy(),
                                  g_sdf_grad=fx["cot_sdf_grad"].numpy())
    gsc = torch.from_numpy(em.repack_bwd(gplanes))
    assert rel_err(gsc, fx["grad_space_cache"]) < GTOL
    names = [f"grad_w_sdf_{i}" for i in range(3)] + [f"grad_w_feature_{i}" for i in range(3)]
    for n, g in zip(names, gw):
        assert rel_err(torch.from_numpy(g), fx[n]) < GTOL, n


@pytest.mark.parametrize("name", ["geometry_c8_r16", "geometry_c32_r16"])
def test_field_backward_with_deformation(em, name):
    """forward_field (sdf + deformation) is differentiable w.r.t. planes, the SDF decoder and the deformation decoder
    (few_step…diffusion.py:375-394; the mesh renderer trains through it): tt_field_bwd vs autograd of the oracle."""
    fx = load_golden(name)
    w = weights_from(fx)
    P, _, C_, R, _ = fx["space_cache"].shape
    cfg = em.config(C_, R, P)
    planes = em.repack(fx["space_cache"].numpy())
    wp = em.pack_weights(np_w(w), C_)
    g = torch.Generator().manual_seed(11)
    M = fx["points"].shape[1]
    cot_s, cot_d = torch.randn(P, M, 1, generator=g), torch.randn(P, M, 3, generator=g)
    sc = fx["space_cache"].clone().requires_grad_(True)
    wr = {k: [t.clone().requires_grad_(True) for t in v] for k, v in w.items()}
    sdf, deform = rp.forward_field(fx["points"], sc, wr, rp.PathConfig())
    loss = (sdf * cot_s).sum() + (deform * cot_d).sum()
    want = torch.autograd.grad(loss, [sc] + wr["sdf"] + wr["deformation"])
    gplanes, gws, gwd = em.field_bwd(planes, wp, cfg, fx["points"].numpy(), g_sdf=cot_s.numpy().ravel(),
                                     g_def=cot_d.numpy().reshape(-1, 3))
    assert rel_err(torch.from_numpy(em.repack_bwd(gplanes)), want[0]) < GTOL
    for i in range(3):
        assert rel_err(torch.from_numpy(gws[i]), want[1 + i]) < GTOL, f"sdf {i}"
        assert rel_err(torch.from_numpy(gwd[i]), want[4 + i]) < GTOL, f"deformation {i}"


@pytest.mark.parametrize("res", [9, 16])        # 16: the regular-grid z-line gather of the warp-specialised kernel
def test_isosurface_grid_points(em, res):
    fx = load_golden("geometry_c8_r16")
    w = weights_from(fx)
    cfg = em.config(8, 16, 1)
    planes = em.repack(fx["space_cache"][:1].numpy())
    wp = em.pack_weights(np_w(w), 8)
    em.set_option("grid_lines", 1 if res == 16 else 0)
    try:
        out = em.geometry_fwd(planes, wp, cfg, None, grid_res=res, normal=False, features=False, deform=True)
    finally:
        em.set_option("grid_lines", 0)
    pts = rp.isosurface_grid_points(res)[None]
    sdf, deform = rp.forward_field(pts, fx["space_cache"][:1], w, rp.PathConfig())
    assert max_abs(torch.from_numpy(out["sdf"]), sdf.reshape(-1)) < TOL
    assert max_abs(torch.from_numpy(out["deformation"]), deform.reshape(-1, 3)) < TOL


def _cfg_for(em, fx):
    P, V, H, W, ns, nimp = [int(v) for v in fx["meta"][:6]]
    C_, R = fx["space_cache"].shape[2], fx["space_cache"].shape[3]
    pc = rp.PathConfig(num_samples_per_ray=ns, num_samples_per_ray_importance=nimp,
                       normal_direction=fx["normal_direction"], rgb_grad_shrink=float(fx["rgb_grad_shrink"]),
                       cos_anneal_ratio=float(fx.get("cos_anneal_ratio", 1.0)))
    cfg = em.config(C_, R, P, rays_per_cache=V * H * W, step=pc.render_step_size, car=pc.cos_anneal_ratio)
    return cfg, pc, (P, V, H, W, ns, nimp)


@pytest.mark.parametrize("name", ["render_train_c8", "render_train_c32", "render_train_front"])
def test_importance_sampler(em, name):
    fx = load_golden(name)
    cfg, pc, (P, V, H, W, ns, nimp) = _cfg_for(em, fx)
    planes = em.repack(fx["space_cache"].numpy())
    wp = em.pack_weights(np_w(weights_from(fx)), cfg.C)
    t = torch.from_numpy(em.importance_sample(planes, wp, cfg, fx["rays_o"].numpy(), fx["rays_d"].numpy(), nimp, ns))
    tv, cdf = proposal_cdf(fx, pc)
    assert_intervals_close(t, torch.cat([fx["t_starts"], fx["t_ends"][:, -1:]], 1), tv.double(), cdf.double())


def test_importance_sampler_stratified(em):
    fx = load_golden("render_train_stratified")
    cfg, pc, (P, V, H, W, ns, nimp) = _cfg_for(em, fx)
    planes = em.repack(fx["space_cache"].numpy())
    wp = em.pack_weights(np_w(weights_from(fx)), cfg.C)
    t = torch.from_numpy(em.importance_sample(planes, wp, cfg, fx["rays_o"].numpy(), fx["rays_d"].numpy(), nimp, ns,
                                              fx["jitter0"].numpy(), fx["jitter1"].numpy()))
    tv, cdf = proposal_cdf(fx, pc, fx["jitter0"])
    assert_intervals_close(t, torch.cat([fx["t_starts"], fx["t_ends"][:, -1:]], 1), tv.double(), cdf.double())


@pytest.mark.parametrize("name", ["render_train_c8", "render_train_c32", "render_train_front", "render_train_shrink",
                                  "render_train_cos_anneal", "render_train_variance"])
def test_render_forward_backward(em, name):
    fx = load_golden(name)
    cfg, pc, (P, V, H, W, ns, nimp) = _cfg_for(em, fx)
    B = P * V
    planes = em.repack(fx["space_cache"].numpy())
    wp = em.pack_weights(np_w(weights_from(fx)), cfg.C)
    o, d = fx["rays_o"].numpy(), fx["rays_d"].numpy()
    fwd = em.render_fwd(planes, wp, cfg, o, d, fx["t_starts"].numpy(), fx["t_ends"].numpy())
    for k in ("sdf", "sdf_orig", "features", "normal", "sdf_grad", "weights"):
        assert max_abs(torch.from_numpy(fwd[k]).reshape(fx["out_" + k].shape), fx["out_" + k]) < TOL, k
    acc = torch.from_numpy(fwd["acc"]).requires_grad_(True)
    explicit_bg = bool(fx["meta"][7])
    bg = torch.ones(3) if explicit_bg else fx["bg"]
    img = compose_images(acc, bg, fx["camera_distances"], fx["c2w"], B, H, W, pc.normal_direction, V)
    for k in [k[4:] for k in fx if k.startswith("out_comp") or k in ("out_opacity", "out_depth", "out_z_variance",
                                                                      "out_disparity")]:
        if k == "comp_rgb_bg":
            continue
        assert max_abs(img[k], fx["out_" + k]) < TOL, k
    # backward: image-space cotangents through compose_images (torch), then the kernels
    cot_keys = [k[4:] for k in fx if k.startswith("cot_")]
    loss = sum((img[k] * fx["cot_" + k]).sum() for k in cot_keys)
    sg = torch.from_numpy(fwd["sdf_grad"])
    nrm = torch.linalg.norm(sg, dim=-1, keepdim=True)
    assert max_abs(img["eikonal_sum"].reshape(-1), ((nrm - 1.0) ** 2).reshape(acc.shape[0], -1).sum(1)) < 1e-3
    if name == "render_train_c32":     # eikonal term of the golden loss through the fused per-ray accumulator ...
        loss = loss + 0.1 * img["eikonal_sum"].sum()
        g_sdf_grad = None
    else:                              # ... or through the per-sample sdf_grad output, like the reference's system
        g_sdf_grad = (0.1 * 2.0 * (nrm - 1.0) * sg / nrm).numpy()
    g_acc, = torch.autograd.grad(loss, acc)
    if H % 4 == 0 and W % 4 == 0:           # image-shape hint: patch-ordered sample lists (k_patch_lists) + tile-merged
        cfg.image_h, cfg.image_w = H, W     # scatter; the other fixtures run the plain / run-length scatter on ray order
        em.set_option("patch_lists", 1); em.set_option("scatter", 2)
    try:
        gplanes, gw, gis = em.render_bwd(planes, wp, cfg, o, d, fx["t_starts"].numpy(), fx["t_ends"].numpy(), fwd,
                                         g_acc.numpy(), g_sdf_grad=g_sdf_grad, rgb_scale=pc.rgb_grad_shrink)
    finally:
        em.set_option("patch_lists", 1); em.set_option("scatter", -1)
    assert rel_err(torch.from_numpy(em.repack_bwd(gplanes)), fx["grad_space_cache"]) < GTOL
    names = [f"grad_w_sdf_{i}" for i in range(3)] + [f"grad_w_feature_{i}" for i in range(3)]
    for n, g in zip(names, gw):
        assert rel_err(torch.from_numpy(g), fx[n]) < GTOL, n
    if name == "render_train_variance":     # d loss / d p with inv_std = exp(10 p): chain rule on the kernel's d loss / d inv_std
        want = float(fx["grad_inv_std_param"])
        got = gis * 10.0 * cfg.inv_std
        assert abs(got - want) < GTOL * max(1.0, abs(want)), (got, want)


def test_inv_std_gradient(em):
    """d loss / d inv_std of the compositing chain against autograd through the oracle."""
    fx = load_golden("render_train_c8")
    cfg, pc, (P, V, H, W, ns, nimp) = _cfg_for(em, fx)
    planes = em.repack(fx["space_cache"].numpy())
    w = weights_from(fx)
    wp = em.pack_weights(np_w(w), cfg.C)
    o, d = fx["rays_o"].numpy(), fx["rays_d"].numpy()
    t0, t1 = fx["t_starts"], fx["t_ends"]
    fwd = em.render_fwd(planes, wp, cfg, o, d, t0.numpy(), t1.numpy())
    g_acc = torch.randn(fwd["acc"].shape, generator=torch.Generator().manual_seed(5))
    g_acc[:, 9] = 0.0
    _, _, gis = em.render_bwd(planes, wp, cfg, o, d, t0.numpy(), t1.numpy(), fwd, g_acc.numpy())
    # oracle: alpha(inv_std) -> weights -> accumulators, everything else held fixed
    inv_std = torch.tensor(cfg.inv_std, requires_grad=True)
    sdf = torch.from_numpy(fwd["sdf"])[:, None]
    normal = torch.from_numpy(fwd["normal"])
    S = t0.shape[1]
    dirs = fx["rays_d"].reshape(-1, 3).repeat_interleave(S, 0)
    tm = ((t0 + t1) / 2).reshape(-1, 1)
    alpha = rp.get_alpha(sdf, normal, dirs, (t1 - t0).reshape(-1, 1), inv_std)[:, 0].reshape(-1, S)
    T = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1 - alpha[:, :-1]], 1), 1)
    wgt = (T * alpha).reshape(-1, 1)
    n_rays = t0.shape[0]
    rgb = rp.sigmoid_mipnerf(torch.from_numpy(fwd["features"]))
    depth = (wgt * tm).reshape(n_rays, S, 1).sum(1)
    acc = torch.cat([wgt.reshape(n_rays, S, 1).sum(1), depth, (wgt * rgb).reshape(n_rays, S, 3).sum(1),
                     (wgt * (tm - depth.repeat_interleave(S, 0)) ** 2).reshape(n_rays, S, 1).sum(1),
                     (wgt * normal).reshape(n_rays, S, 3).sum(1)], 1)
    want, = torch.autograd.grad((acc * g_acc[:, :9]).sum(), inv_std)
    assert abs(gis - want.item()) < 2e-3 * max(1.0, abs(want.item()))


@pytest.mark.parametrize("D", [0, 3, 5])
def test_standalone_compositor(em, D):
    """render_weight_from_alpha + accumulate_along_rays on dense rays (warp-per-ray kernels for the tensor-core families,
    thread-per-ray for the SIMT family) against the restated nerfacc semantics, forward and backward; S not a multiple of
    32, fully opaque and fully empty rays included."""
    from oracle import nerfacc_restated as nf
    g = torch.Generator().manual_seed(D)
    n, S = 7, 45
    alphas = torch.rand(n, S, generator=g) * 0.3
    alphas[1] = 0.0
    alphas[2, 5] = 1.0
    values = torch.randn(n, S, D, generator=g) if D else None
    a = alphas.clone().requires_grad_(True)
    v = values.clone().requires_grad_(True) if D else None
    ray_idx = torch.arange(n).repeat_interleave(S)
    w_ref, T_ref = nf.render_weight_from_alpha(a.reshape(-1), ray_idx, n)
    out_ref = nf.accumulate_along_rays(w_ref, v.reshape(-1, D) if D else None, ray_idx, n)
    w, T, out = em.composite_fwd(alphas.numpy(), None if values is None else values.numpy())
    assert max_abs(torch.from_numpy(w).reshape(-1), w_ref) < 1e-6 and max_abs(torch.from_numpy(T).reshape(-1), T_ref) < 1e-6
    assert max_abs(torch.from_numpy(out), out_ref) < 1e-5
    g_out, g_w = torch.randn(out_ref.shape, generator=g), torch.randn(n, S, generator=g)
    loss = (out_ref * g_out).sum() + (w_ref.reshape(n, S) * g_w).sum()
    want = torch.autograd.grad(loss, [a] + ([v] if D else []))
    ga, gv = em.composite_bwd(alphas.numpy(), None if values is None else values.numpy(), T, g_out.numpy(), g_w.numpy())
    assert max_abs(torch.from_numpy(ga), want[0]) < 1e-5
    if D:
        assert max_abs(torch.from_numpy(gv), want[1]) < 1e-5
