"""GPU parity at the BASELINE configurations and round-2 boundary checks (run on the B200 box with ``-m gpu``).

  * config 1 in full (16 384 rays, R=64, C=32, S=193) and 64x64-ray crops of configs 2 and 3 against the CPU oracle:
    every output of the renderer at 1e-4, gradients included (VERDICT r1 "parity gaps" 1)
  * determinism: repeated field queries and repeated renders are bit-identical in the forward
  * differentiable forward_field (sdf + deformation), interpolate_encodings, closure-driven estimator
  * backward precision: elementwise gradient error of the CUDA path and of a plain fp32 torch run of the oracle on the
    same GPU, both against the fp64 oracle
"""
import json
import os

import pytest
import torch

from oracle import reference_path as rp
from tests.helpers import (assert_intervals_close, build_plugins, load_golden, max_abs, proposal_cdf, rel_err,
                           weights_from)
from triplaneturbo_b200 import ops
from triplaneturbo_b200.synthetic import camera_rays, random_decoder, random_triplanes

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-4       # north_star: 1e-4 on RGB / sigma
GTOL = 2e-3      # gradients, relative to the tensor's max magnitude (see test_backward_precision_against_fp64 for why)


@pytest.fixture(autouse=True, params=[2, 1], ids=["tcgen05-ws", "tcgen05-r1"])
def kernel_family(request):
    ops.set_impl(request.param)
    yield request.param
    ops.set_impl(2)


def _scene(P, V, H, W, R, C, ns, nimp, crop=None, seed=0):
    sc = random_triplanes(P, C, R, seed=seed)
    wts = random_decoder(C, seed=1)
    rays_o, rays_d, c2w, dist = camera_rays(P * V, H, W, seed=2, views_per_prompt=V)
    if crop is not None:            # centre crop of every view: rays that actually hit the volume
        y0, x0 = (H - crop) // 2, (W - crop) // 2
        rays_o = rays_o[:, y0:y0 + crop, x0:x0 + crop].contiguous()
        rays_d = rays_d[:, y0:y0 + crop, x0:x0 + crop].contiguous()
    return sc, wts, rays_o, rays_d, c2w, dist


def _compare_with_oracle(P, V, R, C, ns, nimp, sc, wts, rays_o, rays_d, c2w, dist):
    """Sampler edges, every renderer output and the gradients of a training-style loss: CUDA path vs CPU oracle."""
    B, H, W = rays_o.shape[:3]
    S = ns + nimp + 1
    fx = {"space_cache": sc.to(DEV), **{k: v.to(DEV) for k, v in wts.items()}}
    geom, rend = build_plugins(fx, DEV, ns, nimp, rgb_grad_shrink=0.3)
    rend.train()
    w = geom.decoder_weights()
    for p_ in w:
        p_.requires_grad_(True)
    pc = rp.PathConfig(num_samples_per_ray=ns, num_samples_per_ray_importance=nimp, rgb_grad_shrink=0.3)
    wc = {n: [wts[f"w_{n}_{i}"].clone().requires_grad_(True) for i in range(3)] for n in ("sdf", "feature")}
    # (1) sampler: edges of tt_importance_sample against the oracle's estimator (flat-CDF allowance, tests/helpers.py)
    edges = ops.importance_sample(ops.cached_planes(fx["space_cache"]),
                                  ops.cached_wpack(w[:3], w[3:], geom._deformation_weights(), C),
                                  rend.path_scalars(), rays_o.to(DEV), rays_d.to(DEV), V * H * W, nimp, ns)
    cpu_fx = {"space_cache": sc, "rays_o": rays_o, "rays_d": rays_d, **wts}
    tv, cdf = proposal_cdf(cpu_fx, pc)
    with torch.no_grad():
        t0_ref, t1_ref = rp.sample_intervals(rays_o, rays_d, sc.repeat_interleave(V, 0), wc, pc)
    assert_intervals_close(edges.cpu(), torch.cat([t0_ref, t1_ref[:, -1:]], 1), tv.double(), cdf.double(), tol=3e-5)
    # (2) both sides march the SAME intervals (the CUDA sampler's)
    t0, t1 = edges[:, :-1], edges[:, 1:]
    sc_g = sc.to(DEV).requires_grad_(True)
    out = rend(rays_o.to(DEV), rays_d.to(DEV), None, torch.ones(3, device=DEV), space_cache=sc_g,
               text_embed=torch.zeros(P, 4, device=DEV), camera_distances=dist.to(DEV), c2w=c2w.to(DEV),
               t_starts=t0, t_ends=t1)
    sc_c = sc.clone().requires_grad_(True)
    ref = rp.render_forward(rays_o, rays_d, sc_c, wc, pc, torch.ones(3), dist, c2w, t_starts=t0.cpu(), t_ends=t1.cpu())
    assert torch.equal(out["ray_indices"].cpu(), ref["ray_indices"])                  # bit-exact contract
    # The analytic normal is the derivative of a ReLU network: it is DISCONTINUOUS where a hidden pre-activation crosses zero.
    # At a handful of the ~10^6 samples a pre-activation of the reference lies within 2e-5 of zero, and an implementation
    # whose sums round differently (3xTF32 MMAs vs fp32 FMA chains) may take the other branch there: sdf agrees to 1e-6,
    # the normal of that one sample does not.  Such samples are identified FROM THE ORACLE (|z| < 2e-5), every normal
    # mismatch must be one of them, they must be rare, and everything else has to match to 1e-4.
    n_rays = B * H * W
    with torch.no_grad():
        enc = rp.interpolate_encodings(rp.rescale_points(ref["points"].detach().reshape(1, -1, 3), 1.0), sc, only_geo=True)
        enc = enc.reshape(-1, C)
        z1 = enc @ wts["w_sdf_0"].T
        z2 = z1.relu() @ wts["w_sdf_1"].T
        ambiguous = ((z1.abs() < 2e-5).any(-1) | (z2.abs() < 2e-5).any(-1)) & (enc.abs().sum(-1) > 0)
    dn = (out["normal"].detach().cpu() - ref["normal"].detach()).abs().max(-1).values
    flipped = dn > TOL
    assert not bool((flipped & ~ambiguous).any()), "a normal differs where the reference's ReLU pattern is stable"
    assert int(flipped.sum()) <= max(1e-4 * flipped.numel(), 2), int(flipped.sum())
    sample_ok = ~flipped
    ray_ok = ~flipped.view(n_rays, S).any(1)
    bad = {}
    # images derived from normalize(sum_i w_i n_i): on these synthetic NOISY planes the per-sample normals of a ray point
    # everywhere, |sum| << opacity, and the normalisation amplifies the 1e-5 per-sample differences (condition number
    # opacity / |sum| ~ 10): 3e-4 there, 1e-4 on RGB / opacity / depth as north_star states
    for k, tol in (("comp_rgb", TOL), ("comp_rgb_fg", TOL), ("opacity", TOL), ("depth", TOL), ("z_variance", TOL),
                   ("disparity", TOL), ("comp_normal_cam_vis", 3 * TOL), ("comp_normal_cam_vis_white", 3 * TOL)):
        a, b = out[k].detach().cpu().reshape(n_rays, -1), ref[k].detach().reshape(n_rays, -1)
        err = max_abs(a[ray_ok], b[ray_ok])
        if not err < tol:
            bad[k] = err
    for k in ("sdf", "sdf_orig", "features", "normal", "t_points", "t_intervals", "points", "t_dirs"):
        a, b = out[k].detach().cpu(), ref[k].detach()
        err = max_abs(a[sample_ok], b[sample_ok])
        if not err < TOL:
            bad[k] = err
    wa, wb = out["weights"].detach().cpu().view(n_rays, S), ref["weights"].detach().view(n_rays, S)
    if not max_abs(wa[ray_ok], wb[ray_ok]) < TOL:
        bad["weights"] = max_abs(wa[ray_ok], wb[ray_ok])
    assert not bad, bad
    # comp_normal = normalize(sum_i w_i n_i) is ill-conditioned where the ray's opacity is ~0 (a 1e-7 change of the sum
    # rotates the unit vector): compare it the way every consumer uses it, weighted by the opacity
    # (comp_normal_cam_vis above = (n_cam + 1) / 2 * opacity + ..., REN:475-511)
    dcn = ((out["comp_normal"].detach().cpu() - ref["comp_normal"].detach()).abs() * ref["opacity"].detach()).reshape(n_rays, -1)
    assert float(dcn[ray_ok].max()) < 3 * TOL, "comp_normal"
    # sdf_grad of the synthetic (noisy) planes reaches |g| ~ 50: 1e-4 relative to the largest component
    ga, gb = out["sdf_grad"].detach().cpu(), ref["sdf_grad"].detach()
    assert max_abs(ga[sample_ok], gb[sample_ok]) < TOL * max(1.0, float(gb.abs().max()))
    # (3) gradients of a training-style loss: image cotangents + eikonal on sdf_grad + sparsity (…generator.py:618-714)
    g = torch.Generator().manual_seed(3)
    cots = {k: torch.randn(B, H, W, d, generator=g) for k, d in (("comp_rgb", 3), ("comp_normal_cam_vis", 3), ("disparity", 1))}

    def loss_of(o, dev):
        l_ = sum((o[k] * cots[k].to(dev)).sum() for k in cots)
        l_ = l_ + 0.1 * ((torch.linalg.norm(o["sdf_grad"], ord=2, dim=-1) - 1.0) ** 2).mean() * 100.0
        return l_ + 0.5 * torch.sqrt(o["opacity"] ** 2 + 0.01).mean()
    g_gpu = torch.autograd.grad(loss_of(out, DEV), [sc_g] + w)
    g_cpu = torch.autograd.grad(loss_of(ref, "cpu"), [sc_c] + wc["sdf"] + wc["feature"])
    gerr = {i: rel_err(a.cpu(), b) for i, (a, b) in enumerate(zip(g_gpu, g_cpu))}
    assert all(v < GTOL for v in gerr.values()), gerr


def test_config1_full_against_oracle():
    """BASELINE configs[0] in full: single 64^2 x 32ch triplane, 1 camera 128^2, 64 + 128 samples per ray."""
    P, V, H, W, R, C, ns, nimp = 1, 1, 128, 128, 64, 32, 64, 128
    sc, wts, rays_o, rays_d, c2w, dist = _scene(P, V, H, W, R, C, ns, nimp)
    _compare_with_oracle(P, V, R, C, ns, nimp, sc, wts, rays_o, rays_d, c2w, dist)


@pytest.mark.parametrize("name,R,C,ns,nimp,H", [("config2", 256, 40, 96, 192, 256), ("config3", 256, 32, 64, 128, 512)])
def test_crop_of_bench_configs_against_oracle(name, R, C, ns, nimp, H):
    """64 x 64 centre crop of one view of BASELINE configs[1] / configs[2] (full plane and sample sizes)."""
    sc, wts, rays_o, rays_d, c2w, dist = _scene(1, 1, H, H, R, C, ns, nimp, crop=64)
    _compare_with_oracle(1, 1, R, C, ns, nimp, sc, wts, rays_o, rays_d, c2w, dist)


def test_forward_is_deterministic():
    """200 repeated field queries and 50 repeated training renders: bit-identical forward outputs (a latent race in a
    tcgen05 kernel would show up here); gradients (floating-point atomics) agree to 1e-5 of their scale."""
    C, R = 32, 64
    sc = random_triplanes(2, C, R, seed=4).to(DEV)
    fx = {"space_cache": sc, **{k: v.to(DEV) for k, v in random_decoder(C, seed=1).items()}}
    geom, rend = build_plugins(fx, DEV, 32, 64)
    with torch.no_grad():
        s0, d0 = geom.forward_field_grid(64, sc)
        for _ in range(200):
            s1, d1 = geom.forward_field_grid(64, sc)
            assert torch.equal(s0, s1) and torch.equal(d0, d1)
    rays_o, rays_d, c2w, dist = [t.to(DEV) for t in camera_rays(4, 24, 24, seed=5, views_per_prompt=2)]
    rend.train()
    w = geom.decoder_weights()
    for p_ in w:
        p_.requires_grad_(True)
    first = None
    for _ in range(50):
        sc_g = sc.clone().requires_grad_(True)
        out = rend(rays_o, rays_d, None, torch.ones(3, device=DEV), space_cache=sc_g, text_embed=torch.zeros(2, 4, device=DEV),
                   camera_distances=dist, c2w=c2w)
        loss = out["comp_rgb"].square().sum() + out["opacity"].sum() + 0.1 * ((out["sdf_grad"].norm(dim=-1) - 1) ** 2).sum()
        grads = torch.autograd.grad(loss, [sc_g] + w)
        cur = {k: out[k].detach().clone() for k in ("comp_rgb", "opacity", "depth", "comp_normal", "weights", "sdf", "sdf_grad",
                                                    "features", "t_points")}
        if first is None:
            first, g_first = cur, [x.clone() for x in grads]
        else:
            for k in cur:
                assert torch.equal(cur[k], first[k]), k
            for a, b in zip(grads, g_first):
                assert rel_err(a, b) < 1e-5


@pytest.mark.parametrize("name", ["geometry_c8_r16", "geometry_c32_r16"])
def test_forward_field_is_differentiable(name):
    """forward_field with the deformable grid on (configs/TriplaneTurbo_v1.yaml:100): gradients reach the planes, the SDF
    decoder and the deformation decoder (few_step…diffusion.py:375-394, …mesh_rasterize_renderer.py:449-452)."""
    fx = load_golden(name, DEV)
    cpu = load_golden(name)
    geom, _ = build_plugins(fx, DEV)
    w_s, w_d = geom.sdf_network.weights(), geom.deformation_network.weights()
    for p_ in w_s + w_d:
        p_.requires_grad_(True)
    sc = fx["space_cache"].clone().requires_grad_(True)
    sdf, deform = geom.forward_field(fx["points"], sc)
    assert sdf.requires_grad and deform.requires_grad
    assert max_abs(sdf, fx["field_sdf"]) < TOL and max_abs(deform, fx["field_deformation"]) < TOL
    g = torch.Generator().manual_seed(11)
    cot_s, cot_d = torch.randn(sdf.shape, generator=g), torch.randn(deform.shape, generator=g)
    got = torch.autograd.grad((sdf * cot_s.to(DEV)).sum() + (deform * cot_d.to(DEV)).sum(), [sc] + w_s + w_d)
    wr = {k: [t.clone().requires_grad_(True) for t in v] for k, v in weights_from(cpu).items()}
    sc_c = cpu["space_cache"].clone().requires_grad_(True)
    s_ref, d_ref = rp.forward_field(cpu["points"], sc_c, wr, rp.PathConfig())
    want = torch.autograd.grad((s_ref * cot_s).sum() + (d_ref * cot_d).sum(), [sc_c] + wr["sdf"] + wr["deformation"])
    for i, (a, b) in enumerate(zip(got, want)):
        assert rel_err(a.cpu(), b) < GTOL, i


def test_interpolate_encodings_matches_reference():
    fx = load_golden("geometry_c32_r16", DEV)
    cpu = load_golden("geometry_c32_r16")
    geom, _ = build_plugins(fx, DEV)
    pts = geom.rescale_points(fx["points"])
    geo, tex = geom.interpolate_encodings(pts, fx["space_cache"])
    assert max_abs(geo, fx["enc_geo"]) < 1e-5 and max_abs(tex, fx["enc_tex"]) < 1e-5
    assert max_abs(geom.interpolate_encodings(pts, fx["space_cache"], only_geo=True), fx["enc_geo"]) < 1e-5
    # gradient w.r.t. the space cache through the differentiable repack
    sc = fx["space_cache"].clone().requires_grad_(True)
    geo, tex = geom.interpolate_encodings(pts, sc)
    g = torch.Generator().manual_seed(7)
    cg, ct = torch.randn(geo.shape, generator=g), torch.randn(tex.shape, generator=g)
    got, = torch.autograd.grad((geo * cg.to(DEV)).sum() + (tex * ct.to(DEV)).sum(), sc)
    sc_c = cpu["space_cache"].clone().requires_grad_(True)
    geo_r, tex_r = rp.interpolate_encodings(rp.rescale_points(cpu["points"], 1.0), sc_c)
    want, = torch.autograd.grad((geo_r * cg).sum() + (tex_r * ct).sum(), sc_c)
    assert rel_err(got.cpu(), want) < 1e-5
    # the decoder modules evaluate on their own tensors like the reference's VanillaMLP (networks.py:90-95)
    sdf_direct = geom.sdf_network(fx["enc_geo"].reshape(-1, fx["enc_geo"].shape[-1]))
    assert max_abs(sdf_direct.reshape(-1), fx["out_sdf_orig"].reshape(-1)) < TOL


def test_unsplit_space_cache_is_accepted():
    """the VAE decoder's raw [P,6,2C,R,R] output rendered directly (channel split folded into the repack) is
    bit-identical to `decode` + render, forward and gradient (zeros in the unused channel halves)."""
    fx = load_golden("render_train_c8", DEV)
    P, V, H, W, ns, nimp = [int(v) for v in fx["meta"][:6]]
    geom, rend = build_plugins(fx, DEV, ns, nimp)
    rend.train()
    sc = fx["space_cache"]
    C_ = sc.shape[2]
    g = torch.Generator(device=DEV).manual_seed(3)
    raw = torch.randn(P, 6, 2 * C_, sc.shape[3], sc.shape[4], device=DEV, generator=g)
    raw[:, 0:3, :C_] = sc[:, 0:3]
    raw[:, 3:6, C_:] = sc[:, 3:6]
    assert torch.equal(geom.decode(raw), sc)                       # the reference's split (few_step…:186-196)
    kw = dict(rays_o=fx["rays_o"], rays_d=fx["rays_d"], light_positions=None, bg_color=torch.ones(3, device=DEV),
              text_embed=torch.zeros(P, 4, device=DEV), camera_distances=fx["camera_distances"], c2w=fx["c2w"],
              t_starts=fx["t_starts"], t_ends=fx["t_ends"])
    a = sc.clone().requires_grad_(True)
    b = raw.clone().requires_grad_(True)
    oa, ob = rend(space_cache=a, **kw), rend(space_cache=b, **kw)
    for k in ("comp_rgb", "opacity", "depth", "sdf", "sdf_grad", "features", "weights"):
        assert torch.equal(oa[k], ob[k]), k
    ga, = torch.autograd.grad(oa["comp_rgb"].square().sum() + oa["opacity"].sum(), a)
    gb, = torch.autograd.grad(ob["comp_rgb"].square().sum() + ob["opacity"].sum(), b)
    assert rel_err(gb[:, 0:3, :C_], ga[:, 0:3]) < 1e-5 and rel_err(gb[:, 3:6, C_:], ga[:, 3:6]) < 1e-5
    assert float(gb[:, 0:3, C_:].abs().max()) == 0.0 and float(gb[:, 3:6, :C_].abs().max()) == 0.0
    geom.cfg.fuse_channel_split = True
    assert geom.decode(raw) is raw


def test_estimator_accepts_closures():
    """ImportanceEstimator.sampling with a plain callable (the reference's signature, estimators.py:22-101) takes the
    general route and lands on the same intervals as the fused route driven by the renderer's ProposalSpec."""
    from triplaneturbo_b200.renderer import ImportanceEstimator
    fx = load_golden("render_train_c8", DEV)
    P, V, H, W, ns, nimp = [int(v) for v in fx["meta"][:6]]
    geom, rend = build_plugins(fx, DEV, ns, nimp)
    w = geom.decoder_weights()
    o, d = fx["rays_o"].reshape(-1, 3).contiguous(), fx["rays_d"].reshape(-1, 3).contiguous()
    spec = ImportanceEstimator.ProposalSpec(ops.cached_planes(fx["space_cache"]),
                                            ops.cached_wpack(w[:3], w[3:], geom._deformation_weights(), fx["space_cache"].shape[2]),
                                            rend.path_scalars(), o, d, V * H * W)
    est = ImportanceEstimator()
    n = o.shape[0]
    f0, f1 = est.sampling([spec], [nimp], ns, n, 0.1, 4.0, "uniform", False)
    calls = []

    def closure(t_starts, t_ends):          # an opaque callable: evaluated as is
        calls.append(t_starts.shape)
        return spec(t_starts, t_ends)
    c0, c1 = est.sampling([closure], [nimp], ns, n, 0.1, 4.0, "uniform", False)
    assert calls == [(n, nimp)]
    assert c0.shape == f0.shape == (n, ns + nimp + 1)
    cpu = load_golden("render_train_c8")
    tv, cdf = proposal_cdf(cpu, rp.PathConfig(num_samples_per_ray=ns, num_samples_per_ray_importance=nimp))
    assert_intervals_close(torch.cat([c0, c1[:, -1:]], 1).cpu(), torch.cat([f0, f1[:, -1:]], 1).cpu(), tv.double(), cdf.double(),
                           tol=3e-5)
    assert_intervals_close(torch.cat([c0, c1[:, -1:]], 1).cpu(), torch.cat([cpu["t_starts"], cpu["t_ends"][:, -1:]], 1),
                           tv.double(), cdf.double(), tol=3e-5)


@pytest.mark.parametrize("precise", [False, True], ids=["bwd-1xTF32", "bwd-3xTF32"])
def test_backward_precision_against_fp64(kernel_family, precise):
    """Elementwise gradient error of the CUDA path and of the oracle run in plain fp32 torch on the same GPU, both against
    the oracle in fp64.  The backward layers of the tensor-core families are single-pass TF32 in the colour branch;
    this asserts that the error stays within a small multiple of what an fp32 run (fp32 atomics / summation order) has."""
    name = "render_train_c32"
    fx = load_golden(name, DEV)
    P, V, H, W, ns, nimp = [int(v) for v in fx["meta"][:6]]
    geom, rend = build_plugins(fx, DEV, ns, nimp, rgb_grad_shrink=float(fx["rgb_grad_shrink"]))
    rend.train()
    w = geom.decoder_weights()
    for p_ in w:
        p_.requires_grad_(True)
    cot_keys = [k[4:] for k in fx if k.startswith("cot_")]

    def loss_of(o, f):
        l_ = sum((o[k] * f["cot_" + k]).sum() for k in cot_keys)
        return l_ + 0.1 * ((torch.linalg.norm(o["sdf_grad"], ord=2, dim=-1) - 1.0) ** 2).sum()

    sc = fx["space_cache"].clone().requires_grad_(True)
    out = rend(space_cache=sc, t_starts=fx["t_starts"], t_ends=fx["t_ends"], rays_o=fx["rays_o"], rays_d=fx["rays_d"],
               light_positions=None, bg_color=torch.ones(3, device=DEV), text_embed=torch.zeros(P, 4, device=DEV),
               camera_distances=fx["camera_distances"], c2w=fx["c2w"])
    ops.set_precise_backward(precise)          # TT_FLAG_PRECISE_BWD: colour-decoder backward layers as 3xTF32
    try:
        ours = torch.autograd.grad(loss_of(out, fx), [sc] + w)
    finally:
        ops.set_precise_backward(False)

    def oracle_grads(f):
        pc = rp.PathConfig(num_samples_per_ray=ns, num_samples_per_ray_importance=nimp,
                           rgb_grad_shrink=float(f["rgb_grad_shrink"]))
        wr = {k: [t.clone().requires_grad_(True) for t in v] for k, v in weights_from(f).items() if k != "deformation"}
        s = f["space_cache"].clone().requires_grad_(True)
        o = rp.render_forward(f["rays_o"], f["rays_d"], s, wr, pc, torch.ones(3, dtype=s.dtype, device=s.device),
                              f["camera_distances"], f["c2w"], t_starts=f["t_starts"], t_ends=f["t_ends"])
        return torch.autograd.grad(loss_of(o, f), [s] + wr["sdf"] + wr["feature"])
    g64 = oracle_grads(load_golden(name, "cpu", torch.float64))
    g32 = oracle_grads(fx)                                   # plain fp32 torch on the GPU (cuBLAS / ATen, TF32 off)
    names = ["space_cache"] + [f"w_sdf_{i}" for i in range(3)] + [f"w_feature_{i}" for i in range(3)]
    table = {}
    for n_, a, b, c in zip(names, ours, g32, g64):
        scale = float(c.abs().max()) + 1e-30
        e_ours = float((a.double().cpu() - c).abs().max()) / scale
        e_fp32 = float((b.double().cpu() - c).abs().max()) / scale
        table[n_] = {"ours": e_ours, "fp32_torch": e_fp32}
        assert e_ours < GTOL, (n_, e_ours)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.path.join("gpurun_out", f"r02_backward_precision_impl{kernel_family}_{'3x' if precise else '1x'}.json"), "w") as fh:
        json.dump(table, fh, indent=1)
    # What the table shows: the geometry-side gradients (space cache, SDF decoder) are within 1.3x-6x of the
    # fp32 run's own error against fp64; the colour decoder's weight gradients sit at the single-pass TF32 level
    # (~2.5e-4 of the tensor's scale) where an fp32 run reaches 1e-6.  Bars: every tensor below 1e-3 of its scale, and the
    # tensors that feed the generator (d loss / d space_cache) within 3x of the fp32 run.
    assert all(v["ours"] < 1e-3 for v in table.values()), table
    assert table["space_cache"]["ours"] < 3.0 * table["space_cache"]["fp32_torch"] + 1e-5, table
    if precise:     # first and last layer of the feature network reach the fp32 run's level; W2's gradient keeps the
        for k in ("w_feature_0", "w_feature_2"):     # single-pass K-major contraction over the tile's points (6e-5)
            assert table[k]["ours"] < 1e-5, table
        assert table["w_feature_1"]["ours"] < 1e-4, table
