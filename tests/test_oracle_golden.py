"""Pin the oracle against golden vectors produced by the reference's own Python source
(tests/golden/make_golden.py).  CPU only."""
import pytest
import torch

from oracle import reference_path as rp
from tests.helpers import load_golden, weights_from, max_abs, rel_err

TOL = 2e-5       # same torch ops on the same machine class; slack for op-order in sums
GTOL = 2e-4      # gradients (relative to the tensor's max magnitude)


@pytest.mark.parametrize("name", ["geometry_c8_r16", "geometry_c32_r16"])
def test_geometry_matches_reference(name):
    fx = load_golden(name)
    w = weights_from(fx)
    for ws in w.values():
        for t in ws:
            t.requires_grad_(True)
    cfg = rp.PathConfig()
    sc = fx["space_cache"].clone().requires_grad_(True)
    out = rp.geometry_forward(fx["points"].clone(), sc, w, cfg, output_normal=True)
    for k in ("sdf", "sdf_orig", "features", "normal", "sdf_grad"):
        assert max_abs(out[k], fx["out_" + k]) < TOL, k
    loss = sum((out[k] * fx["cot_" + k]).sum() for k in ("sdf", "features", "normal", "sdf_grad"))
    params = w["sdf"] + w["feature"]
    grads = torch.autograd.grad(loss, [sc] + params)
    assert rel_err(grads[0], fx["grad_space_cache"]) < GTOL
    for i in range(3):
        assert rel_err(grads[1 + i], fx[f"grad_w_sdf_{i}"]) < GTOL
        assert rel_err(grads[4 + i], fx[f"grad_w_feature_{i}"]) < GTOL
    with torch.no_grad():
        sdf, deform = rp.forward_field(fx["points"], fx["space_cache"], w, cfg)
        assert max_abs(sdf, fx["field_sdf"]) < TOL
        assert max_abs(deform, fx["field_deformation"]) < TOL
        assert max_abs(rp.forward_sdf(fx["points"], fx["space_cache"], w, cfg), fx["forward_sdf"]) < TOL
        assert max_abs(rp.export_features(fx["points"][:1], fx["space_cache"][:1], w, cfg),
                       fx["export_features"]) < TOL
        geo, tex = rp.interpolate_encodings(rp.rescale_points(fx["points"], 1.0), fx["space_cache"])
        assert max_abs(geo, fx["enc_geo"]) < TOL and max_abs(tex, fx["enc_tex"]) < TOL
        assert torch.equal(rp.decode_split_channels(fx["triplane"]), fx["decoded"])


def _render(fx, training=True, jitters=None):
    P, V, H, W, ns, nimp = [int(v) for v in fx["meta"][:6]]
    cfg = rp.PathConfig(num_samples_per_ray=ns, num_samples_per_ray_importance=nimp,
                        normal_direction=fx["normal_direction"], rgb_grad_shrink=float(fx["rgb_grad_shrink"]),
                        cos_anneal_ratio=float(fx.get("cos_anneal_ratio", 1.0)))
    w = weights_from(fx)
    explicit_bg = bool(fx["meta"][7])
    bg = torch.ones(3) if explicit_bg else fx["bg"]
    sc = fx["space_cache"].clone()
    t0 = t1 = None
    if jitters is not None:
        B = P * V
        t0, t1 = rp.sample_intervals(fx["rays_o"], fx["rays_d"], sc.repeat_interleave(B // P, 0), w, cfg,
                                     stratified=True, jitters=jitters)
    return cfg, w, sc, bg, t0, t1


@pytest.mark.parametrize("name", ["render_train_c8", "render_train_c32", "render_train_front", "render_train_shrink",
                                  "render_train_cos_anneal", "render_train_variance"])
def test_render_training_matches_reference(name):
    fx = load_golden(name)
    cfg, w, sc, bg, _, _ = _render(fx)
    if name == "render_train_shrink":
        assert abs(cfg.rgb_grad_shrink - 0.307) < 1e-6          # C([0,1,0.01,20000]) at step 14000 (yaml:139)
    if name == "render_train_cos_anneal":
        assert cfg.cos_anneal_ratio == pytest.approx(0.4)
    var_p = None
    if name == "render_train_variance":     # LearnedVariance parameter p with inv_std = exp(10 p) (REN:24-35)
        var_p = torch.tensor(cfg.learned_variance_init, requires_grad=True)
        cfg.learned_variance_init = var_p
    sc.requires_grad_(True)
    for ws in w.values():
        for t in ws:
            t.requires_grad_(True)
    # (1) the sampler reproduces the intervals the reference's estimator produced
    B = fx["rays_o"].shape[0]
    with torch.no_grad():
        t0, t1 = rp.sample_intervals(fx["rays_o"], fx["rays_d"], sc.detach().repeat_interleave(B // sc.shape[0], 0),
                                     w, cfg)
    assert max_abs(t0, fx["t_starts"]) < 2e-6 and max_abs(t1, fx["t_ends"]) < 2e-6
    # (2) marching the reference's own intervals reproduces every output of the reference renderer
    out = rp.render_forward(fx["rays_o"], fx["rays_d"], sc, w, cfg, bg, fx["camera_distances"], fx["c2w"],
                            t_starts=fx["t_starts"], t_ends=fx["t_ends"])
    keys = [k[4:] for k in fx if k.startswith("out_") and k != "out_comp_rgb_bg"]
    assert "weights" in keys and "sdf_grad" in keys and "comp_normal" in keys
    for k in keys:
        if k == "ray_indices":
            assert torch.equal(out[k], fx["out_" + k])       # bit-exact contract
        else:
            assert max_abs(out[k], fx["out_" + k]) < TOL, k
    cot_keys = [k[4:] for k in fx if k.startswith("cot_")]
    loss = sum((out[k] * fx["cot_" + k]).sum() for k in cot_keys)
    loss = loss + 0.1 * ((torch.linalg.norm(out["sdf_grad"], ord=2, dim=-1) - 1.0) ** 2).sum()
    grads = torch.autograd.grad(loss, [sc] + w["sdf"] + w["feature"] + ([var_p] if var_p is not None else []))
    assert rel_err(grads[0], fx["grad_space_cache"]) < GTOL
    for i in range(3):
        assert rel_err(grads[1 + i], fx[f"grad_w_sdf_{i}"]) < GTOL
        assert rel_err(grads[4 + i], fx[f"grad_w_feature_{i}"]) < GTOL
    if var_p is not None:
        assert rel_err(grads[7], fx["grad_inv_std_param"]) < GTOL


def test_render_stratified_matches_reference():
    fx = load_golden("render_train_stratified")
    cfg, w, sc, bg, t0, t1 = _render(fx, jitters=[fx["jitter0"], fx["jitter1"]])
    out = rp.render_forward(fx["rays_o"], fx["rays_d"], sc, w, cfg, bg, fx["camera_distances"], fx["c2w"],
                            t_starts=t0, t_ends=t1)
    for k in ("t_points", "t_intervals", "comp_rgb", "opacity", "weights"):
        assert max_abs(out[k], fx["out_" + k]) < TOL, k


def test_render_eval_matches_reference():
    """Eval: one space cache, V views rendered view by view in 500-point chunks (REN:158-185,364-395)."""
    fx = load_golden("render_eval_c8")
    cfg, w, sc, bg, _, _ = _render(fx, training=False)
    with torch.no_grad():
        out = rp.render_forward(fx["rays_o"], fx["rays_d"], sc, w, cfg, bg, fx["camera_distances"], fx["c2w"],
                                training=False)
    for k in ("comp_rgb", "comp_rgb_fg", "opacity", "depth", "z_variance", "disparity", "comp_normal",
              "comp_normal_cam_vis", "comp_normal_cam_vis_white"):
        assert max_abs(out[k], fx["out_" + k]) < TOL, k
    assert "weights" not in out


def test_patch_renderer_matches_reference():
    fx = load_golden("patch_c8")
    P, V, H, W, ns, nimp, PS, ds = [int(v) for v in fx["meta"]]
    cfg = rp.PathConfig(num_samples_per_ray=ns, num_samples_per_ray_importance=nimp)
    w = weights_from(fx)

    def render_fn(o, d):
        with torch.no_grad():
            return rp.render_forward(o, d, fx["space_cache"], w, cfg, torch.ones(3), fx["camera_distances"],
                                     fx["c2w"])
    out = rp.patch_render(render_fn, fx["rays_o"], fx["rays_d"], PS, ds, tuple(int(v) for v in fx["patch_xy"]))
    for k in ("comp_rgb", "opacity", "depth", "disparity", "comp_normal", "comp_normal_cam_vis"):
        assert max_abs(out[k], fx["out_" + k]) < TOL, k
