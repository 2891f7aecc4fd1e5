"""Generate golden vectors by running the REFERENCE's own Python source on CPU.

    python tests/golden/make_golden.py            # needs /root/reference; writes tests/golden/*.npz

The reference classes are imported through ``ref_harness`` (real module files, third-party packages stubbed;
see its docstring for which stubs carry arithmetic).  The fixtures hold inputs AND outputs, so the tests that
consume them (``tests/test_oracle_golden.py``, ``tests/test_gpu_parity.py``) never need ``/root/reference``.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness  # noqa: E402

ns = ref_harness.load()
ts = sys.modules["threestudio"]


class ConstBackground(ns.background_base.BaseBackground):
    """Stands in for the out-of-scope hash-grid background: returns a fixed per-ray colour tensor."""

    def configure(self):
        self.color = None

    def forward(self, dirs, **kw):
        if self.color is None:  # eval: constant white, like the reference's `eval_color` (yaml :117)
            return torch.ones_like(dirs)
        return self.color.to(dirs).reshape(*dirs.shape[:-1], 3)


def build(C, n_samples, n_imp, normal_direction="camera", rgb_grad_shrink=1.0, seed=1, trainable_variance=False,
          update_step=(0, 0)):
    torch.manual_seed(seed)
    geometry = ns.geometry.StableDiffusionTriplaneDualAttention(dict(
        radius=1.0, normal_type="analytic", sdf_bias="sphere", sdf_bias_params=0.5, rotate_planes="v1",
        split_channels="v1", geo_interpolate="v1", tex_interpolate="v2",
        space_generator_config={"output_dim": 2 * C}, isosurface_deformable_grid=True))
    material = ns.no_material.NoMaterial(dict(n_output_dims=3, color_activation="sigmoid-mipnerf",
                                              requires_normal=True))
    background = ConstBackground({})
    renderer = ns.renderer.GenerativeSpaceSDFVolumeRenderer(dict(
        radius=1.0, use_volsdf=False, trainable_variance=trainable_variance, learned_variance_init=0.4605,
        rgb_grad_shrink=rgb_grad_shrink, estimator="importance", num_samples_per_ray=n_samples,
        num_samples_per_ray_importance=n_imp, near_plane=0.1, far_plane=4.0, train_chunk_size=0,
        randomized=False, normal_direction=normal_direction, eval_chunk_size=500),
        geometry=geometry, material=material, background=background)
    renderer.update_step(*update_step)
    # decoder weights a bit larger than nn.Linear's default so that the SDF crosses zero in the volume
    with torch.no_grad():
        for net in (geometry.sdf_network, geometry.feature_network, geometry.deformation_network):
            for p in net.parameters():
                p.mul_(1.5)
    return geometry, renderer, background


def weights_of(geometry):
    out = {}
    for name, net in (("sdf", geometry.sdf_network), ("feature", geometry.feature_network),
                      ("deformation", geometry.deformation_network)):
        for i, idx in enumerate((0, 2, 4)):
            out[f"w_{name}_{i}"] = net.layers[idx].weight.detach().numpy().copy()
    return out


def cameras(B, H, W, seed, fovy_deg=60.0):
    """Rays as the reference's data module builds them (custom/triplaneturbo/data/…multistep_v2.py:250-337)."""
    g = torch.Generator().manual_seed(seed)
    elev = torch.rand(B, generator=g) * 30.0
    azim = (torch.rand(B, generator=g) + torch.arange(B)) / B * 360.0 - 180.0
    fovy = torch.full((B,), fovy_deg) * np.pi / 180
    dist = (torch.rand(B, generator=g) * 0.2 + 0.8) / torch.tan(0.5 * fovy)
    el, az = elev * np.pi / 180, azim * np.pi / 180
    pos = torch.stack([dist * torch.cos(el) * torch.cos(az), dist * torch.cos(el) * torch.sin(az),
                       dist * torch.sin(el)], -1).float()
    up = torch.tensor([0.0, 0.0, 1.0])[None].repeat(B, 1)
    lookat = F.normalize(-pos, dim=-1)
    right = F.normalize(torch.linalg.cross(lookat, up), dim=-1)
    up = F.normalize(torch.linalg.cross(right, lookat), dim=-1)
    c2w3x4 = torch.cat([torch.stack([right, up, -lookat], dim=-1), pos[:, :, None]], dim=-1)
    c2w = torch.cat([c2w3x4, torch.zeros_like(c2w3x4[:, :1])], dim=1)
    c2w[:, 3, 3] = 1.0
    dirs = ns.ops.get_ray_directions(H, W, focal=1.0)[None].repeat(B, 1, 1, 1)
    focal = 0.5 * H / torch.tan(0.5 * fovy)
    dirs[..., :2] = dirs[..., :2] / focal[:, None, None, None]
    rays_o, rays_d = ns.ops.get_rays(dirs, c2w, keepdim=True)
    return rays_o.contiguous(), rays_d.contiguous(), c2w, dist


def np_(d):
    return {k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def case_geometry(name, C, R, B, N, seed):
    geometry, _, _ = build(C, 8, 16, seed=seed)
    g = torch.Generator().manual_seed(seed + 100)
    space_cache = (torch.randn(B, 6, C, R, R, generator=g) * 0.5).requires_grad_(True)
    points = torch.rand(B, N, 3, generator=g) * 2.2 - 1.1  # some points outside the box -> zero padding
    out = geometry(points.clone(), space_cache, output_normal=True)
    cot = {k: torch.randn(out[k].shape, generator=g) for k in ("sdf", "features", "normal", "sdf_grad")}
    loss = sum((out[k] * cot[k]).sum() for k in cot)
    params = list(geometry.sdf_network.parameters()) + list(geometry.feature_network.parameters())
    grads = torch.autograd.grad(loss, [space_cache] + params)
    with torch.no_grad():
        sdf_f, def_f = geometry.forward_field(points, space_cache)
        sdf_s = geometry.forward_sdf(points, space_cache)
        feat_e = geometry.export(points[:1], space_cache[:1])["features"]
        enc_geo, enc_tex = geometry.interpolate_encodings(geometry.rescale_points(points), space_cache)
        triplane = torch.randn(B, 6, 2 * C, 4, 4, generator=g)
        decoded = geometry.decode(triplane)  # generator shell's forward_decode is the identity
    fix = dict(space_cache=space_cache, points=points, **{"out_" + k: v for k, v in out.items()},
               **{"cot_" + k: v for k, v in cot.items()}, grad_space_cache=grads[0],
               field_sdf=sdf_f, field_deformation=def_f, forward_sdf=sdf_s, export_features=feat_e,
               enc_geo=enc_geo, enc_tex=enc_tex, triplane=triplane, decoded=decoded)
    names = [f"grad_w_sdf_{i}" for i in range(3)] + [f"grad_w_feature_{i}" for i in range(3)]
    fix.update(dict(zip(names, grads[1:])))
    fix = np_(fix)
    fix.update(weights_of(geometry))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **fix)
    print(name, {k: v.shape for k, v in fix.items() if k.startswith("out_")})


def case_render(name, C, R, P, V, H, W, n_samples, n_imp, seed, training=True, normal_direction="camera",
                rgb_grad_shrink=1.0, stratified=False, explicit_bg=False, update_step=(0, 0), cos_anneal_ratio=None,
                trainable_variance=False):
    geometry, renderer, background = build(C, n_samples, n_imp, normal_direction, rgb_grad_shrink, seed=seed,
                                           trainable_variance=trainable_variance, update_step=update_step)
    if cos_anneal_ratio is not None:      # the attribute get_alpha reads (neus_volume_renderer.py:91,100-103)
        renderer.cos_anneal_ratio = cos_anneal_ratio
    g = torch.Generator().manual_seed(seed + 200)
    B = P * V
    space_cache = (torch.randn(P, 6, C, R, R, generator=g) * 0.5).requires_grad_(True)
    rays_o, rays_d, c2w, dist = cameras(B, H, W, seed + 300)
    bg = torch.rand(B, H, W, 3, generator=g)
    background.color = bg if training else None
    text_embed = torch.zeros(P, 4)  # only its batch size is read (REN:110,225)
    renderer.train(training)
    jit = None
    if stratified:
        renderer.randomized = True
        jit = [torch.rand(B * H * W, generator=g), torch.rand(B * H * W, generator=g)]
        queue = list(jit)
        # the nerfacc stub pulls one jitter per importance_sampling call
        orig = ns.pdf.importance_sampling

        def with_jitter(intervals, cdfs, n, strat=False):
            ns.pdf.jitter_for_next_call = queue.pop(0) if strat else None
            return orig(intervals, cdfs, n, strat)
        sys.modules["threestudio.models.estimators"].importance_sampling = with_jitter
    # record the intervals the reference's estimator hands to the marcher (EST:98-101)
    rec = {}
    sampling_orig = renderer.estimator.sampling

    def sampling_rec(*a, **k):
        rec["t_starts"], rec["t_ends"] = sampling_orig(*a, **k)
        return rec["t_starts"], rec["t_ends"]
    renderer.estimator.sampling = sampling_rec
    kw = dict(rays_o=rays_o, rays_d=rays_d, light_positions=torch.zeros(B, 3), space_cache=space_cache,
              text_embed=text_embed, camera_distances=dist, c2w=c2w)
    if explicit_bg:
        kw["bg_color"] = torch.ones(3)
    if training:
        out = renderer(**kw)
    else:
        assert P == 1
        with torch.no_grad():
            out = renderer(**kw)
    if stratified:
        sys.modules["threestudio.models.estimators"].importance_sampling = orig
    fix = dict(space_cache=space_cache, rays_o=rays_o, rays_d=rays_d, c2w=c2w, camera_distances=dist, bg=bg,
               meta=np.array([P, V, H, W, n_samples, n_imp, int(training), int(explicit_bg)]),
               normal_direction=np.array(normal_direction), rgb_grad_shrink=np.array(float(renderer.rgb_grad_shrink)),
               cos_anneal_ratio=np.array(float(renderer.cos_anneal_ratio)),
               trainable_variance=np.array(int(trainable_variance)))
    if jit is not None:
        fix["jitter0"], fix["jitter1"] = jit
    if training:
        fix["t_starts"], fix["t_ends"] = rec["t_starts"], rec["t_ends"]
    for k, v in out.items():
        if torch.is_tensor(v):
            fix["out_" + k] = v
    if training:
        cot = {k: torch.randn(out[k].shape, generator=g) for k in
               ("comp_rgb", "comp_normal_cam_vis" if normal_direction == "camera" else "comp_normal_cam_vis_white",
                "disparity", "opacity", "z_variance", "depth")}
        loss = sum((out[k] * cot[k]).sum() for k in cot)
        loss = loss + 0.1 * ((torch.linalg.norm(out["sdf_grad"], ord=2, dim=-1) - 1.0) ** 2).sum()  # eikonal
        params = list(geometry.sdf_network.parameters()) + list(geometry.feature_network.parameters())
        if trainable_variance:
            params = params + [renderer.variance._inv_std]
        grads = torch.autograd.grad(loss, [space_cache] + params)
        if trainable_variance:
            fix["grad_inv_std_param"] = grads[-1]
            grads = grads[:-1]
        fix.update({"cot_" + k: v for k, v in cot.items()})
        fix["grad_space_cache"] = grads[0]
        names = [f"grad_w_sdf_{i}" for i in range(3)] + [f"grad_w_feature_{i}" for i in range(3)]
        fix.update(dict(zip(names, grads[1:])))
    fix = np_(fix)
    fix.update(weights_of(geometry))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **fix)
    print(name, "rays", B * H * W, "keys", len(fix))


def case_patch(name, seed):
    C, R, P, V, H, W = 8, 16, 1, 2, 12, 12
    geometry, renderer, background = build(C, 8, 16, seed=seed)
    patch = ns.patch.PatchRenderer.__new__(ns.patch.PatchRenderer)
    torch.nn.Module.__init__(patch)
    from types import SimpleNamespace
    patch.cfg = SimpleNamespace(patch_size=4, global_downsample=3, global_detach=False)
    patch.base_renderer = renderer
    g = torch.Generator().manual_seed(seed + 200)
    B = P * V
    space_cache = torch.randn(P, 6, C, R, R, generator=g) * 0.5
    rays_o, rays_d, c2w, dist = cameras(B, H, W, seed + 300)
    renderer.train(True)
    torch.manual_seed(seed)  # PATCH:66-67 draws patch_x, patch_y with torch.randint
    state = torch.get_rng_state()
    px = torch.randint(0, W - 4, (1,)).item()
    py = torch.randint(0, H - 4, (1,)).item()
    torch.set_rng_state(state)
    with torch.no_grad():
        out = patch(rays_o, rays_d, torch.zeros(B, 3), torch.ones(3), space_cache=space_cache,
                    text_embed=torch.zeros(P, 4), camera_distances=dist, c2w=c2w)
    fix = dict(space_cache=space_cache, rays_o=rays_o, rays_d=rays_d, c2w=c2w, camera_distances=dist,
               patch_xy=np.array([px, py]), meta=np.array([P, V, H, W, 8, 16, 4, 3]))
    for k in ("comp_rgb", "opacity", "depth", "disparity", "comp_normal", "comp_normal_cam_vis"):
        fix["out_" + k] = out[k]
    fix = np_(fix)
    fix.update(weights_of(geometry))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **fix)
    print(name, "patch at", px, py)


if __name__ == "__main__":
    torch.set_num_threads(8)
    only = set(sys.argv[1:])      # optional: regenerate only the named fixtures
    if only:
        _cg, _cr, _cp = case_geometry, case_render, case_patch
        case_geometry = lambda name, **k: _cg(name, **k) if name in only else None      # noqa: E731
        case_render = lambda name, **k: _cr(name, **k) if name in only else None        # noqa: E731
        case_patch = lambda name, **k: _cp(name, **k) if name in only else None         # noqa: E731
    case_geometry("geometry_c8_r16", C=8, R=16, B=2, N=257, seed=11)
    case_geometry("geometry_c32_r16", C=32, R=16, B=1, N=128, seed=12)
    case_render("render_train_c8", C=8, R=16, P=2, V=2, H=5, W=6, n_samples=16, n_imp=32, seed=21)
    case_render("render_train_c32", C=32, R=16, P=1, V=2, H=4, W=4, n_samples=64, n_imp=128, seed=22,
                rgb_grad_shrink=[0, 1, 0.01, 20000], explicit_bg=True)
    case_render("render_train_front", C=8, R=16, P=1, V=2, H=4, W=4, n_samples=8, n_imp=16, seed=23,
                normal_direction="front")
    case_render("render_train_stratified", C=8, R=16, P=1, V=1, H=4, W=4, n_samples=8, n_imp=16, seed=24,
                stratified=True)
    case_render("render_eval_c8", C=8, R=16, P=1, V=3, H=4, W=5, n_samples=16, n_imp=32, seed=25,
                training=False, explicit_bg=True)
    case_patch("patch_c8", seed=26)
    # round 2: the parameters the shipped YAML actually trains with
    case_render("render_train_shrink", C=8, R=16, P=1, V=2, H=4, W=5, n_samples=16, n_imp=32, seed=27,
                rgb_grad_shrink=[0, 1, 0.01, 20000], update_step=(0, 14000))          # s = 0.307 (yaml:139)
    case_render("render_train_cos_anneal", C=8, R=16, P=1, V=2, H=4, W=4, n_samples=8, n_imp=16, seed=28,
                cos_anneal_ratio=0.4)
    case_render("render_train_variance", C=8, R=16, P=1, V=2, H=4, W=4, n_samples=16, n_imp=32, seed=29,
                trainable_variance=True)                                               # yaml:135-137 default of the class
