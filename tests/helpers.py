"""Shared helpers for the test-suite: golden fixture loading and weight dictionaries."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name, device="cpu", dtype=None):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = {}
    for k in z.files:
        a = z[k]
        if a.dtype.kind in "US":
            out[k] = str(a)
        else:
            t = torch.from_numpy(a.copy())
            if dtype is not None and t.is_floating_point():
                t = t.to(dtype)
            out[k] = t.to(device)
    return out


def weights_from(fix, names=("sdf", "feature", "deformation")):
    return {n: [fix[f"w_{n}_{i}"] for i in range(3)] for n in names if f"w_{n}_0" in fix}


def max_abs(a, b):
    return (a.double() - b.double()).abs().max().item()


def rel_err(a, b):
    """max |a-b| / (max |b| + tiny): scale-aware error for gradient tensors."""
    return (a.double() - b.double()).abs().max().item() / (b.double().abs().max().item() + 1e-30)


def proposal_cdf(fx, pc, jitter0=None):
    """Oracle: coarse edges t [n, n_imp+1] and the proposal CDF at those edges (EST:72-88)."""
    from oracle import reference_path as rp, nerfacc_restated as nf
    w = weights_from(fx)
    o, d = fx["rays_o"].reshape(-1, 3), fx["rays_d"].reshape(-1, 3)
    n = o.shape[0]
    B = fx["rays_o"].shape[0]
    sc = fx["space_cache"].repeat_interleave(B // fx["space_cache"].shape[0], 0)
    nimp = pc.num_samples_per_ray_importance
    s = nf.quantiles(nimp, n, jitter0 is not None, jitter0, o)
    tv = rp.transform_stot(s, pc.near_plane, pc.far_plane)
    t0, t1 = tv[:, :-1], tv[:, 1:]
    pos = o[:, None] + d[:, None] * (t0 + t1)[..., None] / 2
    with torch.no_grad():
        g = rp.geometry_forward(pos.reshape(B, -1, 3), sc, w, pc)
        inv = rp.inv_std_of(pc.learned_variance_init, g["sdf"])
        sig = rp.proposal_density(g["sdf"], inv, pc.render_step_size).reshape(n, nimp)
        trans, _ = nf.render_transmittance_from_density(t0, t1, sig)
    return tv, 1 - torch.cat([trans, torch.zeros(n, 1)], 1)


def assert_intervals_close(t, t_ref, tv, cdf, tol=2e-5, cdf_tol=2e-5):
    """Sorted interval edges agree within `tol`, except where the proposal CDF is flat: there the inverse CDF is
    ill-conditioned (a 1-ulp change of the CDF moves the edge), so the edge only has to hit the same CDF value."""
    assert t.shape == t_ref.shape
    assert bool((t[:, 1:] >= t[:, :-1]).all()), "edges not sorted"
    bad = (t - t_ref).abs() > tol
    if not bad.any():
        return

    def F(x):
        idx = (torch.searchsorted(tv.contiguous(), x.contiguous(), right=True) - 1).clamp(0, tv.shape[1] - 2)
        a, b = torch.gather(tv, 1, idx), torch.gather(tv, 1, idx + 1)
        ca, cb = torch.gather(cdf, 1, idx), torch.gather(cdf, 1, idx + 1)
        return ca + (x - a) / (b - a) * (cb - ca)
    dF = (F(t.double()) - F(t_ref.double())).abs()
    assert bad.float().mean() < 0.01, f"{bad.sum().item()} of {bad.numel()} edges differ"
    assert bool((dF[bad] < cdf_tol).all()), f"edges differ beyond the flat-CDF allowance: {dF[bad].max().item()}"


def camera_rays(*a, **k):
    from triplaneturbo_b200.synthetic import camera_rays as f
    return f(*a, **k)


def random_decoder(*a, **k):
    from triplaneturbo_b200.synthetic import random_decoder as f
    return f(*a, **k)


def build_plugins(*a, **k):
    from triplaneturbo_b200.synthetic import build_plugins as f
    return f(*a, **k)
