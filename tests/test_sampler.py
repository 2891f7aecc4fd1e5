"""Stand-alone plane sampler (csrc/tt_sampler.cuh, triplaneturbo_b200/sampler.py) against the oracle's bilinear
restatement (oracle/bilinear.py), ATen's grid_sample on the CPU, and autograd's first and second derivatives of the
oracle.  CPU part: the kernel sources compiled for the host (tests/emul); GPU part: the package's functional API
(grid_sample_2d, grid_sample, sample_from_planes of custom/triplaneturbo/models/geometry/utils.py:21-24,127-161 and
extern/grid_sample_gradfix/cuda_gridsample.py:22-79)."""
import numpy as np
import pytest
import torch

from oracle import bilinear, reference_path as rp
from tests.helpers import max_abs, rel_err

TOL = 1e-5      # fp32 sums of 4 (12) products in a fixed order; 1e-4 is the path's tolerance


def _case(N, K, C_, H, W, M, seed, spread=1.15):
    g = torch.Generator().manual_seed(seed)
    planes = torch.randn(N * K, C_, H, W, generator=g)
    grid = (torch.rand(N * K, M, 2, generator=g) * 2 - 1) * spread          # some taps / points out of bounds
    grid[0, 0] = torch.tensor([-1.0, 1.0]); grid[0, 1] = torch.tensor([0.0, 0.0])      # corners, centre
    return planes, grid


def _oracle(planes, grid, K, concat):
    """[N*K,C,H,W], [N*K,M,2] -> [N,M,C] / [N,M,K*C] with the reference's permutes (GUT:137-145)."""
    NK, C_, H, W = planes.shape
    M = grid.shape[1]
    out = bilinear.grid_sample_2d_manual(planes, grid.unsqueeze(1))          # [NK,C,1,M]
    out = out.permute(0, 3, 2, 1).reshape(NK // K, K, M, C_)
    return out.permute(0, 2, 1, 3).reshape(NK // K, M, K * C_) if concat else out.sum(dim=1)


@pytest.fixture(scope="module")
def em():
    from tests.emul.emul_api import Emul
    e = Emul()
    return e


@pytest.mark.parametrize("N,K,C_,H,W,M,concat", [(2, 1, 8, 7, 9, 33, False), (1, 3, 40, 8, 8, 21, False),
                                                  (2, 3, 12, 6, 5, 17, True), (1, 4, 4, 5, 5, 9, False)])
def test_emulated_kernels_against_oracle(em, N, K, C_, H, W, M, concat):
    planes, grid = _case(N, K, C_, H, W, M, seed=11)
    cl = em.to_channel_last(planes.reshape(N * K, C_, H * W).numpy())
    assert np.array_equal(cl, planes.permute(0, 2, 3, 1).reshape(N * K, H * W, C_).numpy())
    assert np.array_equal(em.from_channel_last(cl), planes.reshape(N * K, C_, H * W).numpy())
    cl = cl.reshape(N * K, H, W, C_)
    # ATen agrees with the oracle restatement on the CPU (pins the oracle)
    aten = torch.nn.functional.grid_sample(planes, grid.unsqueeze(1), mode="bilinear", padding_mode="zeros", align_corners=False)
    assert max_abs(aten, bilinear.grid_sample_2d_manual(planes, grid.unsqueeze(1))) < TOL
    # forward
    p_ref, g_ref = planes.clone().requires_grad_(True), grid.clone().requires_grad_(True)
    ref = _oracle(p_ref, g_ref, K, concat)
    out = em.sample_fwd(cl, grid.numpy(), K, concat)
    assert max_abs(torch.from_numpy(out), ref) < TOL
    # first derivative
    gen = torch.Generator().manual_seed(5)
    go = torch.randn(ref.shape, generator=gen)
    gp_ref, gg_ref = torch.autograd.grad(ref, [p_ref, g_ref], go, create_graph=True)
    gp, gg = em.sample_bwd(cl, grid.numpy(), K, concat, go.numpy())
    assert max_abs(torch.from_numpy(gp).permute(0, 3, 1, 2), gp_ref) < TOL * 10
    assert rel_err(torch.from_numpy(gg), gg_ref) < 1e-5
    # second derivative: cotangents on (d/d planes, d/d grid) -> gradients w.r.t. (g_out, planes, grid)
    ggp, ggg = torch.randn(planes.shape, generator=gen), torch.randn(grid.shape, generator=gen)
    go_r = go.clone().requires_grad_(True)
    gp_r, gg_r = torch.autograd.grad(_oracle(p_ref, g_ref, K, concat), [p_ref, g_ref], go_r, create_graph=True)
    r_go, r_p, r_g = torch.autograd.grad([gp_r, gg_r], [go_r, p_ref, g_ref], [ggp, ggg])
    e_go, e_p, e_g = em.sample_bwdbwd(cl, grid.numpy(), K, concat, go.numpy(),
                                      ggp.permute(0, 2, 3, 1).contiguous().numpy(), ggg.numpy())
    assert rel_err(torch.from_numpy(e_go), r_go) < 1e-5
    assert rel_err(torch.from_numpy(e_p).permute(0, 3, 1, 2), r_p) < 1e-5
    assert rel_err(torch.from_numpy(e_g), r_g) < 1e-5


def test_empty_and_bad_arguments(em):
    cl = np.zeros((3, 4, 4, 8), np.float32)
    out = em.sample_fwd(cl, np.zeros((3, 0, 2), np.float32), 3, False)
    assert out.shape == (1, 0, 8)
    rc = em.L.tt_sample_planes_fwd(None, 1, 1, 8, 4, 4, None, 1, 0, None, None)
    assert rc != 0 and b"NULL" in em.L.tt_last_error()
    from tests.emul.emul_api import ptr, aligned_zeros
    a = aligned_zeros((1, 4, 4, 6))
    rc = em.L.tt_sample_planes_fwd(ptr(a), 1, 1, 6, 4, 4, ptr(np.zeros((1, 1, 2), np.float32)), 1, 0, ptr(aligned_zeros((1, 1, 6))), None)
    assert rc != 0 and b"multiple of 4" in em.L.tt_last_error()


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("interp,C_", [("v1", 32), ("v2", 40), ("v4", 8)])
def test_sample_from_planes_matches_reference_and_double_backward(interp, C_):
    import triplaneturbo_b200 as tt
    dev = "cuda:0"
    g = torch.Generator().manual_seed(3)
    N, M, R = 2, 301, 16
    pf = torch.randn(N, 3, C_, R, R, generator=g)
    xyz = (torch.rand(N, M, 3, generator=g) * 2 - 1) * 1.1
    pf_r, xyz_r = pf.clone().requires_grad_(True), xyz.clone().requires_grad_(True)
    feats = torch.tanh(pf_r) if interp == "v4" else pf_r
    ref = rp.sample_from_planes(feats, xyz_r, "v1" if interp == "v4" else interp, sampler=lambda i, gr: bilinear.grid_sample_2d_manual(i, gr))
    pf_g, xyz_g = pf.to(dev).requires_grad_(True), xyz.to(dev).requires_grad_(True)
    out = tt.sample_from_planes(pf_g, xyz_g, interpolate_feat=interp)
    assert out.shape == ref.shape and max_abs(out.cpu(), ref) < TOL
    go = torch.randn(ref.shape, generator=g)
    # first derivative, kept differentiable; then a scalar of it (the eikonal-style use: |d out / d xyz|)
    d_ref = torch.autograd.grad(ref, xyz_r, go, create_graph=True)[0]
    d_gpu = torch.autograd.grad(out, xyz_g, go.to(dev), create_graph=True)[0]
    assert rel_err(d_gpu.cpu(), d_ref) < 1e-5
    r_p, r_x = torch.autograd.grad((d_ref ** 2).sum(), [pf_r, xyz_r])
    g_p, g_x = torch.autograd.grad((d_gpu ** 2).sum(), [pf_g, xyz_g])
    assert rel_err(g_p.cpu(), r_p) < 2e-5 and rel_err(g_x.cpu(), r_x) < 2e-5


@pytest.mark.gpu
def test_grid_sample_2d_api_and_large_random_property():
    import triplaneturbo_b200 as tt
    dev = "cuda:0"
    g = torch.Generator().manual_seed(9)
    x = torch.randn(3, 12, 10, 14, generator=g)
    grid = (torch.rand(3, 5, 7, 2, generator=g) * 2 - 1) * 1.2
    ref = torch.nn.functional.grid_sample(x, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    out = tt.grid_sample_2d(x.to(dev), grid.to(dev), padding_mode="zeros", align_corners=False)
    assert out.shape == ref.shape and max_abs(out.cpu(), ref) < TOL
    assert max_abs(tt.grid_sample(x.to(dev), grid.to(dev)).cpu(), ref) < TOL
    with pytest.raises(NotImplementedError):
        tt.grid_sample_2d(x.to(dev), grid.to(dev), padding_mode="border")
    from triplaneturbo_b200 import _cabi
    with pytest.raises(_cabi.TTError):
        tt.grid_sample_2d(x, grid)                              # CPU tensors: no fallback
    # full-size linearity: sampling a sum of planes = sum of the samples (R=256, C=40, 1 M points)
    a = torch.randn(3, 256, 256, 40, device=dev, generator=torch.Generator(dev).manual_seed(1))
    b = torch.randn(3, 256, 256, 40, device=dev, generator=torch.Generator(dev).manual_seed(2))
    gr = torch.rand(3, 1 << 20, 2, device=dev, generator=torch.Generator(dev).manual_seed(3)) * 2 - 1
    from triplaneturbo_b200.sampler import sample_planes
    lhs = sample_planes(a + b, gr, 3, False)
    rhs = sample_planes(a, gr, 3, False) + sample_planes(b, gr, 3, False)
    assert max_abs(lhs, rhs) < 1e-4


@pytest.mark.gpu
def test_second_derivative_against_the_reference_kernel():
    """tt_sample_planes_bwdbwd vs the reference's own CUDA kernel (grid_sampler_2d_grad2_kernel,
    extern/grid_sample_gradfix/gridsample_cuda.cu:27-210), built by oracle/build_ref.py into oracle/_ref/."""
    import importlib.util
    import os
    so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle", "_ref", "gridsample_grad2_ref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/gridsample_grad2_ref.so not built (needs /root/reference at build time)")
    spec = importlib.util.spec_from_file_location("gridsample_grad2_ref", so)
    ref_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_mod)
    from triplaneturbo_b200.ops import _lib, _ptr, _stream
    from triplaneturbo_b200 import _cabi
    dev = "cuda:0"
    g = torch.Generator().manual_seed(21)
    NK, C_, H, W, M = 3, 32, 16, 12, 4099
    inp = torch.randn(NK, C_, H, W, generator=g).to(dev)
    grid = ((torch.rand(NK, 1, M, 2, generator=g) * 2 - 1) * 1.1).to(dev)
    go = torch.randn(NK, C_, 1, M, generator=g).to(dev)
    ggI = torch.randn(NK, C_, H, W, generator=g).to(dev)
    ggG = torch.randn(NK, 1, M, 2, generator=g).to(dev)
    r_ggo, r_gi, r_gg = ref_mod.grad2_2d(ggI, ggG, go, inp, grid, False, False)      # padding zeros, align_corners False
    cl = lambda t: t.permute(0, 2, 3, 1).contiguous()
    planes, ggp = cl(inp), cl(ggI)
    go_pm = go[:, :, 0, :].permute(0, 2, 1).contiguous()                              # [NK, M, C]
    e_ggo = torch.empty_like(go_pm); e_gp = torch.zeros_like(planes); e_gg = torch.empty(NK, M, 2, device=dev)
    L = _lib()
    _cabi.check(L, L.tt_sample_planes_bwdbwd(_ptr(planes), NK, 1, C_, H, W, _ptr(grid.reshape(NK, M, 2).contiguous()), M, 0,
                                             _ptr(go_pm), _ptr(ggp), _ptr(ggG.reshape(NK, M, 2).contiguous()), _ptr(e_ggo),
                                             _ptr(e_gp), _ptr(e_gg), _stream(torch.device(dev))), "bwdbwd")
    torch.cuda.synchronize()
    assert rel_err(e_ggo.permute(0, 2, 1).unsqueeze(2), r_ggo) < 1e-5
    assert rel_err(e_gp.permute(0, 3, 1, 2), r_gi) < 1e-5
    assert rel_err(e_gg.reshape(NK, 1, M, 2), r_gg) < 1e-5
