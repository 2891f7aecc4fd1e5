#!/usr/bin/env python
"""BASELINE MEASUREMENT (not product code): the reference's algorithm as plain PyTorch on the GPU.

The reference renderer itself cannot run here (nerfacc / tiny-cuda-nn / lightning / omegaconf are absent), so the
"reference GPU renderer" of BASELINE.json's target is approximated by the oracle (oracle/reference_path.py: the
reference's operations one by one — NCHW planes, rotation copies, per-plane grid_sample materialising [3B,C,1,N],
nn.Linear-style matmuls (cuBLAS), cumprod / index_add compositing) executed on cuda:0.  Bilinear sampling is the oracle's
gather-based op (twice differentiable through autograd), or - when oracle/_ref/gridsample_grad2_ref.so exists - the
reference's own sampler stack: ATen grid_sample forward, aten::grid_sampler_2d_backward, and the reference's grad2_2d
CUDA kernel for the backward of the backward (the structure of extern/grid_sample_gradfix/cuda_gridsample.py:22-79).
One fwd+bwd step on a bounded sample of config 2 (1 prompt x 1 view x HxH rays); prints one JSON line.
    python tests/tools/bench_reference_gpu.py [H=96] [manual|ref]"""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import reference_path as rp
from triplaneturbo_b200.synthetic import camera_rays, random_decoder, random_triplanes

H = int(sys.argv[1]) if len(sys.argv) > 1 else 96
REF_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "oracle", "_ref", "gridsample_grad2_ref.so")
mode = sys.argv[2] if len(sys.argv) > 2 else ("ref" if os.path.exists(REF_SO) else "manual")
if mode == "ref":
    import importlib.util
    spec = importlib.util.spec_from_file_location("gridsample_grad2_ref", REF_SO)
    ref_mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref_mod)

    class _Fwd(torch.autograd.Function):
        @staticmethod
        def forward(ctx, inp, grid):
            ctx.save_for_backward(inp, grid)
            return torch.nn.functional.grid_sample(inp, grid, mode="bilinear", padding_mode="zeros", align_corners=False)

        @staticmethod
        def backward(ctx, go):
            inp, grid = ctx.saved_tensors
            return _Bwd.apply(go.contiguous(), inp, grid)

    class _Bwd(torch.autograd.Function):
        @staticmethod
        def forward(ctx, go, inp, grid):
            ctx.save_for_backward(go, inp, grid)
            return torch.ops.aten.grid_sampler_2d_backward(go, inp, grid, 0, 0, False, (True, True))

        @staticmethod
        def backward(ctx, ggi, ggg):
            go, inp, grid = ctx.saved_tensors
            ggi = torch.zeros_like(inp) if ggi is None else ggi.contiguous()
            ggg = torch.zeros_like(grid) if ggg is None else ggg.contiguous()
            o = ref_mod.grad2_2d(ggi, ggg, go, inp, grid, False, False)
            return o[0], o[1], o[2]

    rp.grid_sample_2d_manual = lambda i, g_: _Fwd.apply(i.contiguous(), g_.contiguous())
C, R, ns, nimp = 40, 256, 96, 192
dev = "cuda:0"
sc = random_triplanes(1, C, R, seed=0).to(dev).requires_grad_(True)
wts = random_decoder(C, seed=1)
w = {n: [wts[f"w_{n}_{i}"].to(dev).requires_grad_(True) for i in range(3)] for n in ("sdf", "feature")}
rays_o, rays_d, c2w, dist = [t.to(dev) for t in camera_rays(1, H, H, seed=2)]
pc = rp.PathConfig(num_samples_per_ray=ns, num_samples_per_ray_importance=nimp)
g = torch.Generator().manual_seed(3)
cots = {k: torch.randn(1, H, H, d, generator=g).to(dev) for k, d in (("comp_rgb", 3), ("comp_normal_cam_vis", 3), ("disparity", 1))}


def step():
    out = rp.render_forward(rays_o, rays_d, sc, w, pc, torch.ones(3, device=dev), dist, c2w)
    loss = sum((out[k] * cots[k]).sum() for k in cots)
    loss = loss + 0.1 * ((out["sdf_grad"].norm(dim=-1) - 1.0) ** 2).mean()
    loss = loss + 0.5 * torch.sqrt(out["opacity"] ** 2 + 0.01).mean()
    torch.autograd.grad(loss, [sc] + w["sdf"] + w["feature"])


for _ in range(2):
    step()
torch.cuda.synchronize()
t = time.perf_counter()
reps = 3
for _ in range(reps):
    step()
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / reps
print(json.dumps({"baseline": "oracle (plain PyTorch restatement of the reference) on cuda:0", "sampler": "ATen grid_sample + the reference's grad2_2d kernel" if mode == "ref" else "gather-based manual bilinear", "rays": H * H, "samples_per_ray": ns + nimp + 1,
                  "ms_per_step": dt * 1e3, "rays_per_s": H * H / dt, "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9,
                  "workload": f"1 prompt x 1 view x {H}x{H} rays of config2 (R={R}, C={C})"}))
