#!/usr/bin/env python
"""Benchmark of the triplane volume-rendering hot path: rendered rays/s, forward + backward.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload config2|config3|config1]

A step = one pass of the path over one batch of synthetic input on every rank: importance sampler -> fused
march (forward) -> loss on the rendered images + eikonal + sparsity -> backward to the triplanes and decoder
weights (-> all-reduce of the decoder-weight gradients when N > 1).  Prompts shard across ranks with no data-path
collective (weak scaling: every rank renders its own P prompts).

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how `roofline` and `cpu_baseline` are defined.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: batch=4 prompts, 256^2 x 40ch triplanes, 4 views 256^2, 96 samples/ray, 1xB200
    "config2": dict(P=4, V=4, H=256, W=256, R=256, C=40, ns=96, nimp=192),
    # configs[2] per-GPU share: 4 prompts, 4 views 512^2, C=32, 64+128 samples
    "config3": dict(P=4, V=4, H=512, W=512, R=256, C=32, ns=64, nimp=128),
    # configs[0]: single 64^2 x 32ch triplane, 1 cam 128^2, 64 samples/ray
    "config1": dict(P=1, V=1, H=128, W=128, R=64, C=32, ns=64, nimp=128),
}
METRIC = "rendered rays/sec (fwd+bwd) at 256^3 triplane"
LAMBDA_EIK, LAMBDA_SPARSITY = 0.1, 0.5


def mlp_flops(C, S, nimp):
    """Dense decoder FLOPs per ray (2 x MAC), SURVEY.md 8(d)."""
    f_sdf = 2 * (64 * C + 64 * 64 + 64)
    f_feat = 2 * (192 * C + 64 * 64 + 192)
    return dict(f_sdf=f_sdf, f_feat=f_feat,
                sample=nimp * f_sdf, fwd=S * (2 * f_sdf + f_feat), bwd_geo=S * 4 * f_sdf, bwd_tex=S * 3 * f_feat,
                step_survey=nimp * f_sdf + S * (2 * f_sdf + f_feat) + S * (2 * f_feat + 4 * f_sdf))


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        rows = [r for ts, r in self.rows if t0 <= ts <= t1 + 0.3] or [r for _, r in self.rows]
        sm, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU oracle arm
def oracle_step_fn(wl, sample_hw):
    """One fwd+bwd of the CPU oracle on a bounded sample of the workload: 1 prompt, 1 view, sample_hw^2 rays."""
    from oracle import reference_path as rp
    from triplaneturbo_b200.synthetic import camera_rays, random_decoder, random_triplanes
    C, R, ns, nimp = wl["C"], wl["R"], wl["ns"], wl["nimp"]
    sc = random_triplanes(1, C, R, seed=0).requires_grad_(True)
    wts = random_decoder(C, seed=1)
    w = {n: [wts[f"w_{n}_{i}"].requires_grad_(True) for i in range(3)] for n in ("sdf", "feature")}
    rays_o, rays_d, c2w, dist = camera_rays(1, sample_hw, sample_hw, seed=2)
    pc = rp.PathConfig(num_samples_per_ray=ns, num_samples_per_ray_importance=nimp)
    g = torch.Generator().manual_seed(3)
    cots = {k: torch.randn(1, sample_hw, sample_hw, d, generator=g) for k, d in
            (("comp_rgb", 3), ("comp_normal_cam_vis", 3), ("disparity", 1))}

    def step():
        out = rp.render_forward(rays_o, rays_d, sc, w, pc, torch.ones(3), dist, c2w)
        loss = sum((out[k] * cots[k]).sum() for k in cots)
        loss = loss + LAMBDA_EIK * ((out["sdf_grad"].norm(dim=-1) - 1.0) ** 2).mean()
        loss = loss + LAMBDA_SPARSITY * torch.sqrt(out["opacity"] ** 2 + 0.01).mean()
        torch.autograd.grad(loss, [sc] + w["sdf"] + w["feature"])
        return float(loss.detach())
    return step, sample_hw * sample_hw


def time_oracle(wl, sample_hw, steps, warmup):
    torch.set_num_threads(os.cpu_count())
    step, n_rays = oracle_step_fn(wl, sample_hw)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t = time.perf_counter(); step(); ts.append(time.perf_counter() - t)
    return n_rays, ts


def reference_arm(args, wl, rank):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (the reference itself
    cannot be imported: nerfacc/tinycudann/lightning are absent, nerfacc has no CPU kernels), all host threads."""
    if rank != 0:
        return
    hw = 32
    n_rays, ts = time_oracle(wl, hw, max(args.steps, 1), min(args.warmup, 1))
    total = sum(ts)
    val = n_rays * len(ts) / total
    sample = f"1 prompt x 1 view x {hw}x{hw} rays of {args.workload} (R={wl['R']}, C={wl['C']}, S={wl['ns'] + wl['nimp'] + 1}) per step"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "rays/s", "n_gpus": args.gpus, "steps": len(ts),
        "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * total / len(ts), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, **wl, "sample": sample},
        "cpu_baseline": {"value": val, "unit": "rays/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true", help="for ncu: 1 warm-up + K steps, no e2e / cpu legs (not a bench value)")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return reference_arm(args, wl, rank)
    args.warmup = 1 if args.profile else max(args.warmup, 3)

    # Libraries (NCCL prints its version banner) may write to fd 1: keep stdout for the ONE JSON line by pointing fd 1
    # at stderr until the result is printed.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import triplaneturbo_b200 as tt  # noqa: F401
    from triplaneturbo_b200 import ops
    from triplaneturbo_b200.parallel import allreduce_gradients
    from triplaneturbo_b200.synthetic import camera_rays, random_decoder, random_triplanes
    from tests.helpers import build_plugins

    P, V, H, W, R, C, ns, nimp = (wl[k] for k in ("P", "V", "H", "W", "R", "C", "ns", "nimp"))
    B, S = P * V, ns + nimp + 1
    n_rays = B * H * W
    # host-side inputs (pinned): per-rank seeds like the reference's `seed + rank` (launch.py:168)
    sc_h = random_triplanes(P, C, R, seed=100 + rank).pin_memory()
    rays_o_h, rays_d_h, c2w_h, dist_h = [t.pin_memory() for t in camera_rays(B, H, W, seed=200 + rank, views_per_prompt=V)]
    wts = random_decoder(C, seed=1)
    fx = {"space_cache": sc_h[:1].to(dev), **{k: v.to(dev) for k, v in wts.items()}}
    geom, rend = build_plugins(fx, dev, ns, nimp)
    rend.cfg.return_samples = False
    rend.train()
    params = geom.decoder_weights()
    for p_ in params:
        p_.requires_grad_(True)
    g = torch.Generator().manual_seed(3)
    cots = {k: torch.randn(B, H, W, d, generator=g).to(dev) for k, d in
            (("comp_rgb", 3), ("comp_normal_cam_vis", 3), ("disparity", 1))}
    bg = torch.ones(3, device=dev)
    text_embed = torch.zeros(P, 4, device=dev)
    sc_d = sc_h.to(dev).requires_grad_(True)
    rays_o, rays_d, c2w, cam_d = rays_o_h.to(dev), rays_d_h.to(dev), c2w_h.to(dev), dist_h.to(dev)

    def step(sc, ro, rd, c2w_, cd):
        out = rend(ro, rd, None, bg, space_cache=sc, text_embed=text_embed, camera_distances=cd, c2w=c2w_)
        loss = sum((out[k] * cots[k]).sum() for k in cots)
        loss = loss + LAMBDA_EIK * out["eikonal_sum"].sum() / (n_rays * S)
        loss = loss + LAMBDA_SPARSITY * torch.sqrt(out["opacity"] ** 2 + 0.01).mean()
        grads = torch.autograd.grad(loss, [sc] + params)
        if world > 1:   # the one exchange of the path: decoder-weight gradients (92 KB) in one flat buffer, SURVEY 8(e)
            grads = (grads[0], *allreduce_gradients(grads[1:], average=True))
        return out, loss, grads

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ---------------------------------------------------------------------
    for _ in range(args.warmup):
        step(sc_d, rays_o, rays_d, c2w, cam_d)
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    launches0 = ops.launch_count()
    ops.profile_begin()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    t_wall0 = time.time()
    barrier()
    ev[0].record()
    for _ in range(args.steps):
        step(sc_d, rays_o, rays_d, c2w, cam_d)
    ev[1].record()
    barrier()
    t_wall1 = time.time()
    ms_total = ev[0].elapsed_time(ev[1])
    launches = ops.launch_count() - launches0
    prof = ops.profile_end()
    clk = clocks.stop(t_wall0, t_wall1) if clocks else None

    # ---- end-to-end through the plugin API with host buffers ---------------------------------------------
    def e2e_step():
        sc = sc_h.to(dev, non_blocking=True).requires_grad_(True)
        ro, rd = rays_o_h.to(dev, non_blocking=True), rays_d_h.to(dev, non_blocking=True)
        c2w_, cd = c2w_h.to(dev, non_blocking=True), dist_h.to(dev, non_blocking=True)
        out, loss, grads = step(sc, ro, rd, c2w_, cd)
        img = out["comp_rgb"].detach().to("cpu", non_blocking=True)
        return img, loss.detach().to("cpu", non_blocking=True)
    h2d = sum(t.numel() * 4 for t in (sc_h, rays_o_h, rays_d_h, c2w_h, dist_h))
    d2h = n_rays * 3 * 4 + 4
    ms_e2e = float("nan")
    if not args.profile:
        e2e_step(); barrier()
        ev2 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev2[0].record()
        for _ in range(args.steps):
            e2e_step()
        ev2[1].record()
        barrier()
        ms_e2e = ev2[0].elapsed_time(ev2[1])

    times = torch.tensor([ms_total, ms_e2e], device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = times.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_rays = n_rays * world * args.steps
    value = total_rays / (ms_total * 1e-3)
    e2e_val = total_rays / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel -------------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PF sustained"
    fl = mlp_flops(C, S, nimp)
    by = {}
    for name, ms in prof:
        by.setdefault(name, []).append(ms)
    kern_ms = {k: sum(v) / len(v) for k, v in by.items()}
    kern_tot = {k: sum(v) for k, v in by.items()}
    dom = max(kern_tot, key=kern_tot.get)
    f_sdf, f_feat = fl["f_sdf"], fl["f_feat"]
    # algorithmic dense FLOPs per ray of each kernel (per launch averages; k_geo_tc runs twice per step: proposal
    # pass n_imp*F_sdf and fine pass S*2*F_sdf)
    flops_of = {"k_importance_sample": fl["sample"], "k_render_fwd": fl["fwd"], "k_bwd_geo": fl["bwd_geo"],
                "k_bwd_tex": fl["bwd_tex"], "k_geo_tc": (nimp * f_sdf + S * 2 * f_sdf) / 2.0, "k_tex_tc": S * f_feat,
                "k_bwd_geo_tc": S * 4 * f_sdf, "k_bwd_tex_tc": S * 3 * f_feat}
    dom_flops = flops_of.get(dom, 0) * n_rays
    achieved = dom_flops / (kern_ms[dom] * 1e-3) / 1e12
    planes_bytes = P * 6 * C * R * R * 4
    traffic = None      # dram__bytes_read + dram__bytes_write per launch of the dominant kernel, from the committed
    try:                # ncu --set full capture of this workload (profiles/)
        traffic = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json"))).get(args.workload, {}).get(dom)
    except Exception:
        pass
    roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved / peak_tf, "traffic": traffic, "peak_source": peak_src,
                "kernel_ms": kern_ms[dom], "kernel_share_of_step": kern_tot[dom] / ms_total,
                "algorithmic_flops_per_launch": dom_flops,
                "step_tflops": fl["step_survey"] * n_rays * args.steps / (ms_total * 1e-3) / 1e12,
                "hbm_compulsory_gbs": (planes_bytes * 2 + n_rays * 80) * args.steps / (ms_total * 1e-3) / 1e9,
                # measured DRAM bytes of the dominant kernel / its live duration vs the measured HBM peak: the path is
                # not HBM-bound (it is bound by the L1/L2 gather and reduction path)
                # 3xTF32 (forward kernels) runs 3 TF32 MMAs per product and TF32 peaks at half the bf16 rate: the tensor
                # pipe can deliver at most 1/6 of `peak` as algorithmic FLOPs (1/2 for the single-pass backward kernels)
                "split_precision_ceiling": (1.0 / 6.0) if dom in ("k_geo_tc", "k_tex_tc") else 0.5,
                "hbm_gbs_achieved": (traffic / (kern_ms[dom] * 1e-3) / 1e9) if traffic else None,
                "hbm_peak_gbs": peaks.get("hbm_gbs"),
                "kernels_ms": {k: round(v, 4) for k, v in kern_ms.items()},
                "kernels_ms_per_step": {k: round(v / args.steps, 4) for k, v in kern_tot.items()}}

    cpu = None
    if not args.no_cpu_baseline and not args.profile:
        hw = 16
        nr, ts = time_oracle(wl, hw, 2, 1)
        cpu = {"value": nr * len(ts) / sum(ts), "unit": "rays/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"2 steps of 1 prompt x 1 view x {hw}x{hw} rays at {args.workload} plane/sample sizes "
                         f"(oracle/, torch CPU, {os.cpu_count()} threads)"}

    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps({
        "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, **wl, "rays_per_gpu_per_step": n_rays, "samples_per_ray": S,
                   "l2_policy": "inputs larger than L2 (planes %.0f MB + per-sample state > 126 MB)" % (planes_bytes / 1e6),
                   "parallelism": f"dp{world} over prompts"},
        "e2e": {"value": e2e_val, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "clocks": clk, "roofline": roofline, "cpu_baseline": cpu}))
    sys.stdout.flush()
    os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
