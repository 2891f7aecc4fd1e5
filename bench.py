#!/usr/bin/env python
"""Benchmark of the triplane volume-rendering hot path: rendered rays/s, forward + backward.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Render workloads (a step = importance sampler -> fused march (forward) -> loss on the rendered images + eikonal +
sparsity -> backward to the triplanes and decoder weights -> all-reduce of the decoder-weight gradients when N > 1;
prompts shard across ranks with no data-path collective, weak scaling):
    config3 (default)  the configuration BASELINE.json's metric is quoted on: 4 prompts x 4 views x 512^2 rays per GPU,
                       256^2 x 32-channel triplanes, 64 + 128 samples per ray (configs[2] per-GPU share)
    config2            configs[1]: 4 prompts, 256^2 x 40ch, 4 views 256^2, 96 + 192 samples (also printed as `secondary`)
    config1            configs[0]: 64^2 x 32ch, 1 camera 128^2
    config4            configs[3] render part: 4 parts x PatchRenderer (global 42^2 + patch 40^2, P=2, V=4, drop-in outputs)
                       + the mesh path's 128^3 field query with gradients; the SD generator is replaced by synthetic triplanes
    config3q, config2q a quarter of the rays of config3 / config2 (profiling only: 4x shorter ncu replays)
Operator workloads (a step = one call; metric named in the line): sampler, compositor, config5 (512^3 field query).

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how `roofline`, `dropin` and `cpu_baseline` are defined.
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
    "config3": dict(P=4, V=4, H=512, W=512, R=256, C=32, ns=64, nimp=128),
    "config2": dict(P=4, V=4, H=256, W=256, R=256, C=40, ns=96, nimp=192),
    "config1": dict(P=1, V=1, H=128, W=128, R=64, C=32, ns=64, nimp=128),
    "config3q": dict(P=4, V=4, H=256, W=256, R=256, C=32, ns=64, nimp=128),
    "config2q": dict(P=4, V=4, H=128, W=128, R=256, C=40, ns=96, nimp=192),
    # PRD step: the renderer is entered through the PatchRenderer with 168^2 rays per view: 42^2 global + 40^2 patch
    "config4": dict(P=2, V=4, H=168, W=168, R=256, C=32, ns=64, nimp=128, parts=4, patch=40, downsample=4, grid=128),
}
OP_WORKLOADS = ("sampler", "compositor", "config5")
try:
    METRIC = json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]
except Exception:
    METRIC = "rendered rays/sec (fwd+bwd) at 256³ triplane, 512² views, 1/2/4/8 B200"
# what the path computes in: fp32 everywhere except the decoder layers, which run on tcgen05 as 3xTF32 (split fp32,
# ~1e-6 relative) in the forward and in the SDF backward, single-pass TF32 in the colour backward
DTYPE = "f32 (decoder layers on tcgen05: forward 3xTF32 = fp32-equivalent, backward layers 1xTF32; gathers, compositing, reductions fp32)"
LAMBDA_EIK, LAMBDA_SPARSITY = 0.1, 0.5


def mlp_flops(C, S, nimp):
    """Dense decoder FLOPs per ray (2 x MAC), SURVEY.md 8(d)."""
    f_sdf = 2 * (64 * C + 64 * 64 + 64)
    f_feat = 2 * (192 * C + 64 * 64 + 192)
    return dict(f_sdf=f_sdf, f_feat=f_feat,
                sample=nimp * f_sdf, fwd=S * (2 * f_sdf + f_feat), bwd_geo=S * 4 * f_sdf, bwd_tex=S * 3 * f_feat,
                step_survey=nimp * f_sdf + S * (2 * f_sdf + f_feat) + S * (2 * f_feat + 4 * f_sdf))


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


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
def oracle_step_fns(wl, sample_hw):
    """fwd-only and fwd+bwd of the CPU oracle on a bounded sample of the workload: 1 prompt, 1 view, sample_hw^2 rays."""
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

    def fwd():
        return rp.render_forward(rays_o, rays_d, sc, w, pc, torch.ones(3), dist, c2w)

    def step():
        out = fwd()
        loss = sum((out[k] * cots[k]).sum() for k in cots)
        loss = loss + LAMBDA_EIK * ((out["sdf_grad"].norm(dim=-1) - 1.0) ** 2).mean()
        loss = loss + LAMBDA_SPARSITY * torch.sqrt(out["opacity"] ** 2 + 0.01).mean()
        torch.autograd.grad(loss, [sc] + w["sdf"] + w["feature"])
        return float(loss.detach())
    return fwd, step, sample_hw * sample_hw


def time_fn(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(steps):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    return ts


def cpu_baseline(wl, name, sample_hw=32, steps=5, warmup=3):
    """The oracle (reference port) on the host cores: >= 3 warm-ups, `steps` timed, forward-only and forward+backward."""
    torch.set_num_threads(os.cpu_count())
    fwd, step, n_rays = oracle_step_fns(wl, sample_hw)
    ts = time_fn(step, steps, warmup)
    with torch.no_grad():
        tf = time_fn(fwd, steps, 1)
    S = wl["ns"] + wl["nimp"] + 1
    sample = (f"{warmup} warm-up + {steps} steps of 1 prompt x 1 view x {sample_hw}x{sample_hw} rays at {name} plane/sample sizes "
              f"(R={wl['R']}, C={wl['C']}, S={S}); oracle/ = the reference's algorithm in torch on the CPU, {os.cpu_count()} threads")
    return {"value": n_rays * len(ts) / sum(ts), "unit": "rays/s", "cores": os.cpu_count(), "kind": "port", "sample": sample,
            "median_rays_per_s": n_rays / statistics.median(ts), "fwd_only_rays_per_s": n_rays * len(tf) / sum(tf),
            "cpu_model": cpu_model(), "steps": steps, "warmup": warmup}, ts


def reference_arm(args, wl, rank):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (the reference itself
    cannot be imported: nerfacc/tinycudann/lightning are absent, nerfacc has no CPU kernels), all host threads."""
    if rank != 0:
        return
    hw = 32
    warm = max(min(args.warmup, 3), 1)
    cb, ts = cpu_baseline(wl, args.workload, hw, max(args.steps, 1), warm)
    total = sum(ts)
    val = cb["value"]
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "rays/s", "n_gpus": args.gpus, "steps": len(ts),
        "warmup": warm, "ms_per_step": 1e3 * total / len(ts), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, **wl, "sample": cb["sample"]},
        "cpu_baseline": cb,
        "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------------ render workloads
class RenderBench:
    """Inputs, plugins and the step function of one render workload on this rank's GPU."""

    def __init__(self, name, wl, dev, rank, world, dropin=False):
        from triplaneturbo_b200.synthetic import build_plugins, camera_rays, random_decoder, random_triplanes
        self.name, self.wl, self.dev, self.world, self.dropin = name, wl, dev, world, dropin
        P, V, H, W, R, C, ns, nimp = (wl[k] for k in ("P", "V", "H", "W", "R", "C", "ns", "nimp"))
        self.B, self.S, self.n_rays = P * V, ns + nimp + 1, P * V * H * W
        # host-side inputs (pinned): per-rank seeds like the reference's `seed + rank` (launch.py:168)
        self.sc_h = random_triplanes(P, C, R, seed=100 + rank).pin_memory()
        self.rays_h = [t.pin_memory() for t in camera_rays(self.B, H, W, seed=200 + rank, views_per_prompt=V)]
        wts = random_decoder(C, seed=1)
        fx = {"space_cache": self.sc_h[:1].to(dev), **{k: v.to(dev) for k, v in wts.items()}}
        self.geom, self.rend = build_plugins(fx, dev, ns, nimp, rgb_grad_shrink=[0, 1, 0.01, 20000])
        self.rend.cfg.return_samples = bool(dropin)
        self.rend.update_step(0, 10000)              # mid-training value of the shipped rgb_grad_shrink schedule (yaml:139)
        self.rend.train()
        self.params = self.geom.decoder_weights()
        for p_ in self.params:
            p_.requires_grad_(True)
        g = torch.Generator().manual_seed(3)
        self.cots = {k: torch.randn(self.B, H, W, d, generator=g).to(dev) for k, d in
                     (("comp_rgb", 3), ("comp_normal_cam_vis", 3), ("disparity", 1))}
        self.bg = torch.ones(3, device=dev)
        self.text_embed = torch.zeros(P, 4, device=dev)
        self.sc_d = self.sc_h.to(dev).requires_grad_(True)
        self.rays_d = [t.to(dev) for t in self.rays_h]

    def step(self, sc, ro, rd, c2w_, cd):
        from triplaneturbo_b200.parallel import allreduce_gradients
        out = self.rend(ro, rd, None, self.bg, space_cache=sc, text_embed=self.text_embed, camera_distances=cd, c2w=c2w_)
        loss = sum((out[k] * self.cots[k]).sum() for k in self.cots)
        if self.dropin:   # the reference system's own eikonal term on the per-sample gradient (…generator.py:690-714)
            loss = loss + LAMBDA_EIK * ((torch.linalg.norm(out["sdf_grad"], ord=2, dim=-1) - 1.0) ** 2).mean()
        else:             # same loss from the per-ray sums the kernels accumulate (no per-sample tensors leave the library)
            loss = loss + LAMBDA_EIK * out["eikonal_sum"].sum() / (self.n_rays * self.S)
        loss = loss + LAMBDA_SPARSITY * torch.sqrt(out["opacity"] ** 2 + 0.01).mean()
        grads = torch.autograd.grad(loss, [sc] + self.params)
        if self.world > 1:   # the one exchange of the path: decoder-weight gradients (92 KB) in one flat buffer, SURVEY 8(e)
            grads = (grads[0], *allreduce_gradients(grads[1:], average=True))
        return out, loss, grads

    def step_resident(self):
        return self.step(self.sc_d, *self.rays_d)

    def _upload(self):
        """H2D of one step's inputs (pinned host buffers) on the copy stream; returns the device tensors + a ready event."""
        dev = self.dev
        with torch.cuda.stream(self.copy_stream):
            sc = self.sc_h.to(dev, non_blocking=True)
            rays = [t.to(dev, non_blocking=True) for t in self.rays_h]
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        return sc, rays, ev

    def step_e2e(self):
        """One step with host buffers: this step's inputs were uploaded on the copy stream while the previous step computed
        (every step still uploads its own inputs and reads its own result back inside the timed region)."""
        if not hasattr(self, "copy_stream"):
            self.copy_stream = torch.cuda.Stream(self.dev)
            self._next = self._upload()
        sc, rays, ev = self._next
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(ev)
        for t in [sc] + rays:
            t.record_stream(cur)
        self._next = self._upload()                       # overlaps with this step's kernels
        sc = sc.requires_grad_(True)
        out, loss, grads = self.step(sc, *rays)
        done = torch.cuda.Event()
        done.record(cur)
        with torch.cuda.stream(self.copy_stream):         # D2H of the result off the compute stream
            self.copy_stream.wait_event(done)
            img = out["comp_rgb"].detach()
            img.record_stream(self.copy_stream)
            img_h = img.to("cpu", non_blocking=True)
            loss_h = loss.detach().to("cpu", non_blocking=True)
        return img_h, loss_h

    def h2d_bytes(self):
        return sum(t.numel() * 4 for t in [self.sc_h] + self.rays_h)

    def d2h_bytes(self):
        return self.n_rays * 3 * 4 + 4


def run_render(args, wl, rank, world, local, real_stdout):
    import torch.distributed as dist
    dev = torch.device("cuda", local)
    from triplaneturbo_b200 import ops

    if args.workload == "config4":
        rb = make_prd_bench(args.workload, wl, dev, rank, world)
    else:
        rb = RenderBench(args.workload, wl, dev, rank, world, dropin=False)
    P, V, H, W, R, C, ns, nimp = (wl[k] for k in ("P", "V", "H", "W", "R", "C", "ns", "nimp"))
    S, n_rays = rb.S, rb.n_rays

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(steps):
            fn()
        ev[1].record()
        barrier()
        return ev[0].elapsed_time(ev[1])

    # ---- device-resident timing ---------------------------------------------------------------------
    for _ in range(args.warmup):
        rb.step_resident()
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    launches0 = ops.launch_count()
    ops.profile_begin()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    t_wall0 = time.time()
    barrier()
    ev[0].record()
    for _ in range(args.steps):
        rb.step_resident()
    ev[1].record()
    barrier()
    t_wall1 = time.time()
    ms_total = ev[0].elapsed_time(ev[1])
    launches = ops.launch_count() - launches0
    prof = ops.profile_end()
    clk = clocks.stop(t_wall0, t_wall1) if clocks else None

    # ---- end-to-end through the plugin API with host buffers ---------------------------------------------
    ms_e2e = float("nan")
    if not args.profile:
        ms_e2e = timed(rb.step_e2e, args.steps, 1)

    # ---- drop-in figure: the same step with the reference system's per-sample outputs (return_samples=True) and the
    # eikonal loss taken from out["sdf_grad"] as …generator.py:690-714 does -------------------------------------
    ms_drop, drop_steps = float("nan"), min(args.steps, 3)
    if not args.profile and not args.no_dropin and args.workload != "config4":
        rb.dropin = True
        rb.rend.cfg.return_samples = True
        ms_drop = timed(rb.step_resident, drop_steps, 1)
        rb.dropin = False
        rb.rend.cfg.return_samples = False
    # ---- eval-mode forward (no gradients, no per-sample outputs): sampler + march only -------------------------------
    ms_eval, eval_steps = float("nan"), min(args.steps, 3)
    if not args.profile and args.workload != "config4":
        rb.rend.eval()

        def eval_fwd():       # the reference's eval path takes ONE space cache and its views per call (REN:150-185)
            V_ = rb.wl["V"]
            outs = []
            with torch.no_grad():
                for p_ in range(rb.wl["P"]):
                    sl = slice(p_ * V_, (p_ + 1) * V_)
                    outs.append(rb.rend(rb.rays_d[0][sl], rb.rays_d[1][sl], None, rb.bg, space_cache=rb.sc_d.detach()[p_:p_ + 1],
                                        text_embed=rb.text_embed[p_:p_ + 1], camera_distances=rb.rays_d[3][sl],
                                        c2w=rb.rays_d[2][sl])["comp_rgb"])
            return outs
        ms_eval = timed(eval_fwd, eval_steps, 1)
        rb.rend.train()
    h2d, d2h = rb.h2d_bytes(), rb.d2h_bytes()
    del rb
    ops.clear_caches()
    torch.cuda.empty_cache()

    # ---- secondary line: BASELINE configs[1] (device-resident, short) -------------------------------------------
    ms_sec, sec_steps, sec_rays = float("nan"), 3, 0
    if args.workload == "config3" and not args.profile and not args.no_secondary:
        rb2 = RenderBench("config2", dict(WORKLOADS["config2"]), dev, rank, world, dropin=False)
        ms_sec = timed(rb2.step_resident, sec_steps, 3)
        sec_rays = rb2.n_rays
        del rb2
        ops.clear_caches()
        torch.cuda.empty_cache()

    times = torch.tensor([ms_total, ms_e2e, ms_drop, ms_sec, ms_eval], device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, ms_drop, ms_sec, ms_eval = times.tolist()
    if rank != 0:
        return

    total_rays = n_rays * world * args.steps
    value = total_rays / (ms_total * 1e-3)
    e2e_val = total_rays / (ms_e2e * 1e-3) if ms_e2e == ms_e2e else None

    # ---- roofline of the dominant kernel -------------------------------------------------------------------
    peaks = load_peaks()
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PF sustained"
    fl = mlp_flops(C, S, nimp)
    by = {}
    for name, ms in prof:
        by.setdefault(name, []).append(ms)
    kern_ms = {k: sum(v) / len(v) for k, v in by.items()}
    kern_tot = {k: sum(v) for k, v in by.items()}
    f_sdf, f_feat = fl["f_sdf"], fl["f_feat"]
    # algorithmic dense FLOPs per ray of each decoder kernel, per launch (the geometry forward kernel runs twice
    # per step: proposal pass n_imp*F_sdf and fine pass S*2*F_sdf, so its per-launch average is half their sum)
    geo_fwd = (nimp * f_sdf + S * 2 * f_sdf) / 2.0
    flops_of = {"k_importance_sample": fl["sample"], "k_render_fwd": fl["fwd"], "k_bwd_geo": fl["bwd_geo"],
                "k_bwd_tex": fl["bwd_tex"], "k_geo_tc": geo_fwd, "k_geo_ws": geo_fwd, "k_tex_tc": S * f_feat,
                "k_tex_ws": S * f_feat, "k_tex_tc1": S * f_feat, "k_bwd_geo_tc": S * 4 * f_sdf, "k_bwd_geo_ws": S * 4 * f_sdf,
                "k_bwd_tex_tc": S * 3 * f_feat, "k_bwd_tex_ws": S * 3 * f_feat}
    cand = {k: v for k, v in kern_tot.items() if k in flops_of} or kern_tot
    dom = max(cand, key=cand.get)
    dom_flops = flops_of.get(dom, 0) * n_rays
    achieved = dom_flops / (kern_ms[dom] * 1e-3) / 1e12
    planes_bytes = P * 6 * C * R * R * 4
    traffic = None      # dram__bytes_read + dram__bytes_write per launch of the dominant kernel, from the committed
    try:                # ncu --set full capture of this workload (profiles/dram_traffic.json)
        traffic = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json"))).get(args.workload, {}).get(dom)
    except Exception:
        pass
    fwd3 = dom in ("k_geo_tc", "k_tex_tc", "k_tex_tc1", "k_geo_ws", "k_tex_ws", "k_bwd_geo_ws")
    roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved / peak_tf, "traffic": traffic, "peak_source": peak_src,
                "kernel_ms": kern_ms[dom], "kernel_share_of_step": kern_tot[dom] / ms_total,
                "algorithmic_flops_per_launch": dom_flops,
                "step_tflops": fl["step_survey"] * n_rays * args.steps / (ms_total * 1e-3) / 1e12,
                "step_frac": fl["step_survey"] * n_rays * args.steps / (ms_total * 1e-3) / 1e12 / peak_tf,
                "hbm_compulsory_gbs": (planes_bytes * 2 + n_rays * 80) * args.steps / (ms_total * 1e-3) / 1e9,
                # 3xTF32 runs 3 TF32 MMAs per product and TF32 peaks at half the bf16 rate: the tensor pipe can deliver at
                # most 1/6 of `peak` as algorithmic FLOPs for those kernels (1/2 for single-pass TF32 kernels)
                "split_precision_ceiling": (1.0 / 6.0) if fwd3 else 0.5,
                "hbm_gbs_achieved": (traffic / (kern_ms[dom] * 1e-3) / 1e9) if traffic else None,
                "hbm_peak_gbs": peaks.get("hbm_gbs"),
                "kernels_ms": {k: round(v, 4) for k, v in kern_ms.items()},
                "kernels_ms_per_step": {k: round(v / args.steps, 4) for k, v in kern_tot.items()}}

    cpu = None
    if not args.no_cpu_baseline and not args.profile:
        cpu, _ = cpu_baseline(wl, args.workload, 32, 5, 3)

    line = {
        "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": {"workload": args.workload, **wl, "rays_per_gpu_per_step": n_rays, "samples_per_ray": S,
                   "l2_policy": "inputs larger than L2 (planes %.0f MB + per-sample state > 126 MB)" % (planes_bytes / 1e6),
                   "parallelism": f"dp{world} over prompts", "rgb_grad_shrink": "C([0,1,0.01,20000]) at step 10000"},
        "e2e": {"value": e2e_val, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "clocks": clk, "roofline": roofline, "cpu_baseline": cpu}
    if ms_drop == ms_drop:
        line["dropin"] = {"value": n_rays * world * drop_steps / (ms_drop * 1e-3), "unit": "rays/s",
                          "ms_per_step": ms_drop / drop_steps, "steps": drop_steps,
                          "what": "same step with return_samples=True (every per-sample output of the reference renderer "
                                  "materialised, colour decoder at every sample) and the eikonal loss from out['sdf_grad']"}
    if ms_eval == ms_eval:
        line["eval_forward"] = {"value": n_rays * world * eval_steps / (ms_eval * 1e-3), "unit": "rays/s",
                                "ms_per_step": ms_eval / eval_steps, "steps": eval_steps,
                                "what": "no-grad eval render (importance sampler + march, no per-sample outputs, no masks)"}
    if ms_sec == ms_sec:
        line["secondary"] = {"workload": "config2", **WORKLOADS["config2"], "value": sec_rays * world * sec_steps / (ms_sec * 1e-3),
                             "unit": "rays/s", "ms_per_step": ms_sec / sec_steps, "steps": sec_steps, "warmup": 3}
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(line))
    sys.stdout.flush()
    os.dup2(2, 1)


def make_prd_bench(name, wl, dev, rank, world):
    """config4: see PrdBench's docstring."""
    import triplaneturbo_b200 as tt  # noqa: F401
    from triplaneturbo_b200.renderer import PatchRenderer

    rb = RenderBench(name, wl, dev, rank, world, dropin=True)
    pr = PatchRenderer.__new__(PatchRenderer)
    torch.nn.Module.__init__(pr)
    pr.cfg = PatchRenderer.Config(patch_size=wl["patch"], global_downsample=wl["downsample"])
    pr.base_renderer = rb.rend
    ds, psz, parts, grid = wl["downsample"], wl["patch"], wl["parts"], wl["grid"]
    H, W = wl["H"], wl["W"]
    rb.n_rays = rb.B * ((H // ds) * (W // ds) + psz * psz) * parts
    dparams = rb.geom.deformation_network.weights()
    for p_ in dparams:
        p_.requires_grad_(True)
    g = torch.Generator().manual_seed(5)
    lin = torch.linspace(-1.0, 1.0, grid)
    pts = torch.stack(torch.meshgrid(lin, lin, lin, indexing="ij"), -1).reshape(1, -1, 3).repeat(wl["P"], 1, 1).to(dev)
    cot_sdf = (torch.randn(wl["P"], grid ** 3, 1, generator=g) * 1e-3).to(dev)
    cot_def = (torch.randn(wl["P"], grid ** 3, 3, generator=g) * 1e-3).to(dev)
    # upsampled cotangents: the PatchRenderer returns H x W images (global upsampled + patch pasted)
    g2 = torch.Generator().manual_seed(3)
    rb.cots = {k: torch.randn(rb.B, H, W, d, generator=g2).to(dev) for k, d in
               (("comp_rgb", 3), ("comp_normal_cam_vis", 3), ("disparity", 1))}

    def step(sc, ro, rd, c2w_, cd):
        from triplaneturbo_b200.parallel import allreduce_gradients
        total = None
        for _ in range(parts):
            out = pr(ro, rd, None, rb.bg, space_cache=sc, text_embed=rb.text_embed, camera_distances=cd, c2w=c2w_)
            loss = sum((out[k] * rb.cots[k]).sum() for k in rb.cots)
            loss = loss + LAMBDA_SPARSITY * torch.sqrt(out["opacity"] ** 2 + 0.01).mean()
            sdf, deform = rb.geom.forward_field(pts, sc)            # mesh path: isosurface grid query with gradients
            loss = loss + (sdf * cot_sdf).sum() + (deform * cot_def).sum()
            grads = torch.autograd.grad(loss, [sc] + rb.params + dparams)
            total = grads if total is None else [a + b for a, b in zip(total, grads)]
        if world > 1:
            total = (total[0], *allreduce_gradients(total[1:], average=True))
        return out, loss, total
    rb.step = step
    return rb


# ------------------------------------------------------------------------------------------------ operator workloads
def run_op(args, real_stdout):
    """Stand-alone operators of the path against the roof that binds them."""
    from triplaneturbo_b200 import ops
    from triplaneturbo_b200.sampler import PLANES, project_onto_planes, sample_planes
    from triplaneturbo_b200.synthetic import build_plugins, camera_rays, random_decoder, random_triplanes
    dev = torch.device("cuda", 0)
    peaks = load_peaks()
    hbm = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6.65 TB/s"

    def timed(fn):
        for _ in range(max(args.warmup, 3)):
            fn()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        launches0 = ops.launch_count()
        ev[0].record()
        for _ in range(args.steps):
            fn()
        ev[1].record()
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[1]) / args.steps, (ops.launch_count() - launches0)

    extra = {}
    if args.workload == "sampler":
        # sample_from_planes, shipped texture mode (v2: three planes concatenated), samples along camera rays (ray-major)
        N, K, R, C, S = 4, 3, 256, 32, 64
        gen = torch.Generator(dev).manual_seed(0)
        planes = torch.randn(N * K, R, R, C, device=dev, generator=gen)
        ro, rd, _, _ = camera_rays(N, 256, 256, seed=2)
        t = torch.linspace(0.6, 2.6, S)
        pts = (ro.reshape(N, -1, 1, 3) + rd.reshape(N, -1, 1, 3) * t.view(1, 1, S, 1)).reshape(N, -1, 3).to(dev)
        grid = project_onto_planes(PLANES, pts).contiguous()
        M = grid.shape[1]
        units, unit, metric = N * M, "points/s", "sample_from_planes (v2, 3 planes concatenated) forward, points/s"
        OS = K * C
        alg = N * M * (8 * K + 4 * OS) + planes.numel() * 4
        ms, launches = timed(lambda: sample_planes(planes, grid, K, True))
        pl, gr = planes.clone().requires_grad_(True), grid.clone().requires_grad_(True)
        go = torch.randn(N, M, OS, device=dev, generator=gen)
        ms_fb, _ = timed(lambda: torch.autograd.grad(sample_planes(pl, gr, K, True), [pl, gr], go))
        alg_b = N * M * (8 * K + 4 * OS) + 8 * K * N * M + 2 * planes.numel() * 4
        extra = {"backward": {"ms": ms_fb - ms, "achieved_gbs": alg_b / ((ms_fb - ms) * 1e-3) / 1e9,
                              "frac": alg_b / ((ms_fb - ms) * 1e-3) / 1e9 / hbm, "algorithmic_bytes": alg_b}}
        bound, peak, peak_src, ach = "hbm", hbm, hbm_src, alg / (ms * 1e-3) / 1e9
        cfg = {"workload": "sampler", "N": N, "K": K, "R": R, "C": C, "M": M, "mode": "v2 concat", "points": "camera rays, ray-major"}
    elif args.workload == "compositor":
        n, S, D = 1 << 20, 193, 3
        gen = torch.Generator(dev).manual_seed(0)
        alphas = torch.rand(n, S, device=dev, generator=gen) * 0.1
        values = torch.randn(n, S, D, device=dev, generator=gen)
        units, unit, metric = n, "rays/s", "render_weight_from_alpha + accumulate_along_rays (dense rays, D=3), rays/s"
        alg = n * S * (4 + 4 * D) + n * S * 8 + n * 4 * D           # alphas + values read, weights + trans written, out
        ms, launches = timed(lambda: ops.composite_fwd(alphas, values))
        bound, peak, peak_src, ach = "hbm", hbm, hbm_src, alg / (ms * 1e-3) / 1e9
        cfg = {"workload": "compositor", "n_rays": n, "S": S, "D": D}
    else:
        res, C, R = 512, 32, 256
        sc = random_triplanes(1, C, R, seed=0).to(dev)
        fx = {"space_cache": sc, **{k: v.to(dev) for k, v in random_decoder(C, seed=1).items()}}
        geom, _ = build_plugins(fx, dev, 64, 128)
        units, unit = res ** 3, "points/s"
        metric = "mesh-export field query: SDF + deformation at the 512^3 isosurface grid (BASELINE configs[4]), points/s"
        f = 2 * (64 * C + 64 * 64 + 64) + 2 * (64 * C + 64 * 64 + 192)
        alg = units * f
        with torch.no_grad():
            ms, launches = timed(lambda: geom.forward_field_grid(res, sc))
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PF sustained"
        bound, ach = "tensor", alg / (ms * 1e-3) / 1e12
        cfg = {"workload": "config5", "grid": res, "R": R, "C": C, "P": 1}
    line = {"metric": metric, "value": units / (ms * 1e-3), "unit": unit, "n_gpus": 1, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if bound == "hbm" else DTYPE, "data": "synthetic", "config": cfg,
            "e2e": None, "gpu_launches": launches,
            "roofline": {"bound": bound, "achieved": ach, "peak": peak, "unit": "GB/s" if bound == "hbm" else "TFLOP/s",
                         "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                         ("algorithmic_bytes_per_launch" if bound == "hbm" else "algorithmic_flops_per_launch"): alg, **extra},
            "cpu_baseline": None}
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    print(json.dumps(line))
    sys.stdout.flush()
    os.dup2(2, 1)


# ------------------------------------------------------------------------------------------------ entry
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--option", action="append", help="library experiment switch name=value (tt_set_option); not for headline numbers")
    ap.add_argument("--workload", default="config3", choices=sorted(WORKLOADS) + list(OP_WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dropin", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--kernels", type=int, default=None, help="kernel family: 2 warp-specialised tcgen05 (default), 1 round-1 tcgen05, 0 SIMT")
    ap.add_argument("--profile", action="store_true", help="for ncu: 1 warm-up + K steps, no e2e / dropin / cpu legs (not a bench value)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        wl = dict(WORKLOADS[args.workload if args.workload in WORKLOADS else "config3"])
        return reference_arm(args, wl, rank)
    args.warmup = 1 if args.profile else max(args.warmup, 3)

    # Libraries (NCCL prints its version banner) may write to fd 1: keep stdout for the ONE JSON line by pointing fd 1
    # at stderr until the result is printed.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import triplaneturbo_b200 as tt  # noqa: F401
    from triplaneturbo_b200 import ops
    if args.kernels is not None:
        ops.set_impl(args.kernels)
    for kv in args.option or []:          # experiment switches of the library (tt_set_option), e.g. --option scatter=2
        k_, v_ = kv.split("=")
        ops.set_option(k_, int(v_))
    if args.workload in OP_WORKLOADS:
        if rank == 0:
            run_op(args, real_stdout)
    else:
        run_render(args, dict(WORKLOADS[args.workload]), rank, world, local, real_stdout)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
