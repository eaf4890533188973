#!/usr/bin/env python
"""Headline benchmark: blurry views / s (forward + backward, F sub-frames) -- BASELINE.json's metric.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2] [--impl ours|reference]

One "step" = one blurry view per GPU: sub-frame poses from the Bezier control points, all F sub-frame
renders, their mean, L1 loss against a ground-truth image, and the complete backward down to the
Gaussian parameters and the control points.  N > 1 (torchrun, one rank per GPU): every rank renders a
different blurry view over the same replicated Gaussians (views of a batch are sharded across GPUs,
weak scaling) and the Gaussian gradients are summed with one NCCL all-reduce inside the timed region.

`--impl reference` times the reference's own implementation on the same GPU: its unmodified CUDA
extension from baseline/_ref driven by the reference's per-sub-frame Python loop (restated in
oracle/pose_torch.py because /root/reference is not present on the GPU box).
"""
import argparse
import ctypes as C
import json
import math
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "blurry_views_per_s_fwd_bwd"
UNIT = "views/s"


# ------------------------------------------------------------------------------------------------
def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its banner / NCCL_DEBUG output to stdout by default: keep stdout for the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self):
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
class RefCamera:
    def __init__(self, cam, device):
        self.image_width, self.image_height = cam.width, cam.height
        self.FoVx, self.FoVy, self.znear, self.zfar = cam.fovx, cam.fovy, cam.znear, cam.zfar
        self.projection_matrix = cam.projection_matrix_t().to(device)


def build_workload(config, rank, device):
    from deblurgs_b200 import synthetic
    from deblurgs_b200.motion import CameraMotionModule, GaussianParams
    P, W, H, F, order = synthetic.get_config(config)
    cam = synthetic.make_camera(W, H)
    scene = synthetic.make_scene(P, cam, seed=0).to(device)            # same Gaussians on every rank
    traj = synthetic.make_trajectory(F, order, seed=1 + rank).to(device)  # one view per rank
    gt_host = synthetic.make_target(cam, seed=2 + rank).pin_memory()
    bg = synthetic.make_background().to(device)
    gaussians = GaussianParams.from_scene(scene)
    base = torch.tensor([synthetic.BASE_SE3], dtype=torch.float32, device=device)
    cmm = CameraMotionModule([RefCamera(cam, device)], base, curve_order=order, num_subframes=F)
    with torch.no_grad():
        cmm._trans._control_points.copy_(traj.ctrl_trans[None])
        cmm._rot._control_points.copy_(traj.ctrl_rot[None])
    cmm.link_gaussian(gaussians)
    return dict(P=P, W=W, H=H, F=F, order=order, cam=cam, scene=scene, traj=traj, gt_host=gt_host, bg=bg,
                gaussians=gaussians, cmm=cmm)


def flat_grad_views(params):
    """One flat gradient buffer; every parameter's .grad is a view into it (single all-reduce)."""
    n = sum(p.numel() for p in params)
    flat = torch.zeros(n, dtype=torch.float32, device=params[0].device)
    off = 0
    for p in params:
        p.grad = flat[off:off + p.numel()].view_as(p)
        off += p.numel()
    return flat


def make_step_ours(w, world, lambda_t_smooth=0.0):
    from deblurgs_b200.loss import blur_photometric_loss
    cmm, g = w["cmm"], w["gaussians"]
    gparams = g.parameters()
    cparams = cmm.parameters()
    flat = flat_grad_views(gparams) if world > 1 else None

    def step(gt, gt_ready=None):
        if flat is not None:
            flat.zero_()
        else:
            for p in gparams:
                p.grad = None
        for p in cparams:
            p.grad = None
        out = cmm.query(0, "all", background=w["bg"])
        if gt_ready is not None:   # ground truth uploaded on a side stream while the view rendered
            torch.cuda.current_stream().wait_event(gt_ready)
        # mean |blurred - gt| (+ lambda * mean |subframes[1:] - subframes[:-1]| for --loss smooth), fused
        loss = blur_photometric_loss(out["blurred"], out["subframes"], gt, lambda_t_smooth)
        loss.backward()
        if flat is not None:
            import torch.distributed as dist
            dist.all_reduce(flat)
        return loss
    return step


def make_step_reference(w, lambda_t_smooth=0.0):
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    from oracle import pose_torch as pt   # the reference's Python loop, restated (test infrastructure)
    cmm, g, cam = w["cmm"], w["gaussians"], w["cam"]
    ref_cam = cmm.original_cam[0]
    params = g.parameters() + cmm.parameters()

    def render(view, proj, center, bg):   # gaussian_renderer/__init__.py:18-90
        xyz = g.get_xyz
        screenspace_points = torch.zeros_like(xyz, dtype=xyz.dtype, requires_grad=True, device="cuda") + 0
        screenspace_points.retain_grad()
        rs = GaussianRasterizationSettings(
            image_height=int(cam.height), image_width=int(cam.width), tanfovx=math.tan(cam.fovx * 0.5),
            tanfovy=math.tan(cam.fovy * 0.5), bg=bg, scale_modifier=1.0, z_near=g.z_near, z_far=g.z_far,
            use_sigmoid=g.use_sigmoid, sh_degree=g.active_sh_degree, campos=center, prefiltered=False, debug=False)
        img, depth, radii = GaussianRasterizer(raster_settings=rs)(
            means3D=xyz, means2D=screenspace_points, shs=g.get_features, colors_precomp=None,
            opacities=g.get_opacity, scales=g.get_scaling, rotations=g.get_rotation, cov3D_precomp=None,
            viewmatrix=view, projmatrix=proj)
        return img

    def step(gt, gt_ready=None):
        for p in params:
            p.grad = None
        nu = pt.sample_nu(cmm._nu[0], cmm.n_subframes)
        poses = pt.trajectory(cmm._trans._control_points[0], cmm._rot._control_points[0], nu,
                              ref_cam.projection_matrix)
        subframes = torch.stack([render(v, p, c, w["bg"]) for (v, p, c) in poses])   # scene/motion.py:141-145
        blurred = subframes.mean(dim=0)
        if gt_ready is not None:
            torch.cuda.current_stream().wait_event(gt_ready)
        loss = (blurred - gt).abs().mean()
        if lambda_t_smooth != 0.0:   # batchwise_smoothness_loss, utils/loss_utils.py:80-93
            loss = loss + lambda_t_smooth * (subframes[1:] - subframes[:-1]).abs().mean()
        loss.backward()
        return loss
    return step


# ------------------------------------------------------------------------------------------------
def workload_stats(w):
    """V, D, E, K, E_b of this rank's view (measurement only, outside the timed region)."""
    from deblurgs_b200 import _lib, rasterizer as rz
    lib = _lib.load()
    g, cmm, cam = w["gaussians"], w["cmm"], w["cam"]
    with torch.no_grad():
        view, proj, campos = cmm.get_trajectory_tensors(0)
        out = rz._forward_batched(g.get_xyz.detach(), g.get_features.detach(), None, g.get_opacity.detach(),
                                  g.get_scaling.detach(), g.get_rotation.detach(), None, view, proj, campos,
                                  w["bg"], cam.height, cam.width, cam.tanfovx, cam.tanfovy, 1.0, g.z_near,
                                  g.z_far, g.active_sh_degree, False, g.use_sigmoid, False, float(w["F"]))
        color, depth, radii, _, D, geom, binning, img = out
        stats = torch.zeros(3, dtype=torch.int64, device=color.device)
        rc = lib.dgs_debug_workload(_lib.ptr(geom), _lib.ptr(binning), _lib.ptr(img), w["P"], w["F"], cam.width,
                                    cam.height, D, _lib.ptr(stats),
                                    C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "dgs_debug_workload")
        E, K, Eb = (int(v) for v in stats.tolist())
        V = int((radii > 0).sum())
    return dict(V=V, D=int(D), E=E, K=K, E_b=Eb)


def algorithmic_bytes(w, st, sort_bits):
    """SURVEY.md 8(d) HBM byte model per blurry view, split per stage."""
    P, F, W, H = w["P"], w["F"], w["W"], w["H"]
    M = w["scene"].shs.shape[1]
    G = 44 + 12 * M
    V, D = st["V"], st["D"]
    passes = math.ceil(sort_bits / 8)
    return {
        "preprocess_fwd": P * G + V * 48,
        "scan": 2 * 4 * P * F,
        "duplicate": V * 16 + D * 12,
        "sort": D * 12 * 2 * passes,
        "tile_ranges": D * 8,
        "render_fwd": D * (4 + 44) + F * H * W * (16 + 8),
        "blur_mean": F * H * W * 12 + H * W * 12,
        "bwd_memset": V * 40,
        "render_bwd": D * (4 + 44) + F * H * W * (16 + 8) + V * 40,
        "preprocess_bwd": P * G + P * G + V * 40 + V * 48,
    }


def read_profile(lib):
    n = lib.dgs_profile_num_stages()
    ms = (C.c_double * n)()
    calls = (C.c_int64 * n)()
    rc = lib.dgs_profile_read(ms, calls, n, 1)
    if rc != 0:
        raise RuntimeError("dgs_profile_read failed")
    return {lib.dgs_profile_stage_name(i).decode(): (ms[i], calls[i]) for i in range(n)}


def cpu_baseline(config_name, seconds_budget=25.0):
    """The numpy oracle (test infrastructure, the checker) timed on one host core on a bounded sample of
    the same workload: whole sub-frames of the benchmark scene, forward + backward, until the budget is
    spent (at least one)."""
    import numpy as np
    from deblurgs_b200 import synthetic
    from oracle import pose_torch as pt, raster_np as rn
    P, W, H, F, order = synthetic.get_config(config_name)
    cam = synthetic.make_camera(W, H)
    scene = synthetic.make_scene(P, cam, seed=0)
    traj = synthetic.make_trajectory(F, order, seed=1)
    bg = synthetic.make_background().numpy()
    poses = pt.trajectory(traj.ctrl_trans, traj.ctrl_rot, traj.nu, cam.projection_matrix_t())
    rng = np.random.default_rng(0)
    dpix = rng.standard_normal((3, H, W)) / (3 * H * W * F)
    ddep = np.zeros((1, H, W))
    a = [t.numpy() for t in (scene.means3D, scene.scales, scene.rotations, scene.opacities, scene.shs)]
    done, t0 = 0, time.perf_counter()
    while done < F and (done == 0 or time.perf_counter() - t0 < seconds_budget):
        v, p, c = (t.detach().numpy() for t in poses[done])
        fw = rn.forward(a[0], a[1], a[2], a[3], a[4], 3, v, p, c, bg, W, H, cam.tanfovx, cam.tanfovy)
        rn.backward(fw, a[0], a[1], a[2], a[4], 3, v, p, c, bg, W, H, cam.tanfovx, cam.tanfovy, dpix, ddep)
        done += 1
    dt = time.perf_counter() - t0
    return {"value": 1.0 / (dt / done * F), "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "%d of %d sub-frames of %s (fwd+bwd, numpy oracle, 1 thread) in %.1f s; "
                      "views/s extrapolated to F=%d; host has %d cores" % (done, F, config_name, dt, F, os.cpu_count())}


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="c2")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--loss", default="l1", choices=["l1", "smooth"],
                    help="l1: mean|blur - gt| (headline); smooth: + 1e-3 * temporal smoothness of the sub-frames "
                         "(SURVEY 8d's second variant: non-uniform per-sub-frame gradients)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=25.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank, world, local = dist_setup(args.gpus)
    if args.impl == "reference" and rank != 0:
        return   # the reference is single-GPU: rank 0 alone runs it
    device = torch.device("cuda", local)
    eff_world = 1 if args.impl == "reference" else world
    w = build_workload(args.config, rank, device)
    F = w["F"]
    lam = 1e-3 if args.loss == "smooth" else 0.0

    if args.impl == "ours":
        from deblurgs_b200 import _lib
        lib = _lib.load()
        step = make_step_ours(w, eff_world, lam)
    else:
        lib = None
        if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "diff_gaussian_rasterization")):
            print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref is not installed"}))
            return
        step = make_step_reference(w, lam)

    gt_dev = w["gt_host"].to(device)
    for _ in range(args.warmup):
        step(gt_dev)
    barrier(eff_world)

    # ---- device-resident timing: inputs already in HBM, CUDA events, max over ranks
    if lib is not None:
        lib.dgs_profile_read(None, None, 0, 1)
        lib.dgs_launch_count(1)
        lib.dgs_profile_enable(1)
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(eff_world)
    e0.record()
    for _ in range(args.steps):
        step(gt_dev)
    e1.record()
    barrier(eff_world)
    sampler.stop_flag = True
    sampler.join()
    ms_total = e0.elapsed_time(e1)
    launches = None
    prof = None
    if lib is not None:
        launches = int(lib.dgs_launch_count(0))
        prof = read_profile(lib)
        lib.dgs_profile_enable(0)
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if eff_world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = eff_world * 1000.0 / ms_step

    # ---- end to end through the public API: pinned-host ground truth in, loss scalar out, every step
    # (the upload runs on a side stream into one of two preallocated device buffers, as a data loader's
    # prefetch would, and is awaited before the loss; the loss read-back makes every step synchronous)
    barrier(eff_world)
    copy_stream = torch.cuda.Stream(device)
    gt_bufs = [torch.empty_like(gt_dev), torch.empty_like(gt_dev)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    t_prev, step_wall = t0, []
    for k in range(args.steps):
        gt = gt_bufs[k & 1]
        with torch.cuda.stream(copy_stream):
            gt.copy_(w["gt_host"], non_blocking=True)
            gt_ready = copy_stream.record_event()
        loss = step(gt, gt_ready)
        loss_host = loss.item()
        t_now = time.perf_counter()
        step_wall.append((t_now - t_prev) * 1e3)
        t_prev = t_now
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
    if eff_world > 1:
        import torch.distributed as dist
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = eff_world * args.steps / float(t_e2e.item())

    if eff_world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    clocks = sampler.summary()
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": eff_world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "ms_per_subframe": ms_step / F,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: %d Gaussians, %dx%d, num_subframes=%d, SH degree 3, se3 Bezier order %d, "
                               "one blurry view per GPU per step (fwd+bwd, %s), inputs larger than L2"
                               % (args.config, w["P"], w["W"], w["H"], F, w["order"],
                                  "L1 loss" if lam == 0.0 else "L1 + 1e-3 temporal-smoothness loss"),
                   "parallelism": "views-dp%d" % eff_world, "loss_last": loss_host},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(w["gt_host"].numel() * 4),
                "d2h_bytes_per_step": 4,
                "ms_per_step_min_median_max": [round(min(step_wall), 3), round(statistics.median(step_wall), 3),
                                               round(max(step_wall), 3)]},
    }
    if args.impl == "reference":
        out["impl"] = "reference"
        out["reference_kind"] = ("reference CUDA extension (baseline/_ref, unmodified, sm_100 build) on the same "
                                 "B200, driven by the reference's per-sub-frame loop")
        out["cpu_baseline"] = {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                               "sample": "full workload; the reference implementation of this path is itself a "
                                         "CUDA extension, so this arm runs on the GPU, host cores only drive it"}
        out["gpu_launches"] = 0
        print(json.dumps(out))
        return

    out["gpu_launches"] = launches
    st = workload_stats(w)
    tb, sb = C.c_int(0), C.c_int(0)
    lib.dgs_key_bits(w["W"], w["H"], F, C.byref(tb), C.byref(sb))
    abytes = algorithmic_bytes(w, st, 32 + tb.value + sb.value)
    stage_ms = {k: (v[0] / max(v[1], 1)) for k, v in prof.items() if v[1] > 0}
    out["stage_ms_per_step"] = {k: round(v[0] / args.steps, 4) for k, v in prof.items() if v[1] > 0}
    out["workload_stats"] = st
    sm_mhz = clocks["sm_mhz"] or 1965.0
    n_sm = torch.cuda.get_device_properties(device).multi_processor_count
    fp32_peak = n_sm * 128 * 2 * sm_mhz * 1e6 / 1e12
    flops = {"render_fwd": 14 * st["E"] + 18 * st["K"], "render_bwd": 16 * st["E_b"] + 88 * st["K"]}
    # dominant kernel among the stages with an algorithmic work figure (under a profiler the tiny latency-bound
    # stages can show the largest event times)
    dom = max((k for k in stage_ms if k in flops or k in abytes), key=stage_ms.get)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(args.config, {}).get(dom)
    if dom in flops:
        ach = flops[dom] / (stage_ms[dom] * 1e-3) / 1e12
        out["roofline"] = {"kernel": dom, "bound": "fp32", "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s",
                           "frac": ach / fp32_peak, "traffic": traffic,
                           "peak_source": "derived: %d SMs x 128 lanes x 2 x %.0f MHz (SM clock sampled during the "
                                          "timed region); MEASURED_PEAKS.json has no fp32 entry" % (n_sm, sm_mhz),
                           "algorithmic_flops_per_launch": flops[dom], "ms_per_launch": stage_ms[dom]}
        if traffic:   # why the bound is not HBM: measured DRAM traffic of the same kernel against the copy bandwidth
            hbm_pk, _ = measured_peaks()
            out["roofline"]["dram_gbs"] = traffic / (stage_ms[dom] * 1e-3) / 1e9
            out["roofline"]["dram_frac_of_hbm_peak"] = out["roofline"]["dram_gbs"] / hbm_pk
    else:
        hbm, src = measured_peaks()
        ach = abytes[dom] / (stage_ms[dom] * 1e-3) / 1e9
        out["roofline"] = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s",
                           "frac": ach / hbm, "traffic": traffic, "peak_source": src,
                           "algorithmic_bytes_per_launch": abytes[dom], "ms_per_launch": stage_ms[dom]}
    hbm, src = measured_peaks()
    hbm_stages = ["preprocess_fwd", "scan", "duplicate", "sort", "tile_ranges", "bwd_memset", "preprocess_bwd"]
    hb = sum(abytes[k] for k in hbm_stages if k in stage_ms)
    ht = sum(stage_ms[k] for k in hbm_stages if k in stage_ms)
    out["roofline_hbm_group"] = {"kernels": hbm_stages, "bound": "hbm", "achieved": hb / (ht * 1e-3) / 1e9,
                                 "peak": hbm, "unit": "GB/s", "frac": hb / (ht * 1e-3) / 1e9 / hbm,
                                 "peak_source": src, "algorithmic_bytes": hb, "ms": ht}
    if not args.no_cpu_baseline and eff_world == 1:
        out["cpu_baseline"] = cpu_baseline(args.config, args.cpu_budget)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
