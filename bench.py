#!/usr/bin/env python3
"""bench.py — the headline benchmark of the render hot path on B200 (contract: see the task statement / DESIGN.md §7).

    python bench.py --gpus N --steps K --warmup W            # N = 1; for N > 1 launch under torch.distributed.run
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port) on host cores

A *step* is one full render of the named workload (default: BASELINE.json configs[1], Cornell box with light importance
sampling, 1024x1024, 10 000 spp, depth 50) through the C ABI of include/wrt.h.  Metric: Mrays/s, a ray being one
closest-hit query issued by the integrator (render.zig:215), counted on the device.

  value     device-resident: scene already uploaded, frame left in HBM (wrt_render_device); time = CUDA events on the
            launching stream inside the library (wrt_stats.render_ms), max over ranks; N > 1 adds the NCCL gather.
  e2e       the user-facing call with HOST buffers: wrt_upload_scene (H2D of the scene arrays) + wrt_render into a pinned
            host framebuffer (D2H), wall clock around the calls with a device synchronize on both sides.
  roofline  dominant kernel = render_kernel; algorithmic bytes per ray from SURVEY.md §8(d) x rays per launch / its
            event-timed duration, against the measured HBM copy peak (MEASURED_PEAKS.json).  These configs are
            cache-resident, so the binding bound is FP64 issue: `roofline_issue` reports rays/s x F_ray against a DFMA
            micro-kernel measured live.
  cpu_baseline  the oracle (CPU port of the reference algorithm, reference-like RNG) on all host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

# BASELINE.json configs -> workloads.  B_ray / F_ray: algorithmic bytes and FP64 instructions per ray, SURVEY.md §8(d).
WORKLOADS = {
    "C1": dict(name="README CLI example: emissive 400x400, 128 spp, depth 10", scene="emissive", width=400, height=400, spp=128,
               depth=10, b_ray=204, f_ray=285),
    "C2": dict(name="cornell_box 1024x1024, 10000 spp, depth 50 (light importance sampling, pdf.zig mixture)", scene="cornell_box",
               width=1024, height=1024, spp=10000, depth=50, b_ray=332, f_ray=354,
               # dram__bytes_read.sum + dram__bytes_write.sum of render_kernel / rays of that launch, from the ncu --set full
               # capture of this workload at 64 spp (profiles/r01_ncu_render_kernel_packet_summary.txt): 471.5 MB / 3.38e8 rays
               dram_b_per_ray_ncu=1.395),
    "C3": dict(name="balls (book-1 final, ~484 spheres through BVH) 1920x1080, 512 spp, depth 50", scene="balls", width=1920,
               height=1080, spp=512, depth=50, b_ray=572, f_ray=658),
    "C4": dict(name="earth (image-textured sphere + diffuse lights) 1920x1080, 1024 spp, depth 20", scene="earth", width=1920,
               height=1080, spp=1024, depth=20, b_ray=240, f_ray=335),
    "C5": dict(name="synthetic 2^20 spheres/quads 3840x2160, 1024 spp, depth 20", scene="synthetic", width=3840, height=2160,
               spp=1024, depth=20, b_ray=1236, f_ray=1262, n_prims=1 << 20),
}
METRIC = "Mrays/s (samples x bounces) and render wall-time at 1/2/4/8 B200 vs host-CPU ref"
UNIT = "Mrays/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="wrt", choices=["wrt", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--cull", default="tight", choices=["tight", "reference"])
    ap.add_argument("--spp", type=int, default=0, help="development only: override samples per pixel (marks the line reduced)")
    ap.add_argument("--res", default="", help="development only: override the frame as WxH (marks the line reduced)")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cpu-sample-spp", type=int, default=0, help="spp of the bounded CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json, copy bandwidth)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def scene_images(scene: str):
    """Texel bytes for image-textured workloads: the reference's assets as decoded by the reference's own vendored stb_image
    (tools/make_texel_fixtures.py; the reference checkout does not travel to the GPU box).  C4's earth.png is full resolution."""
    if scene not in ("earth", "rtw_final", "shrek_quads"):
        return None, None
    assets = importlib.import_module("zig-weekend-raytracer_b200.assets")
    return assets.reference_images(), ("reference assets decoded by the reference's stb_image v2.28 (data/texels; earth.png and "
                                       "wap.jpg full resolution, me.jpg decimated 4x)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.rows = []
        self.proc = None
        self.thread = None
        self.device_index = device_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device_index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.rows.append(line.strip())
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ---------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own algorithm on the host cores (the oracle port; the Zig binary cannot be built here)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_sample_spp(wl, override):
    if override:
        return override
    return {"C1": 16, "C2": 16, "C3": 4, "C4": 8, "C5": 1}[wl["key"]]


def run_cpu(wl, spp_sample, seed, threads=None):
    sys.path.insert(0, str(ROOT / "oracle"))
    import wro_py as wro  # the checker, timed here as the CPU baseline only
    imgs, _ = scene_images(wl["scene"])
    sc = wro.OracleScene(wl["scene"], seed=1, n_prims=wl.get("n_prims", 0), images=imgs)
    cam = sc.camera(wl["width"], wl["height"])
    p = sc.params(wl["width"], wl["height"], spp_sample, wl["depth"], seed=seed)
    threads = threads or wro.host_threads()
    _, st = sc.render(cam, p, wro.RNG_REFERENCE, threads=threads)
    sc.close()
    return st.rays, st.paths, st.seconds, threads


def reference_arm(args, wl):
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    spp_sample = cpu_sample_spp(wl, args.cpu_sample_spp)
    if wl["key"] == "C5":
        wl = dict(wl, n_prims=1 << 14)  # the reference's culling visits every leaf (aabb.zig:80-101): bound the sample
    sample = (f"{spp_sample} of {wl['spp']} spp per pixel over the full {wl['width']}x{wl['height']} frame, depth {wl['depth']}"
              + (", 2^14 of 2^20 primitives" if wl["key"] == "C5" else ""))
    for _ in range(args.warmup):
        run_cpu(wl, max(1, spp_sample // 4), args.seed)
    rays = secs = 0.0
    threads = 0
    for _ in range(args.steps):
        r, _, s, threads = run_cpu(wl, spp_sample, args.seed)
        rays += r; secs += s
    value = rays / secs / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "scene": wl["scene"], "width": wl["width"], "height": wl["height"], "spp": wl["spp"],
                   "depth": wl["depth"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference = CPU port of the reference algorithm (oracle/, same job decomposition and culling); the Zig reference "
                "cannot be built in this image (no Zig toolchain)",
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# own arm
# ---------------------------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    wl = dict(WORKLOADS[args.workload], key=args.workload)
    if args.res:
        wl["width"], wl["height"] = (int(v) for v in args.res.lower().split("x"))
    if args.impl == "reference":
        return reference_arm(args, wl)

    rank, world, local = dist_env()
    if world != args.gpus and world > 1:
        print(f"warning: WORLD_SIZE={world} but --gpus {args.gpus}", file=sys.stderr)
    n_gpus = world

    import numpy as np
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the render back end has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))

    wrt = importlib.import_module("zig-weekend-raytracer_b200")
    host = importlib.import_module("zig-weekend-raytracer_b200.host")

    spp = args.spp or wl["spp"]
    W, H, depth = wl["width"], wl["height"], wl["depth"]
    cull = wrt.WRT_CULL_TIGHT if args.cull == "tight" else wrt.WRT_CULL_REFERENCE
    imgs, img_note = scene_images(wl["scene"])
    scene = host.HostScene(wl["scene"], seed=1, synthetic_prims=wl.get("n_prims", 0), images=imgs)
    flat = scene.flat()
    cam = scene.camera(W, H)
    ctx = wrt.Context(local)
    ctx.upload_scene(flat)
    params = scene.params(W, H, spp, depth, seed=args.seed, cull_mode=cull, row_shard_index=rank, row_shard_count=world)
    rows_local = ctx.local_rows(params)
    rows_pad = (H + world - 1) // world
    LANES = 4
    d_fb = torch.zeros((rows_pad, W, LANES), dtype=torch.float64, device=f"cuda:{local}")
    gather_list = [torch.zeros_like(d_fb) for _ in range(world)] if (world > 1 and rank == 0) else None
    d_full = torch.zeros((H, W, LANES), dtype=torch.float64, device=f"cuda:{local}") if rank == 0 else None
    h_fb = torch.zeros((H, W, LANES), dtype=torch.float64).pin_memory() if rank == 0 else None

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    distributed = importlib.import_module("zig-weekend-raytracer_b200.distributed")

    def gather_frame():
        """NCCL gather of the row shards to rank 0 + interleave into the full frame (N > 1 only)."""
        if dist:
            distributed.gather_frame(dist, d_fb, H, rank, world, out=d_full, gather_list=gather_list)

    def device_step():
        ctx.render_device(cam, params, d_fb.data_ptr(), LANES * 8)
        st = ctx.stats()
        g_ms = 0.0
        if dist:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); gather_frame(); e1.record(); torch.cuda.synchronize()
            g_ms = e0.elapsed_time(e1)
        return st.rays, st.paths, st.render_ms, st.kernel_ms, g_ms, st.kernel_launches

    fp64_peak = ctx.fp64_issue_peak() if rank == 0 else 0.0
    fp32_peak = ctx.fp32_issue_peak() if rank == 0 else 0.0

    for _ in range(args.warmup):
        device_step()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t_wall0 = time.perf_counter()
    rays = paths = 0
    dev_ms = kern_ms = gath_ms = 0.0
    launches = 0
    for _ in range(args.steps):
        r, p_, ms, kms, gms, nl = device_step()
        rays += r; paths += p_; dev_ms += ms + gms; kern_ms += kms; gath_ms += gms; launches += nl
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall0)
    clocks = sampler.stop() if rank == 0 else None

    # whole-job aggregate: sum of rays over ranks / max time over ranks
    t = torch.tensor([dev_ms, kern_ms, wall_ms, gath_ms], dtype=torch.float64, device=f"cuda:{local}")
    c = torch.tensor([rays, paths, launches], dtype=torch.float64, device=f"cuda:{local}")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    dev_ms_max, kern_ms_max, wall_ms_max, gath_ms_max = t.tolist()
    rays_all, paths_all, launches_all = c.tolist()
    value = rays_all / (dev_ms_max * 1e-3) / 1e6

    # ---- e2e: host buffers in, host framebuffer out, every step ----
    barrier()
    t0 = time.perf_counter()
    e2e_rays = 0
    for _ in range(args.steps):
        ctx.upload_scene(flat)  # H2D of the scene arrays (the step's inputs)
        if dist:
            ctx.render_device(cam, params, d_fb.data_ptr(), LANES * 8)
            e2e_rays += ctx.stats().rays
            gather_frame()
            if rank == 0:
                h_fb.copy_(d_full, non_blocking=False)
        else:
            ctx.render(cam, params, lanes=LANES, out=h_fb.numpy())  # D2H into the pinned host framebuffer
            e2e_rays += ctx.stats().rays
    barrier()
    e2e_s = time.perf_counter() - t0
    e = torch.tensor([e2e_rays], dtype=torch.float64, device=f"cuda:{local}")
    ts = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{local}")
    if dist:
        dist.all_reduce(e, op=dist.ReduceOp.SUM)
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
    e2e_value = e.item() / ts.item() / 1e6
    h2d = scene.input_bytes() + 256 + 136  # scene arrays + camera + params structs
    d2h = H * W * LANES * 8 + 32           # framebuffer + ray counters

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return 0

    mean = float(np.nanmean(h_fb.numpy()[..., :3]))
    hbm_peak, peak_src = measured_peaks()
    rays_per_s_kernel = rays_all / (kern_ms_max * 1e-3)
    ach_gbs = rays_per_s_kernel * wl["b_ray"] / 1e9 / n_gpus  # per GPU, the kernel alone
    ach_issue = rays_per_s_kernel * wl["f_ray"] / n_gpus
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "scene": wl["scene"], "width": W, "height": H, "spp": spp, "depth": depth,
                   "cull": args.cull, "parallelism": f"row-interleaved shards x{n_gpus}" + (" + NCCL gather" if n_gpus > 1 else ""),
                   "l2_policy": "scene is cache-resident by design; per-step traffic is the framebuffer (> L2 only for C5)",
                   "textures": img_note, "reduced_spp": bool(args.spp), "reduced_frame": bool(args.res)},
        "wall_ms_per_step": wall_ms_max / args.steps,
        "rays_per_step": rays_all / args.steps, "paths_per_step": paths_all / args.steps,
        "gather_ms_per_step": gath_ms_max / args.steps,
        "mean_radiance": mean,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches_all),
        "roofline": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                     "traffic": (wl["dram_b_per_ray_ncu"] * rays_all / args.steps) if "dram_b_per_ray_ncu" in wl else None,
                     "traffic_unit": "bytes per launch (ncu DRAM bytes per ray at 64 spp x rays of this launch)",
                     "algorithmic_bytes_per_launch": wl["b_ray"] * rays_all / args.steps,
                     "kernel": "render_kernel", "kernel_ms_per_launch": kern_ms_max / args.steps,
                     "algorithmic_bytes_per_ray": wl["b_ray"], "peak_source": peak_src,
                     "note": "cache-resident scene: HBM is not the binding bound here, see roofline_issue"},
        "roofline_issue": {"bound": "fp64_issue", "achieved": ach_issue / 1e9, "peak": fp64_peak / 1e9, "unit": "G FP64 instr/s",
                           "frac": (ach_issue / fp64_peak) if fp64_peak else None, "algorithmic_fp64_instr_per_ray": wl["f_ray"],
                           "peak_source": "DFMA micro-kernel measured live (wrt_fp64_issue_peak)",
                           "fp32_issue_peak": fp32_peak / 1e9, "fp32_note": "FFMA micro-kernel (wrt_fp32_issue_peak): the pipe the box tests run on"},
    }
    if not args.no_cpu_baseline and n_gpus == 1:
        spp_s = cpu_sample_spp(wl, args.cpu_sample_spp)
        wl_cpu = dict(wl, n_prims=1 << 14) if wl["key"] == "C5" else wl
        r, _, s, thr = run_cpu(wl_cpu, spp_s, args.seed)
        line["cpu_baseline"] = {"value": r / s / 1e6, "unit": UNIT, "cores": thr, "kind": "port",
                                "sample": f"{spp_s} of {wl['spp']} spp per pixel over the full {W}x{H} frame, depth {depth}"
                                          + (", 2^14 of 2^20 primitives" if wl["key"] == "C5" else ""), "seconds": s}
    print(json.dumps(line))
    if dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
