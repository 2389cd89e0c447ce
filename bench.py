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
    ap.add_argument("--cull", default="auto", choices=["auto", "tight", "reference"])
    ap.add_argument("--warmup-spp", type=int, default=0, help="spp of the warm-up steps (0 = the workload's; for minute-long frames)")
    ap.add_argument("--fused-e2e", action="store_true",
                    help="minute-long frames: time each step ONCE — the e2e call (upload + render into host memory) — and take "
                         "`value` from the device events of that same render instead of rendering every frame twice")
    ap.add_argument("--engine", default="auto", choices=["auto", "megakernel", "wavefront"], help="development: force an engine")
    ap.add_argument("--chunks", type=int, default=0, help="development: pin the sample-chunk count (WRT_FLAG_CHUNKS); 0 = the library's rule")
    ap.add_argument("--shard", default="rows", choices=["rows", "samples"], help="N > 1: what the devices split")
    ap.add_argument("--no-all-workloads", dest="all_workloads", action="store_false",
                    help="skip the short runs of the other BASELINE configs after the headline timing")
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


CPU_FLAGS = {"libwro.so": "-O2 -ffp-contract=off -fno-fast-math (portable build, the one the parity tests load)",
             "libwro_native.so": "-O3 -march=native -ffp-contract=off -fno-fast-math (built on this host)"}


def native_oracle():
    """oracle/libwro_native.so, built ON this host (so -march=native is this host's ISA); None when it cannot be built."""
    try:
        subprocess.check_call(["make", "-s", "-C", str(ROOT / "oracle"), "libwro_native.so"], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL, timeout=300)
        lib = ROOT / "oracle" / "libwro_native.so"
        return lib if lib.exists() else None
    except Exception:
        return None


def run_cpu(wl, spp_sample, seed, threads=None, lib=None):
    """One bounded CPU render of the workload with the oracle (the checker, timed here as the CPU baseline only).  `lib`
    selects the build; it runs in a child process because the binding loads one library per process."""
    if lib is not None:
        code = ("import json,sys; sys.path.insert(0, %r); import bench; wl = json.loads(sys.argv[1]); "
                "print(json.dumps(bench.run_cpu(wl, int(sys.argv[2]), int(sys.argv[3]))))" % str(ROOT))
        out = subprocess.run([sys.executable, "-c", code, json.dumps(wl), str(spp_sample), str(seed)], capture_output=True, text=True,
                             env=dict(os.environ, WRO_LIB=str(lib)), timeout=1800)
        if out.returncode != 0:
            raise RuntimeError("CPU leg failed: " + out.stderr[-400:])
        return tuple(json.loads(out.stdout.strip().splitlines()[-1]))
    sys.path.insert(0, str(ROOT / "oracle"))
    import wro_py as wro  # the checker, timed here as the CPU baseline only
    imgs, _ = scene_images(wl["scene"])
    sc = wro.OracleScene(wl["scene"], seed=1, n_prims=wl.get("n_prims", 0), images=imgs)
    cam = sc.camera(wl["width"], wl["height"])
    p = sc.params(wl["width"], wl["height"], spp_sample, wl["depth"], seed=seed)
    threads = threads or wro.host_threads()
    _, st = sc.render(cam, p, wro.RNG_REFERENCE, threads=threads)
    sc.close()
    return int(st.rays), int(st.paths), float(st.seconds), int(threads)


def cpu_baseline_record(wl, spp_sample, seed, W, H, depth):
    """The CPU number reported beside the GPU one: the faster of the two builds of the oracle, both stated with their flags."""
    # C5: the reference's culling visits every leaf (aabb.zig:80-101), so its sample is bounded in primitives AND frame
    # (2^14 primitives, a quarter of the frame's width and height): ~10 s instead of minutes
    wl_cpu = dict(wl, n_prims=1 << 14, width=wl["width"] // 4, height=wl["height"] // 4) if wl["key"] == "C5" else wl
    if wl["key"] == "C5":
        W, H = wl_cpu["width"], wl_cpu["height"]
    builds = {}
    native = native_oracle()
    for name, lib in (("libwro_native.so", native), ("libwro.so", ROOT / "oracle" / "libwro.so")):
        if lib is None:
            continue
        try:
            r, _, s_, thr = run_cpu(wl_cpu, spp_sample, seed, lib=lib)
            builds[name] = {"value": r / s_ / 1e6, "seconds": s_, "cores": thr, "flags": CPU_FLAGS[name]}
        except Exception as e:  # the portable build always exists; a failed native build is only noted
            builds[name] = {"error": str(e)[:200]}
    ok = {k: v for k, v in builds.items() if "value" in v}
    best = max(ok, key=lambda k: ok[k]["value"])
    return {"value": ok[best]["value"], "unit": UNIT, "cores": ok[best]["cores"], "kind": "port", "build": best,
            "flags": ok[best]["flags"], "seconds": ok[best]["seconds"],
            "sample": f"{spp_sample} of {wl['spp']} spp per pixel over the {W}x{H} frame, depth {depth}"
                      + (", 2^14 of 2^20 primitives, frame reduced 4x per side" if wl["key"] == "C5" else ""),
            "builds": builds,
            "note": "throughput is spp-independent; a full-spp CPU frame time quoted from it is an extrapolation"}


def reference_arm(args, wl):
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    spp_sample = cpu_sample_spp(wl, args.cpu_sample_spp)
    if wl["key"] == "C5":  # the reference's culling visits every leaf (aabb.zig:80-101): bound the sample
        wl = dict(wl, n_prims=1 << 14, width=wl["width"] // 4, height=wl["height"] // 4)
    sample = (f"{spp_sample} of {wl['spp']} spp per pixel over the {wl['width']}x{wl['height']} frame, depth {wl['depth']}"
              + (", 2^14 of 2^20 primitives, frame reduced 4x per side" if wl["key"] == "C5" else ""))
    native = native_oracle()
    lib = native or (ROOT / "oracle" / "libwro.so")
    for _ in range(args.warmup):
        run_cpu(wl, max(1, spp_sample // 4), args.seed, lib=lib)
    rays = secs = 0.0
    threads = 0
    for _ in range(args.steps):
        r, _, s, threads = run_cpu(wl, spp_sample, args.seed, lib=lib)
        rays += r; secs += s
    value = rays / secs / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "scene": wl["scene"], "width": wl["width"], "height": wl["height"], "spp": wl["spp"],
                   "depth": wl["depth"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample, "build": lib.name,
                         "flags": CPU_FLAGS[lib.name]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference = CPU port of the reference algorithm (oracle/, same job decomposition and culling); the Zig reference "
                "cannot be built in this image (no Zig toolchain)",
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# own arm
# ---------------------------------------------------------------------------------------------------------------------
def measure_workload(wrt, host, ctx, key, spp_override, seed, steps, warmup):
    """One short device-resident measurement of another BASELINE config on this GPU (the `workloads` sub-record)."""
    import torch
    wl = dict(WORKLOADS[key], key=key)
    spp = spp_override or wl["spp"]
    imgs, _ = scene_images(wl["scene"])
    scene = host.HostScene(wl["scene"], seed=1, synthetic_prims=wl.get("n_prims", 0), images=imgs)
    ctx.upload_scene(scene.flat())
    up_first_ms = ctx.stats().upload_ms  # first upload of a scene of this size on this context: pays the driver's first mapping of the buffers
    ctx.upload_scene(scene.flat())
    up_ms = ctx.stats().upload_ms
    tree_ms, tree_dev = ctx.stats().tree_build_ms, ctx.stats().tree_build_device
    W, H = wl["width"], wl["height"]
    cam = scene.camera(W, H)
    params = scene.params(W, H, spp, wl["depth"], seed=seed, cull_mode=wrt.WRT_CULL_AUTO)
    d_fb = torch.zeros((H, W, 4), dtype=torch.float64, device=f"cuda:{ctx.device}")
    for _ in range(warmup):
        ctx.render_device(cam, params, d_fb.data_ptr(), 32)
    rays = ms = kms = 0.0
    for _ in range(steps):
        ctx.render_device(cam, params, d_fb.data_ptr(), 32)
        st = ctx.stats()
        rays += st.rays; ms += st.render_ms; kms += st.kernel_ms
    hbm_peak, _ = measured_peaks()
    rec = {"workload": wl["name"], "spp": spp, "reduced_spp": spp != wl["spp"], "steps": steps, "warmup": warmup,
           "value": rays / (ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms / steps, "rays_per_step": rays / steps,
           "upload_ms": up_ms, "upload_ms_first": up_first_ms, "tree_build_ms": tree_ms, "tree_build_on_device": bool(tree_dev), "traversal": ("packet" if ctx.stats().program_ops <= 96 else
                         ("wavefront, persistent per-lane extend over four-wide records" if ctx.stats().n_tree_records >= 16384 and tree_dev
                          else "ordered, per lane")),
           "roofline_hbm_frac": (rays / (kms * 1e-3)) * wl["b_ray"] / 1e9 / hbm_peak,
           "algorithmic_bytes_per_ray": wl["b_ray"], "algorithmic_fp64_instr_per_ray": wl["f_ray"]}
    scene.close() if hasattr(scene, "close") else None
    return rec


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
    cull = {"auto": wrt.WRT_CULL_AUTO, "tight": wrt.WRT_CULL_TIGHT, "reference": wrt.WRT_CULL_REFERENCE}[args.cull]
    flags = wrt.WRT_FLAG_SHARD_SAMPLES if args.shard == "samples" else 0
    flags |= {"auto": 0, "megakernel": wrt.WRT_FLAG_ENGINE_MEGAKERNEL, "wavefront": wrt.WRT_FLAG_ENGINE_WAVEFRONT}[args.engine]
    if args.chunks:
        flags |= wrt.WRT_FLAG_CHUNKS(args.chunks)
    imgs, img_note = scene_images(wl["scene"])
    scene = host.HostScene(wl["scene"], seed=1, synthetic_prims=wl.get("n_prims", 0), images=imgs)
    flat = scene.flat()
    cam = scene.camera(W, H)
    ctx = wrt.Context(local)
    ctx.upload_scene(flat)
    LANES = 4
    if dist:
        # multi-GPU goes through the library's own boundary (include/wrt.h, Multi-GPU (2)): rank 0 makes the NCCL id, torch
        # only carries it to the other ranks; the shard gather runs inside wrt_render_sharded
        box = [wrt.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.comm_init(box[0], rank, world)
    params = scene.params(W, H, spp, depth, seed=args.seed, cull_mode=cull, flags=flags)
    d_fb = torch.zeros((H, W, LANES), dtype=torch.float64, device=f"cuda:{local}") if not dist else None
    h_fb = torch.zeros((H, W, LANES), dtype=torch.float64).pin_memory() if rank == 0 else None

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        """Scene resident, frame left in HBM (rank 0's for N > 1: shards rendered, gathered and assembled on the device)."""
        if dist:
            ctx.render_sharded(cam, params, out=None)
        else:
            ctx.render_device(cam, params, d_fb.data_ptr(), LANES * 8)
        st = ctx.stats()
        return st.rays, st.paths, st.render_ms, st.kernel_ms, st.gather_ms, st.kernel_launches

    fp64_peak = ctx.fp64_issue_peak() if rank == 0 else 0.0
    fp32_peak = ctx.fp32_issue_peak() if rank == 0 else 0.0

    if args.warmup_spp:
        full_spp = params.samples_per_pixel
        params.samples_per_pixel = args.warmup_spp
    for _ in range(args.warmup):
        device_step()
    if args.warmup_spp:
        params.samples_per_pixel = full_spp

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t_wall0 = time.perf_counter()
    rays = paths = 0
    dev_ms = kern_ms = gath_ms = 0.0
    launches = 0
    e2e_rays = 0
    for _ in range(args.steps):
        if args.fused_e2e:  # one render per step: the user-facing call, device-timed inside
            ctx.upload_scene(flat)
            if dist:
                ctx.render_sharded(cam, params, out=h_fb.numpy() if rank == 0 else None)
            else:
                ctx.render(cam, params, lanes=LANES, out=h_fb.numpy())
            st = ctx.stats()
            r, p_, ms, kms, gms, nl = st.rays, st.paths, st.kernel_ms, st.kernel_ms, st.gather_ms, st.kernel_launches
            e2e_rays += r
        else:
            r, p_, ms, kms, gms, nl = device_step()
        rays += r; paths += p_; dev_ms += ms + gms; kern_ms += kms; gath_ms += gms; launches += nl
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall0)
    clocks = sampler.stop() if rank == 0 else None

    # whole-job aggregate: sum of rays over ranks / max time over ranks
    t = torch.tensor([dev_ms, kern_ms, wall_ms, gath_ms], dtype=torch.float64, device=f"cuda:{local}")
    c = torch.tensor([rays, paths, launches], dtype=torch.float64, device=f"cuda:{local}")
    k_all = [torch.zeros(1, dtype=torch.float64, device=f"cuda:{local}") for _ in range(world)]
    if dist:
        dist.all_gather(k_all, torch.tensor([kern_ms / args.steps], dtype=torch.float64, device=f"cuda:{local}"))
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    else:
        k_all[0][0] = kern_ms / args.steps
    kernel_ms_per_rank = [float(k.item()) for k in k_all]
    dev_ms_max, kern_ms_max, wall_ms_max, gath_ms_max = t.tolist()
    rays_all, paths_all, launches_all = c.tolist()
    value = rays_all / (dev_ms_max * 1e-3) / 1e6

    # ---- e2e: host buffers in, host framebuffer out, every step ----
    barrier()
    t0 = time.perf_counter()
    for _ in range(0 if args.fused_e2e else args.steps):
        ctx.upload_scene(flat)  # H2D of the scene arrays (the step's inputs)
        if dist:
            ctx.render_sharded(cam, params, out=h_fb.numpy() if rank == 0 else None)  # gather + D2H of the full frame on rank 0
        else:
            ctx.render(cam, params, lanes=LANES, out=h_fb.numpy())  # D2H into the pinned host framebuffer
        e2e_rays += ctx.stats().rays
    barrier()
    e2e_s = (wall_ms * 1e-3) if args.fused_e2e else (time.perf_counter() - t0)
    e = torch.tensor([e2e_rays], dtype=torch.float64, device=f"cuda:{local}")
    ts = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{local}")
    if dist:
        dist.all_reduce(e, op=dist.ReduceOp.SUM)
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
    e2e_value = e.item() / ts.item() / 1e6
    h2d = (scene.input_bytes() + 256 + 136) * n_gpus  # scene arrays + camera + params structs, to every device
    d2h = H * W * LANES * 8 + 32 * n_gpus            # framebuffer (rank 0) + ray counters
    stats0 = ctx.stats()

    if rank != 0:
        if dist:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    mean = float(np.nanmean(h_fb.numpy()[..., :3]))
    hbm_peak, peak_src = measured_peaks()
    rays_per_s_kernel = rays_all / (kern_ms_max * 1e-3)
    ach_gbs = rays_per_s_kernel * wl["b_ray"] / 1e9 / n_gpus  # per GPU, the kernel alone
    ach_issue = rays_per_s_kernel * wl["f_ray"] / n_gpus
    cache_resident = wl["key"] != "C5"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "scene": wl["scene"], "width": W, "height": H, "spp": spp, "depth": depth,
                   "engine": args.engine, "sample_chunks": (args.chunks or "library rule"), "cull": args.cull + (" -> " + ("reference" if stats0.cull_mode_used == wrt.WRT_CULL_REFERENCE else "tight")),
                   "parallelism": (f"{'sample-range' if args.shard == 'samples' else 'row-interleaved'} shards x{n_gpus}"
                                   + (" + NCCL gather inside wrt_render_sharded" if n_gpus > 1 else "")),
                   "l2_policy": "scene is cache-resident by design; per-step traffic is the framebuffer (> L2 only for C5)",
                   "textures": img_note, "reduced_spp": bool(args.spp), "reduced_frame": bool(args.res),
                   "warmup_spp": args.warmup_spp or spp,
                   "timing": ("fused: each step is ONE e2e call (upload + render into host memory); `value` = rays / device-event "
                              "time of the kernel + gather of those same calls") if args.fused_e2e else "separate device-resident and e2e steps"},
        "wall_ms_per_step": wall_ms_max / args.steps,
        "rays_per_step": rays_all / args.steps, "paths_per_step": paths_all / args.steps,
        "gather_ms_per_step": gath_ms_max / args.steps,
        "kernel_ms_per_rank": {"min": min(kernel_ms_per_rank), "max": max(kernel_ms_per_rank), "all": kernel_ms_per_rank},
        "mean_radiance": mean,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches_all),
        "roofline": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                     "traffic": (wl["dram_b_per_ray_ncu"] * rays_all / args.steps / n_gpus) if "dram_b_per_ray_ncu" in wl else None,
                     "traffic_unit": "bytes per launch (ncu DRAM bytes per ray x rays of this launch; source in profiles/README.md)",
                     "algorithmic_bytes_per_launch": wl["b_ray"] * rays_all / args.steps / n_gpus,
                     "kernel": ("wavefront iteration loop (wf_extend_ordered_kernel = 81 % of its GPU time, profiles/README.md)"
                                if launches_all / args.steps / n_gpus > 16 else "render_kernel"),
                     "kernel_ms_per_launch": kern_ms_max / args.steps,
                     "algorithmic_bytes_per_ray": wl["b_ray"], "peak_source": peak_src,
                     "note": ("NOMINAL for this workload: the scene is cache-resident (real DRAM traffic is `traffic`, a fraction of a "
                              "percent of the algorithmic bytes), so this is not achieved HBM bandwidth; the binding bound is "
                              "roofline_issue") if cache_resident else
                             "2^20 primitives: node / primitive fetches miss L1 and contend in L2; HBM is the bound to report"},
        "roofline_issue": {"bound": "fp64_issue", "achieved": ach_issue / 1e9, "peak": fp64_peak / 1e9, "unit": "G FP64 instr/s",
                           "frac": (ach_issue / fp64_peak) if fp64_peak else None, "algorithmic_fp64_instr_per_ray": wl["f_ray"],
                           "peak_source": "DFMA micro-kernel measured live (wrt_fp64_issue_peak)",
                           "fp32_issue_peak": fp32_peak / 1e9, "fp32_note": "FFMA micro-kernel (wrt_fp32_issue_peak): the pipe the box tests run on"},
    }
    if not args.no_cpu_baseline and n_gpus == 1:
        line["cpu_baseline"] = cpu_baseline_record(wl, cpu_sample_spp(wl, args.cpu_sample_spp), args.seed, W, H, depth)
    if args.all_workloads and n_gpus == 1 and not args.spp and not args.res:
        # the other BASELINE configs, measured by whoever runs this command (short: 1 warm-up + 2 steps each; C5 at reduced spp,
        # its full-spp campaign is in profiles/)
        others = {}
        for key, spp_o in (("C1", 0), ("C3", 0), ("C4", 0), ("C5", 16)):
            if key == wl["key"]:
                continue
            try:
                others[key] = measure_workload(wrt, host, ctx, key, spp_o, args.seed, steps=2, warmup=1)
            except Exception as ex:  # a sub-record must never lose the headline
                others[key] = {"error": str(ex)[:200]}
        line["workloads"] = others
    print(json.dumps(line))
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
