#!/bin/bash
# where the time of an upload goes when a context moves from small scenes to the 2^20-primitive one (WRT_TRACE_BUILD=1)
set -u
mkdir -p gpurun_out
WRT_TRACE_BUILD=1 python - > gpurun_out/r02_trace3.txt 2>&1 <<'P'
import sys, importlib, time, os
sys.path.insert(0,'.')
wrt = importlib.import_module("zig-weekend-raytracer_b200")
host = importlib.import_module("zig-weekend-raytracer_b200.host")
import torch
big = host.HostScene("synthetic", seed=1, synthetic_prims=1<<20)
with wrt.Context(0) as ctx:
    for name, w, h, spp in (("cornell_box", 1024, 1024, 64), ("balls", 1920, 1080, 8)):
        sc = host.HostScene(name, seed=1)
        ctx.upload_scene(sc.flat())
        cam = sc.camera(w, h); params = sc.params(w, h, spp, 20, seed=1, cull_mode=wrt.WRT_CULL_AUTO)
        fb = torch.zeros((h, w, 4), dtype=torch.float64, device="cuda:0")
        ctx.render_device(cam, params, fb.data_ptr(), 32)
        print(name, "upload_ms", round(ctx.stats().upload_ms, 2), flush=True)
    for k in range(3):
        t = time.time(); ctx.upload_scene(big.flat()); dt = (time.time() - t) * 1e3
        print("synthetic upload", k, "wall", round(dt, 1), "upload_ms", round(ctx.stats().upload_ms, 1), "tree", round(ctx.stats().tree_build_ms, 1), flush=True)
P
echo "rc=$?"; grep -v "^wrt trace: compile: records 0.0" gpurun_out/r02_trace3.txt | tail -40
