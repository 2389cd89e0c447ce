"""Development tool: upload the 2^20-primitive scene once, then render the C5 frame (3840x2160, depth 20) at a reduced sample
count under a list of environment settings that wrt_render reads per call (WRT_WF_NODE_BURST, WRT_WF_LEAF_BURST,
WRT_WF_NODE_SHIFT, WRT_WF_POOL, WRT_WF_PIPELINES, WRT_WF_SORT, WRT_WF_SORT_SHIFT).  Prints Mrays/s per setting and checks that
the frame's mean radiance does not move.  usage: python tools/c5_sweep.py SPP "A=1 B=2" "A=3" ...  ("" = defaults)"""
import importlib
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
wrt = importlib.import_module("zig-weekend-raytracer_b200")
host = importlib.import_module("zig-weekend-raytracer_b200.host")
import torch  # noqa: E402

spp = int(sys.argv[1])
settings = sys.argv[2:] or [""]
W, H, depth = 3840, 2160, 20
scene = host.HostScene("synthetic", seed=1, synthetic_prims=1 << 20)
with wrt.Context(0) as ctx:
    ctx.upload_scene(scene.flat())
    cam = scene.camera(W, H)
    d_fb = torch.zeros((H, W, 4), dtype=torch.float64, device="cuda:0")
    warm = scene.params(W, H, 2, depth, seed=1, cull_mode=wrt.WRT_CULL_AUTO)
    ctx.render_device(cam, warm, d_fb.data_ptr(), 32)
    params = scene.params(W, H, spp, depth, seed=1, cull_mode=wrt.WRT_CULL_AUTO)
    ref_mean = None
    for s in settings:
        pairs = [kv.split("=", 1) for kv in s.split()]
        for k, v in pairs:
            os.environ[k] = v
        ctx.render_device(cam, params, d_fb.data_ptr(), 32)
        st = ctx.stats()
        mean = float(d_fb[..., :3].mean().item())
        ref_mean = mean if ref_mean is None else ref_mean
        print(f"{s or 'defaults':60s} {st.rays / (st.render_ms * 1e-3) / 1e6:8.1f} Mrays/s  {st.render_ms:9.1f} ms  launches {st.kernel_launches}"
              f"  mean {mean!r}{'' if mean == ref_mean else '  <-- FRAME DIFFERS'}", flush=True)
        for k, _ in pairs:
            del os.environ[k]
