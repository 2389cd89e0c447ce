"""Development tool: one eighth of the C5 frame (row shard 0 of 8) on one GPU under different pool sizes — what each device of
an 8-GPU render does.  usage: python tools/c5_shard_sweep.py SPP SHARDS "WRT_WF_POOL=4" ..."""
import importlib
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
wrt = importlib.import_module("zig-weekend-raytracer_b200")
host = importlib.import_module("zig-weekend-raytracer_b200.host")
import torch  # noqa: E402

spp, shards = int(sys.argv[1]), int(sys.argv[2])
settings = sys.argv[3:] or [""]
W, H, depth = 3840, 2160, 20
scene = host.HostScene("synthetic", seed=1, synthetic_prims=1 << 20)
with wrt.Context(0) as ctx:
    ctx.upload_scene(scene.flat())
    cam = scene.camera(W, H)
    rows = (H + shards - 1) // shards
    d_fb = torch.zeros((rows, W, 4), dtype=torch.float64, device="cuda:0")
    def params(s, chunks=0):
        return scene.params(W, H, s, depth, seed=1, cull_mode=wrt.WRT_CULL_AUTO, row_shard_index=0, row_shard_count=shards,
                            flags=wrt.WRT_FLAG_CHUNKS(chunks) if chunks else 0)
    ctx.render_device(cam, params(2), d_fb.data_ptr(), 32)
    for s in settings:
        pairs = [kv.split("=", 1) for kv in s.split()]
        chunks = 0
        for k, v in pairs:
            if k == "CHUNKS":   # pseudo-setting: WRT_FLAG_CHUNKS(n) in the params
                chunks = int(v)
            os.environ[k] = v
        ctx.render_device(cam, params(spp, chunks), d_fb.data_ptr(), 32)
        st = ctx.stats()
        print(f"{s or 'defaults':40s} {st.rays / (st.render_ms * 1e-3) / 1e6:8.1f} Mrays/s  {st.render_ms:9.1f} ms  launches {st.kernel_launches}"
              f"  mean {float(d_fb[..., :3].mean().item())!r}", flush=True)
        for k, _ in pairs:
            del os.environ[k]
