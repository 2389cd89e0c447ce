#!/usr/bin/env python3
"""Regenerates tests/golden/*.npz from the oracle (run in the build container).

The reference (Zig) cannot be executed here, and its tests hold no vectors for this path, so these fixtures pin the
ORACLE's outputs (regression guard + the vectors the CUDA path is compared with on the GPU box):
  prim_ids / t_bits : gate-1 dump, primary-ray closest hit of samples [0, n_primary) of every pixel
  radiance          : linear f64 frame of the counter-RNG render (Philox stream shared with the device)
Image-textured scenes use the procedural stand-in texels of oracle/wro_py.procedural_image (the reference assets do
not travel to the GPU box).  `assets_*.npz` additionally records checksums of the reference's own assets decoded by
the reference's vendored stb_image (oracle/_ref/libstbi.so) when /root/reference is mounted.
"""
import ctypes as C
import hashlib
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT / "oracle"))
import wro_py as wro  # noqa: E402

CASES = {
    # name: (width, height, spp, depth, n_primary, scene_seed, n_prims_arg, render seed)
    "cornell_box": (64, 64, 8, 50, 4, 1, 0, 42),
    "emissive": (50, 50, 8, 10, 4, 1, 0, 42),
    "balls": (64, 36, 4, 50, 4, 1, 0, 42),
    "rtw_final": (48, 48, 4, 20, 2, 1, 0, 42),
    "shrek_quads": (40, 40, 4, 10, 4, 1, 0, 42),
    "earth": (48, 27, 4, 20, 4, 1, 0, 42),
    "synthetic": (32, 18, 2, 20, 2, 1, 2048, 42),
}


def main():
    images = {
        "wap.jpg": wro.procedural_image("wap.jpg", 300, 292),
        "me.jpg": wro.procedural_image("me.jpg", 231, 308),
        "earth.png": wro.procedural_image("earth.png", 512, 256),
    }
    for name, (w, h, spp, depth, n_primary, scene_seed, n_prims_arg, seed) in CASES.items():
        sc = wro.OracleScene(name, seed=scene_seed, n_prims=n_prims_arg, images=images)
        cam = sc.camera(w, h)
        p = sc.params(w, h, spp, depth, seed=seed)
        ids, t = sc.primary_hits(cam, p, n_primary)
        fb, st = sc.render(cam, p, wro.RNG_COUNTER)
        np.savez_compressed(HERE / f"{name}.npz", width=w, height=h, spp=spp, depth=depth, n_primary=n_primary,
                            scene_seed=scene_seed, n_prims_arg=n_prims_arg, seed=seed, prim_ids=ids,
                            t_bits=t.view(np.uint64), radiance=fb[..., :3].copy(), rays=st.rays, paths=st.paths,
                            n_prims=sc.n_prims)
        print(name, "prims", sc.n_prims, "rays", st.rays, "hit fraction", float((ids != 0xFFFFFFFF).mean()))
        sc.close()

    stbi = ROOT / "oracle" / "_ref" / "libstbi.so"
    assets = Path("/root/reference/assets")
    if stbi.exists() and assets.exists():
        lib = C.CDLL(str(stbi))
        lib.wro_stbi_load.restype = C.POINTER(C.c_ubyte)
        lib.wro_stbi_load.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.wro_stbi_free.argtypes = [C.POINTER(C.c_ubyte)]
        out = {}
        for fn in ["earth.png", "wap.jpg", "me.jpg"]:
            w, h, c = C.c_int(), C.c_int(), C.c_int()
            ptr = lib.wro_stbi_load(str(assets / fn).encode(), C.byref(w), C.byref(h), C.byref(c))
            arr = np.ctypeslib.as_array(ptr, shape=(h.value, w.value, c.value)).copy()
            lib.wro_stbi_free(ptr)
            key = fn.replace(".", "_")
            out[key + "_shape"] = np.array(arr.shape)
            out[key + "_sha256"] = np.frombuffer(hashlib.sha256(arr.tobytes()).digest(), np.uint8)
            out[key + "_corner"] = arr[:4, :4].copy()
            if fn == "wap.jpg":
                out[key + "_pixels"] = arr  # 300x292x3 = 263 KB raw, compresses well
            print(fn, arr.shape)
        np.savez_compressed(HERE / "assets_reference_decode.npz", **out)


if __name__ == "__main__":
    main()
