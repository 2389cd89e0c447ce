#!/usr/bin/env python3
"""Regenerates tests/golden/*.npz from the oracle (run in the build container).

The reference (Zig) cannot be executed here, and its tests hold no vectors for this path, so these fixtures pin the
ORACLE's outputs (regression guard + the vectors the CUDA path is compared with on the GPU box):
  prim_ids / t_bits : gate-1 dump, primary-ray closest hit of samples [0, n_primary) of every pixel
  radiance          : linear f64 frame of the counter-RNG render (Philox stream shared with the device)
Image-textured scenes use the reference's assets as decoded by the reference's vendored stb_image
(tools/make_texel_fixtures.py -> zig-weekend-raytracer_b200/data/texels; me.jpg decimated 4x).
"""
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
    sys.path.insert(0, str(ROOT))
    import importlib
    images = importlib.import_module("zig-weekend-raytracer_b200.assets").reference_images()
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


if __name__ == "__main__":
    main()
