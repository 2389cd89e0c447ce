"""Texels of the reference's image assets as the reference's own decoder produces them (tools/make_texel_fixtures.py).

The Zig host decodes assets/*.jpg|png with the vendored stb_image (src/image.zig:12-17) and hands the bytes over the C ABI
(wrt_scene.texels); the C++ host mirror does the same when it was built against that header (host/wrh_image.cpp).  Where the
reference checkout is absent (the GPU box), tests and bench.py take the same bytes from data/texels/."""
from __future__ import annotations

import hashlib
import json
import lzma
from pathlib import Path

import numpy as np

TEXEL_DIR = Path(__file__).resolve().parent / "data" / "texels"


def manifest() -> dict:
    return json.loads((TEXEL_DIR / "manifest.json").read_text())


def reference_texels(name: str) -> np.ndarray:
    """uint8 [H][W][C] texels of assets/<name> (me.jpg: decimated 4x, see the manifest)."""
    info = manifest()["images"][name]
    raw = np.frombuffer(lzma.decompress((TEXEL_DIR / f"{name}.rgb8.xz").read_bytes()), np.uint8).reshape(info["shape"])
    out = np.cumsum(raw.astype(np.uint32), axis=1, dtype=np.uint32).astype(np.uint8)  # undo the left-neighbour difference
    if hashlib.sha256(out.tobytes()).hexdigest() != info["sha256"]:
        raise RuntimeError(f"texel fixture {name} is corrupt")
    return out


def reference_images() -> dict:
    return {name: reference_texels(name) for name in manifest()["images"]}
