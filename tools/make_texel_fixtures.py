#!/usr/bin/env python3
"""Decode the reference's assets with the reference's OWN decoder and commit the texels (run in the build container).

assets/earth.png, wap.jpg, me.jpg are loaded by zstbi.Image.loadFromFile (src/image.zig:12-17) = the vendored stb_image v2.28.
oracle/_ref/libstbi.so is that header compiled where it lies under /root/reference (oracle/Makefile target `ref`), so the
bytes written here are exactly what Image.getPixel (image.zig:23-36) reads in the reference.  /root/reference does not exist
on the GPU box: the decoded texels travel as  zig-weekend-raytracer_b200/data/texels/<name>.rgb8.xz  (left-neighbour
difference per row, then xz — lossless) with manifest.json (shape, sha256 of the raw decode, decimation).

  earth.png  2048 x 1024 x 3   full resolution (config C4's texture)
  wap.jpg     300 x  292 x 3   full resolution
  me.jpg     2316 x 3088 x 3   every 4th texel of every 4th row (579 x 772): the full decode is 21 MB; its sha256 is kept
                               and checked against the decoder whenever /root/reference is mounted
"""
import ctypes as C
import hashlib
import json
import lzma
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "zig-weekend-raytracer_b200" / "data" / "texels"
ASSETS = Path("/root/reference/assets")
DECIMATE = {"earth.png": 1, "wap.jpg": 1, "me.jpg": 4}


def decode(path: Path) -> np.ndarray:
    lib = C.CDLL(str(ROOT / "oracle" / "_ref" / "libstbi.so"))
    lib.wro_stbi_load.restype = C.POINTER(C.c_ubyte)
    lib.wro_stbi_load.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.wro_stbi_free.argtypes = [C.POINTER(C.c_ubyte)]
    w, h, c = C.c_int(), C.c_int(), C.c_int()
    ptr = lib.wro_stbi_load(str(path).encode(), C.byref(w), C.byref(h), C.byref(c))
    if not ptr:
        raise RuntimeError(f"stb_image cannot decode {path}")
    arr = np.ctypeslib.as_array(ptr, shape=(h.value, w.value, c.value)).copy()
    lib.wro_stbi_free(ptr)
    return arr


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    manifest = {"decoder": "stb_image v2.28 as vendored by the reference (libs/zstbi/libs/stbi/stb_image.h), compiled where it lies",
                "encoding": "uint8 [H][W][C]; per row: byte - left neighbour (mod 256), then xz", "images": {}}
    for name, step in DECIMATE.items():
        full = decode(ASSETS / name)
        kept = np.ascontiguousarray(full[::step, ::step])
        d = kept.astype(np.int16)
        d[:, 1:] -= kept[:, :-1].astype(np.int16)
        blob = lzma.compress((d & 255).astype(np.uint8).tobytes(), preset=9)
        (OUT / f"{name}.rgb8.xz").write_bytes(blob)
        manifest["images"][name] = {"full_shape": list(full.shape), "full_sha256": hashlib.sha256(full.tobytes()).hexdigest(),
                                    "decimation": step, "shape": list(kept.shape),
                                    "sha256": hashlib.sha256(kept.tobytes()).hexdigest()}
        print(name, full.shape, "->", kept.shape, len(blob), "bytes")
    (OUT / "manifest.json").write_text(json.dumps(manifest, indent=1))


if __name__ == "__main__":
    main()
