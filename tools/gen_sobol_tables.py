#!/usr/bin/env python3
"""Extract the Sobol / van-der-Corput tables used by the reference sampler into a binary blob.

The reference keeps three constant tables in src/math/sobolmatrices.zig (adapted there from
PBRT-v4, Joe & Kuo direction numbers):
  SobolMatrices32      [1024*52] u32   (sobolmatrices.zig:42)
  VdCSobolMatrices     [25][52]  u64   (sobolmatrices.zig:8926, jagged rows zero-padded to 52)
  VdCSobolMatricesInv  [26][52]  u64   (sobolmatrices.zig:9052, jagged rows zero-padded to 52)
They are data, not code: this script parses the literals mechanically (no retyping) and writes

  magic "WRTSOBL1" | u32 n_dims | u32 matrix_size | u32 n_vdc | u32 n_vdc_inv
  | u32[n_dims*matrix_size] | u64[n_vdc*matrix_size] | u64[n_vdc_inv*matrix_size]

to zig-weekend-raytracer_b200/data/sobol_tables.bin (little endian).  Both the oracle and the CUDA
library embed that blob with .incbin.  Run once in the build container (needs /root/reference).
"""
import re
import struct
import sys
from pathlib import Path

SRC = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/math/sobolmatrices.zig")
OUT = Path(__file__).resolve().parent.parent / "zig-weekend-raytracer_b200" / "data" / "sobol_tables.bin"


def strip_comments(text: str) -> str:
    return re.sub(r"//[^\n]*", "", text)


def main() -> None:
    text = strip_comments(SRC.read_text())
    n_dims = int(re.search(r"NSobolDimensions\s*=\s*(\d+)", text).group(1))
    msize = int(re.search(r"SobolMatrixSize\s*=\s*(\d+)", text).group(1))

    m = re.search(r"SobolMatrices32\s*=\s*\[[^\]]*\]u32\{(.*?)\};", text, re.S)
    sob = [int(x, 16) for x in re.findall(r"0x[0-9a-fA-F]+", m.group(1))]
    assert len(sob) == n_dims * msize, (len(sob), n_dims * msize)

    def jagged(name: str):
        mm = re.search(name + r"\s*=\s*\[_\]\[SobolMatrixSize\]u64\{(.*?)\}\)\};", text, re.S)
        body = mm.group(1) + "})"
        rows = re.findall(r"pad\(SobolMatrixSize,\s*\[_\]u64\{(.*?)\}\)", body, re.S)
        out = []
        for r in rows:
            vals = [int(x, 16) for x in re.findall(r"0x[0-9a-fA-F]+", r)]
            assert len(vals) <= msize
            out.append(vals + [0] * (msize - len(vals)))
        return out

    vdc = jagged(r"VdCSobolMatrices")
    vdc_inv = jagged(r"VdCSobolMatricesInv")
    assert len(vdc) == 25 and len(vdc_inv) == 26, (len(vdc), len(vdc_inv))

    blob = b"WRTSOBL1" + struct.pack("<4I", n_dims, msize, len(vdc), len(vdc_inv))
    blob += struct.pack("<%dI" % len(sob), *sob)
    for rows in (vdc, vdc_inv):
        for r in rows:
            blob += struct.pack("<%dQ" % msize, *r)
    OUT.parent.mkdir(parents=True, exist_ok=True)
    OUT.write_bytes(blob)
    print(f"wrote {OUT} ({len(blob)} bytes): dims={n_dims} size={msize} vdc={len(vdc)} inv={len(vdc_inv)}")


if __name__ == "__main__":
    main()
