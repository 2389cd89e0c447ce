"""In-tree build of the native pieces.

  libwrt.so   the product: CUDA kernels + C ABI (csrc/), nvcc, sm_100a only

It is built next to its sources so that it travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
DATA = PKG / "data"
LIBWRT = PKG / "libwrt.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",          # binary64 ops stay unfused in source order; culling uses explicit fma()
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-Xcompiler", "-pthread",
    "-Xcompiler", "-ffp-contract=off",  # host side of the same rule (the host and device tree builders must agree bit for bit)
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the CUDA back end cannot be built (there is no CPU fallback)")


def _newer(target: Path, sources) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(Path(s).stat().st_mtime <= t for s in sources)


def build_wrt(force: bool = False, verbose: bool = False, extra_flags=()) -> Path:
    """Each .cu is compiled to its own object (in parallel, only when it or a header changed), then linked."""
    from concurrent.futures import ThreadPoolExecutor
    cu = [CSRC / "wrt_api.cu", CSRC / "wrt_kernels.cu", CSRC / "wrt_program.cu", CSRC / "wrt_multi.cu", CSRC / "wrt_build.cu"]
    hdrs = [CSRC / "wrt_device.cuh", CSRC / "wrt_treebuild.cuh", CSRC / "wrt_kernels.h", CSRC / "wrt_program.h", CSRC / "wrt_ctx.h", ROOT / "include" / "wrt.h",
            Path(__file__)]
    blob_deps = [CSRC / "wrt_sobol_blob.c", DATA / "sobol_tables.bin", Path(__file__)]
    objs = [c.with_suffix(".o") for c in cu]
    blob_o = CSRC / "wrt_sobol_blob.o"
    if extra_flags:
        force = True
    if not force and _newer(LIBWRT, cu + hdrs + blob_deps):
        return LIBWRT
    cc = os.environ.get("CC", "gcc")
    if force or not _newer(blob_o, blob_deps):
        subprocess.check_call([cc, "-c", "-fPIC", "-O2", f'-DWRT_SOBOL_BLOB="{DATA / "sobol_tables.bin"}"',
                               str(CSRC / "wrt_sobol_blob.c"), "-o", str(blob_o)])

    def compile_one(pair):
        src, obj = pair
        if not force and _newer(obj, [src] + hdrs):
            return
        cmd = [_nvcc(), *NVCC_FLAGS, *extra_flags, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)

    with ThreadPoolExecutor(max_workers=len(cu)) as pool:
        list(pool.map(compile_one, zip(cu, objs)))
    subprocess.check_call([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(LIBWRT), *map(str, objs),
                           str(blob_o), "-ldl", "-Xcompiler", "-pthread"])
    return LIBWRT


HOST = PKG / "host"
LIBWRTH = PKG / "libwrth.so"
CLI = PKG / "weekend-raytracer"
HOST_FLAGS = ["-O2", "-g", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fvisibility=hidden", "-Wall", "-Wextra", "-pthread"]


def build_host(force: bool = False) -> Path:
    """Host-side mirror of the reference interface (scene construction, flattening, Renderer.render, PPM writer) as
    libwrth.so, plus the `weekend-raytracer` command line driver.  Both link libwrt.so through $ORIGIN."""
    lib_srcs = [HOST / n for n in ("wrh_entity.cpp", "wrh_scenes.cpp", "wrh_render.cpp", "wrh_writer.cpp", "wrh_image.cpp", "wrh_capi.cpp")]
    hdrs = [HOST / n for n in ("wrh_math.hpp", "wrh_rng.hpp", "wrh_scene.hpp", "wrh_writer.hpp")] + [ROOT / "include" / "wrt.h"]
    cxx = os.environ.get("CXX", "g++")
    link = [f"-L{PKG}", "-lwrt", "-Wl,-rpath,$ORIGIN"]
    # JPEG / PNG decode = the reference's vendored stb_image.h, compiled where it lies in the reference checkout (never copied)
    stbi_flags = []
    for cand in (os.environ.get("WRT_STBI_INCLUDE"), "/root/reference/libs/zstbi/libs/stbi"):
        if cand and (Path(cand) / "stb_image.h").exists():
            stbi_flags = ["-DWRH_HAVE_STBI", f"-I{cand}"]
            break
    if not stbi_flags and LIBWRTH.exists() and CLI.exists() and not force:
        return LIBWRTH  # no header here (e.g. the GPU box): keep the artefacts built where the reference checkout was
    if force or not _newer(LIBWRTH, lib_srcs + hdrs + [LIBWRT, Path(__file__)]):
        subprocess.check_call([cxx, *HOST_FLAGS, *stbi_flags, "-shared", "-o", str(LIBWRTH), *map(str, lib_srcs), *link])
    if force or not _newer(CLI, lib_srcs + hdrs + [HOST / "main.cpp", LIBWRT, Path(__file__)]):
        subprocess.check_call([cxx, *HOST_FLAGS, *stbi_flags, "-o", str(CLI), str(HOST / "main.cpp"),
                               *map(str, lib_srcs[:-1]), *link])
    return LIBWRTH


if __name__ == "__main__":
    force = "--force" in sys.argv
    print(build_wrt(force=force, verbose="-v" in sys.argv))
    print(build_host(force=force))
