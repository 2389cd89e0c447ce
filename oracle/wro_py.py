"""ctypes binding of the ORACLE (oracle/libwro.so) — test infrastructure, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
The POD structures at the boundary are the ones of include/wrt.h, mirrored in zig-weekend-raytracer_b200/abi.py
(loaded here by file path so that importing the oracle never loads the product library).
"""
from __future__ import annotations

import ctypes as C
import importlib.util
import os
import subprocess
from pathlib import Path

import numpy as np

ORACLE_DIR = Path(__file__).resolve().parent
ROOT = ORACLE_DIR.parent
# WRO_LIB=<file> selects another build of the same sources (bench.py: libwro_native.so, -O3 -march=native on the GPU box's host)
LIB_PATH = Path(os.environ["WRO_LIB"]) if os.environ.get("WRO_LIB") else ORACLE_DIR / "libwro.so"

_spec = importlib.util.spec_from_file_location("wrt_abi", ROOT / "zig-weekend-raytracer_b200" / "abi.py")
abi = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(abi)

RNG_REFERENCE = 0  # Xoshiro256++ / ziggurat restatement of the reference's std.Random use
RNG_COUNTER = 1    # the Philox counter stream shared with the device

SCENES = ["balls", "shrek_quads", "emissive", "cornell_box", "rtw_final", "earth", "synthetic"]


class ImageIn(C.Structure):
    _fields_ = [("name", C.c_char_p), ("width", C.c_uint32), ("height", C.c_uint32), ("num_components", C.c_uint32),
                ("data", C.c_void_p)]


class Flat(C.Structure):
    _fields_ = [("scene", abi.Scene), ("owner", C.c_void_p)]


class RenderStats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("rays", C.c_uint64), ("seconds", C.c_double), ("threads", C.c_uint32)]


def build():
    subprocess.check_call(["make", "-s", "-C", str(ORACLE_DIR), "libwro.so"])


def _load() -> C.CDLL:
    if not LIB_PATH.exists():
        build()
    lib = C.CDLL(str(LIB_PATH))
    vp = C.c_void_p
    lib.wro_scene_build.restype = vp
    lib.wro_scene_build.argtypes = [C.c_char_p, C.c_uint64, C.c_uint32, C.POINTER(ImageIn), C.c_uint32]
    lib.wro_scene_from_flat.restype = vp
    lib.wro_scene_from_flat.argtypes = [vp]
    lib.wro_scene_destroy.argtypes = [vp]
    lib.wro_scene_destroy.restype = None
    lib.wro_scene_set_no_cull.argtypes = [vp, C.c_int]
    lib.wro_scene_set_no_cull.restype = None
    lib.wro_scene_camera.argtypes = [vp, C.c_uint32, C.c_uint32, vp]
    lib.wro_scene_camera.restype = None
    lib.wro_scene_camera_desc.argtypes = [vp, vp]
    lib.wro_scene_camera_desc.restype = None
    lib.wro_scene_background.argtypes = [vp, vp]
    lib.wro_scene_background.restype = None
    lib.wro_scene_n_prims.argtypes = [vp]
    lib.wro_scene_n_prims.restype = C.c_uint32
    lib.wro_scene_flatten.argtypes = [vp, C.POINTER(Flat)]
    lib.wro_flat_free.argtypes = [C.POINTER(Flat)]
    lib.wro_flat_free.restype = None
    lib.wro_scene_prim_table.argtypes = [vp, vp, vp, vp]
    lib.wro_render.argtypes = [vp, vp, vp, C.c_int, C.c_uint32, vp, C.c_size_t,
                               C.POINTER(RenderStats)]
    lib.wro_primary_hits.argtypes = [vp, vp, vp, C.c_uint32, C.c_uint32, vp, vp]
    lib.wro_trace_rays.argtypes = [vp, vp, vp, C.c_uint64, C.c_double, vp, vp, vp, vp, vp, vp]
    lib.wro_light_pdf_values.argtypes = [vp, vp, vp, C.c_uint64, vp]
    lib.wro_sobol_pixel_samples.argtypes = [C.c_uint32, C.c_uint32, vp, vp, vp, C.c_uint64, vp, vp]
    lib.wro_sobol_dimension_samples.argtypes = [vp, vp, C.c_uint64, C.c_uint32, C.c_uint32, vp]
    lib.wro_sobol_get1d_sequence.argtypes = [C.c_uint32] * 7 + [C.c_uint32, vp]
    lib.wro_encode_color.argtypes = [vp, vp]
    lib.wro_encode_color.restype = None
    lib.wro_encode_image.argtypes = [vp, C.c_size_t, C.c_uint64, vp]
    lib.wro_encode_image.restype = None
    lib.wro_size_of_line.argtypes = [vp]
    lib.wro_size_of_line.restype = C.c_uint32
    lib.wro_size_of_digit.argtypes = [C.c_uint8]
    lib.wro_size_of_digit.restype = C.c_uint32
    lib.wro_counter_rng_bits.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
    lib.wro_counter_rng_bits.restype = C.c_uint64
    lib.wro_kat_splitmix64.argtypes = [C.c_uint64, C.c_uint32, vp]
    lib.wro_kat_splitmix64.restype = None
    lib.wro_kat_xoshiro256pp.argtypes = [vp, C.c_uint32, vp]
    lib.wro_kat_xoshiro256pp.restype = None
    lib.wro_kat_default_prng.argtypes = [C.c_uint64, C.c_uint32, vp, vp]
    lib.wro_kat_default_prng.restype = None
    lib.wro_murmur2_hash_u32_with_seed.argtypes = [C.c_uint32, C.c_uint32]
    lib.wro_murmur2_hash_u32_with_seed.restype = C.c_uint32
    lib.wro_math_cross.argtypes = [vp, vp, vp]
    lib.wro_math_cross.restype = None
    lib.wro_math_dot.argtypes = [vp, vp]
    lib.wro_math_dot.restype = C.c_double
    lib.wro_math_length.argtypes = [vp]
    lib.wro_math_length.restype = C.c_double
    lib.wro_math_normalize.argtypes = [vp, vp]
    lib.wro_math_normalize.restype = None
    return lib


lib = _load()


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class OracleScene:
    """A scene of the oracle's catalogue (or rebuilt from flat arrays) plus the reference semantics on it."""

    def __init__(self, name: str | None = None, seed: int = 1, n_prims: int = 0, images: dict | None = None,
                 flat: abi.Scene | None = None):
        self._keep = []
        if flat is not None:
            self._h = lib.wro_scene_from_flat(C.byref(flat))
            self.name = "from_flat"
        else:
            arr = None
            n = 0
            if images:
                n = len(images)
                arr = (ImageIn * n)()
                for i, (nm, img) in enumerate(images.items()):
                    img = np.ascontiguousarray(img, dtype=np.uint8)
                    assert img.ndim == 3
                    self._keep.append(img)
                    arr[i].name = nm.encode()
                    arr[i].height, arr[i].width, arr[i].num_components = img.shape
                    arr[i].data = img.ctypes.data
            self._h = lib.wro_scene_build(name.encode(), seed, n_prims, arr, n)
            self.name = name
        if not self._h:
            raise ValueError(f"oracle: cannot build scene {name!r}")
        self._flat = None

    def close(self):
        if self._flat is not None:
            lib.wro_flat_free(C.byref(self._flat))
            self._flat = None
        if self._h:
            lib.wro_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- description ----
    @property
    def n_prims(self) -> int:
        return lib.wro_scene_n_prims(self._h)

    def camera(self, width: int, height: int) -> abi.Camera:
        cam = abi.Camera()
        lib.wro_scene_camera(self._h, width, height, C.byref(cam))
        return cam

    def camera_desc(self) -> np.ndarray:
        out = np.zeros(12)
        lib.wro_scene_camera_desc(self._h, _ptr(out))
        return out

    def background(self) -> np.ndarray:
        out = np.zeros(3)
        lib.wro_scene_background(self._h, _ptr(out))
        return out

    def params(self, width: int, height: int, spp: int, depth: int, seed: int = 1, **kw) -> abi.Params:
        p = abi.Params(width=width, height=height, samples_per_pixel=spp, max_ray_bounce_depth=depth, seed=seed,
                       row_shard_index=0, row_shard_count=1)
        bg = self.background()
        for k in range(3):
            p.background_color[k] = bg[k]
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    def flatten(self) -> abi.Scene:
        """The wrt_scene a Zig shim would hand to wrt_upload_scene (valid while this object lives)."""
        if self._flat is None:
            f = Flat()
            rc = lib.wro_scene_flatten(self._h, C.byref(f))
            if rc != 0:
                raise RuntimeError("wro_scene_flatten failed")
            self._flat = f
        return self._flat.scene

    def prim_table(self):
        n = self.n_prims
        kinds = np.zeros(n, np.uint32)
        mats = np.zeros(n, np.uint32)
        centers = np.zeros((n, 3))
        lib.wro_scene_prim_table(self._h, _ptr(kinds), _ptr(mats), _ptr(centers))
        return kinds, mats, centers

    def set_no_cull(self, flag: bool):
        lib.wro_scene_set_no_cull(self._h, int(flag))

    # ---- reference semantics ----
    def render(self, cam: abi.Camera, params: abi.Params, rng_mode: int = RNG_COUNTER, threads: int | None = None,
               lanes: int = 4):
        cnt = params.row_shard_count or 1
        rows = 0 if params.row_shard_index >= params.height else (params.height - params.row_shard_index + cnt - 1) // cnt
        fb = np.zeros((rows, params.width, lanes), dtype=np.float64)
        st = RenderStats()
        rc = lib.wro_render(self._h, C.byref(cam), C.byref(params), rng_mode, threads or host_threads(), _ptr(fb), lanes * 8,
                            C.byref(st))
        if rc != 0:
            raise RuntimeError("wro_render failed")
        return fb, st

    def render_into(self, fb: np.ndarray, cam: abi.Camera, params: abi.Params, rng_mode: int = RNG_COUNTER,
                    threads: int | None = None):
        st = RenderStats()
        rc = lib.wro_render(self._h, C.byref(cam), C.byref(params), rng_mode, threads or host_threads(), _ptr(fb),
                            fb.shape[-1] * 8, C.byref(st))
        if rc != 0:
            raise RuntimeError("wro_render failed")
        return st

    def primary_hits(self, cam: abi.Camera, params: abi.Params, n_samples: int, threads: int | None = None):
        ids = np.zeros((params.height, params.width, n_samples), np.uint32)
        t = np.zeros((params.height, params.width, n_samples), np.float64)
        lib.wro_primary_hits(self._h, C.byref(cam), C.byref(params), n_samples, threads or host_threads(), _ptr(ids), _ptr(t))
        return ids, t

    def trace_rays(self, origins, directions, tmin: float = 1e-4):
        origins = np.ascontiguousarray(origins, np.float64)
        directions = np.ascontiguousarray(directions, np.float64)
        n = origins.shape[0]
        out = {
            "prim_id": np.zeros(n, np.uint32), "t": np.zeros(n, np.float64), "point": np.zeros((n, 3)),
            "normal": np.zeros((n, 3)), "uv": np.zeros((n, 2)), "front_face": np.zeros(n, np.uint32),
        }
        lib.wro_trace_rays(self._h, _ptr(origins), _ptr(directions), n, tmin, _ptr(out["prim_id"]), _ptr(out["t"]),
                           _ptr(out["point"]), _ptr(out["normal"]), _ptr(out["uv"]), _ptr(out["front_face"]))
        return out

    def light_pdf_values(self, origins, directions):
        origins = np.ascontiguousarray(origins, np.float64)
        directions = np.ascontiguousarray(directions, np.float64)
        out = np.zeros(origins.shape[0])
        rc = lib.wro_light_pdf_values(self._h, _ptr(origins), _ptr(directions), origins.shape[0], _ptr(out))
        if rc != 0:
            raise RuntimeError("scene has no lights")
        return out


def sobol_pixel_samples(width, height, cols, rows, sample_idx):
    cols = np.ascontiguousarray(cols, np.uint32)
    rows = np.ascontiguousarray(rows, np.uint32)
    sidx = np.ascontiguousarray(sample_idx, np.uint32)
    n = cols.shape[0]
    index = np.zeros(n, np.uint64)
    offsets = np.zeros((n, 2))
    lib.wro_sobol_pixel_samples(width, height, _ptr(cols), _ptr(rows), _ptr(sidx), n, _ptr(index), _ptr(offsets))
    return index, offsets


def sobol_dimension_samples(sobol_index, dimension, owen_fast: bool, seed: int):
    idx = np.ascontiguousarray(sobol_index, np.uint64)
    dim = np.ascontiguousarray(dimension, np.uint32)
    out = np.zeros(idx.shape[0], np.float32)
    lib.wro_sobol_dimension_samples(_ptr(idx), _ptr(dim), idx.shape[0], int(owen_fast), seed, _ptr(out))
    return out


def sobol_get1d_sequence(width, height, col, row, sample_idx, owen_fast: bool, seed: int, n: int):
    out = np.zeros(n)
    lib.wro_sobol_get1d_sequence(width, height, col, row, sample_idx, int(owen_fast), seed, n, _ptr(out))
    return out


def encode_color(rgb) -> np.ndarray:
    rgb = np.ascontiguousarray(rgb, np.float64)
    out = np.zeros(3, np.uint8)
    lib.wro_encode_color(_ptr(rgb), _ptr(out))
    return out


def encode_image(fb: np.ndarray) -> np.ndarray:
    fb = np.ascontiguousarray(fb, np.float64)
    n = fb.shape[0] * fb.shape[1]
    out = np.zeros((fb.shape[0], fb.shape[1], 3), np.uint8)
    lib.wro_encode_image(_ptr(fb), fb.shape[2] * 8, n, _ptr(out))
    return out


def procedural_image(name: str, width: int, height: int) -> np.ndarray:
    """Deterministic RGB test image used where a reference asset cannot travel (no /root/reference on the GPU box)."""
    y, x = np.mgrid[0:height, 0:width]
    seed = sum(name.encode())
    r = (x * 255 // max(width - 1, 1)) ^ ((y * 7 + seed) & 0xFF)
    g = (y * 255 // max(height - 1, 1)) ^ ((x * 3 + seed * 5) & 0xFF)
    b = ((x // 8 + y // 8) % 2) * 200 + ((x * y + seed) % 56)
    return np.stack([r, g, b], axis=-1).astype(np.uint8)
