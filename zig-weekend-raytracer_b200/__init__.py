"""zig-weekend-raytracer_b200 — B200-native render back end for j-helland/zig-weekend-raytracer.

The product is ``libwrt.so`` (hand-written sm_100a CUDA behind the C ABI of ``include/wrt.h``).  This module is a
thin ctypes binding of that ABI — the same calls a Zig ``extern fn`` block makes — used by the
tests, ``bench.py`` and ``__graft_entry__``.  There is no CPU fallback: if the shared library is missing the import
fails, and if no CUDA device is present ``Context()`` raises.

The package directory name contains a hyphen; import it with ``importlib.import_module("zig-weekend-raytracer_b200")``.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
import os as _os
# WRT_LIB=<file> loads another build of the same sources (development: kernel variants side by side)
LIB_PATH = Path(_os.environ["WRT_LIB"]) if _os.environ.get("WRT_LIB") else PKG_DIR / "libwrt.so"

from .abi import *  # noqa: F401,F403  (structures and constants of include/wrt.h)
from .abi import Camera, Params, Scene, Stats


# every symbol include/wrt.h declares
ABI_SYMBOLS = [
    "wrt_create", "wrt_destroy", "wrt_last_error", "wrt_abi_version", "wrt_upload_scene", "wrt_render",
    "wrt_render_device", "wrt_encode_rgb8", "wrt_primary_hits", "wrt_trace_rays", "wrt_sobol_pixel_samples",
    "wrt_sobol_dimension_samples", "wrt_get_stats", "wrt_fp64_issue_peak", "wrt_fp32_issue_peak", "wrt_format_ppm", "wrt_check_scene", "wrt_build_trees",
    "wrt_group_create", "wrt_group_destroy", "wrt_group_last_error", "wrt_group_size", "wrt_group_ctx", "wrt_group_upload_scene",
    "wrt_group_render", "wrt_group_encode_rgb8", "wrt_group_get_stats", "wrt_comm_unique_id", "wrt_comm_init", "wrt_render_sharded",
]


class WrtError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"wrt error {code}: {message}")
        self.code = code
        self.message = message


def _load() -> C.CDLL:
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} is missing: build it with `python {PKG_DIR / 'build.py'}` "
                          "(the CUDA back end has no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    vp = C.c_void_p
    lib.wrt_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.wrt_create.restype = C.c_int
    lib.wrt_destroy.argtypes = [vp]
    lib.wrt_destroy.restype = None
    lib.wrt_last_error.argtypes = [vp]
    lib.wrt_last_error.restype = C.c_char_p
    lib.wrt_abi_version.restype = C.c_uint32
    lib.wrt_upload_scene.argtypes = [vp, vp]
    lib.wrt_render.argtypes = [vp, vp, vp, vp, C.c_size_t]
    lib.wrt_render_device.argtypes = [vp, vp, vp, vp, C.c_size_t]
    lib.wrt_encode_rgb8.argtypes = [vp, vp]
    lib.wrt_primary_hits.argtypes = [vp, vp, vp, C.c_uint32, vp, vp]
    lib.wrt_trace_rays.argtypes = [vp, vp, vp, C.c_uint64, C.c_double, C.c_uint32, vp, vp, vp, vp, vp, vp]
    lib.wrt_sobol_pixel_samples.argtypes = [vp, C.c_uint32, C.c_uint32, vp, vp, vp, C.c_uint64, vp, vp]
    lib.wrt_sobol_dimension_samples.argtypes = [vp, vp, vp, C.c_uint64, C.c_uint32, C.c_uint32, vp]
    lib.wrt_get_stats.argtypes = [vp, vp]
    lib.wrt_fp64_issue_peak.argtypes = [vp, vp]
    lib.wrt_fp32_issue_peak.argtypes = [vp, vp]
    lib.wrt_format_ppm.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp, C.c_uint64, vp]
    lib.wrt_check_scene.argtypes = [vp, vp, C.c_char_p, C.c_size_t]
    lib.wrt_build_trees.argtypes = [vp, C.c_int, vp, vp, C.c_size_t, vp, C.c_size_t, C.c_char_p, C.c_size_t]
    lib.wrt_group_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
    lib.wrt_group_destroy.argtypes = [vp]
    lib.wrt_group_destroy.restype = None
    lib.wrt_group_last_error.argtypes = [vp]
    lib.wrt_group_last_error.restype = C.c_char_p
    lib.wrt_group_size.argtypes = [vp]
    lib.wrt_group_ctx.argtypes = [vp, C.c_int]
    lib.wrt_group_ctx.restype = vp
    lib.wrt_group_upload_scene.argtypes = [vp, vp]
    lib.wrt_group_render.argtypes = [vp, vp, vp, vp, C.c_size_t]
    lib.wrt_group_encode_rgb8.argtypes = [vp, vp]
    lib.wrt_group_get_stats.argtypes = [vp, vp]
    lib.wrt_comm_unique_id.argtypes = [vp]
    lib.wrt_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.wrt_render_sharded.argtypes = [vp, vp, vp, vp, C.c_size_t]
    for name in ABI_SYMBOLS:
        if name not in ("wrt_destroy", "wrt_last_error", "wrt_abi_version", "wrt_group_destroy", "wrt_group_last_error", "wrt_group_ctx"):
            getattr(lib, name).restype = C.c_int
    return lib


lib = _load()


def check_scene(scene) -> "SceneInfo":
    """Compile `scene` on the host (no device needed) and check the structure of the result; raises WrtError."""
    info = SceneInfo()
    err = C.create_string_buffer(512)
    rc = lib.wrt_check_scene(C.addressof(scene), C.addressof(info), err, len(err))
    if rc != 0:
        raise WrtError(rc, err.value.decode(errors="replace"))
    return info


def build_trees(scene, device: int = -1, records: bool = True):
    """The ordered traversal's tree build on its own (wrt_build_trees): on CUDA device `device`, or on the host threads
    when device < 0.  Returns (TreeInfo, child-pair records as uint8[n, 64], four-wide records as uint8[n, 128]); the two
    builders write the same bytes.  Raises WrtError (no CPU fallback for device >= 0)."""
    import numpy as np
    info = TreeInfo()
    err = C.create_string_buffer(512)
    rc = lib.wrt_build_trees(C.addressof(scene), int(device), C.addressof(info), None, 0, None, 0, err, len(err))
    if rc != 0:
        raise WrtError(rc, err.value.decode(errors="replace"))
    if not records:
        return info, None, None
    r2 = np.zeros((info.n_records2, 64), np.uint8)
    r4 = np.zeros((info.n_records4, 128), np.uint8)
    rc = lib.wrt_build_trees(C.addressof(scene), int(device), C.addressof(info), r2.ctypes.data_as(C.c_void_p), r2.nbytes,
                             r4.ctypes.data_as(C.c_void_p), r4.nbytes, err, len(err))
    if rc != 0:
        raise WrtError(rc, err.value.decode(errors="replace"))
    return info, r2, r4


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """One wrt_ctx bound to one CUDA device (include/wrt.h)."""

    def __init__(self, device: int = 0, _borrowed: int | None = None):
        self._owned = _borrowed is None
        if _borrowed is not None:  # a member of a Group: the group owns it
            self._h = C.c_void_p(_borrowed)
            self.device = device
            return
        h = C.c_void_p()
        rc = lib.wrt_create(device, C.byref(h))
        if rc != 0:
            raise WrtError(rc, (lib.wrt_last_error(None) or b"").decode())
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            if self._owned:
                lib.wrt_destroy(self._h)
            self._h = None

    # ---- one process per device (include/wrt.h, Multi-GPU (2)) ----
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * WRT_COMM_ID_BYTES)()
        rc = lib.wrt_comm_unique_id(buf)
        if rc != 0:
            raise WrtError(rc, (lib.wrt_group_last_error(None) or b"").decode())
        return bytes(buf)

    def comm_init(self, unique_id: bytes, rank: int, n_ranks: int):
        assert len(unique_id) == WRT_COMM_ID_BYTES
        buf = (C.c_uint8 * WRT_COMM_ID_BYTES).from_buffer_copy(unique_id)
        self._check(lib.wrt_comm_init(self._h, buf, rank, n_ranks))

    def render_sharded(self, cam: Camera, params: Params, out: np.ndarray | None = None, lanes: int = 4):
        """Collective: every rank renders its shard, rank 0 receives the frame.  `out` (rank 0): host array
        (height, width, lanes) f64, or None to leave the assembled frame on the device."""
        if out is not None:
            assert out.dtype == np.float64 and out.flags.c_contiguous and out.shape == (params.height, params.width, lanes)
        self._check(lib.wrt_render_sharded(self._h, C.byref(cam), C.byref(params), _ptr(out), lanes * 8))
        return out

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc != 0:
            raise WrtError(rc, (lib.wrt_last_error(self._h) or b"").decode())

    def upload_scene(self, scene: Scene):
        self._check(lib.wrt_upload_scene(self._h, C.byref(scene)))

    def stats(self) -> Stats:
        s = Stats()
        self._check(lib.wrt_get_stats(self._h, C.byref(s)))
        return s

    def fp64_issue_peak(self) -> float:
        """Measured binary64 FMA issue rate of the device, thread-level FMAs per second."""
        out = C.c_double()
        self._check(lib.wrt_fp64_issue_peak(self._h, C.byref(out)))
        return out.value

    def format_ppm(self, width: int, height: int, rgb8: np.ndarray | None = None) -> tuple[np.ndarray, int]:
        """The reference's PPM file image (header + packed pixel lines + NUL tail) of the last rendered frame, or of
        `rgb8` (H x W x 3 uint8), formatted on the device.  Returns (file bytes, content length)."""
        header = f"P3\n{width} {height}\n255\n".encode()
        out = np.zeros(len(header) + 12 * width * height, np.uint8)
        n = C.c_uint64()
        src = None
        if rgb8 is not None:
            rgb8 = np.ascontiguousarray(rgb8, dtype=np.uint8)
            assert rgb8.size == width * height * 3
            src = _ptr(rgb8)
        self._check(lib.wrt_format_ppm(self._h, src, width, height, _ptr(out), out.size, C.byref(n)))
        return out, int(n.value)

    def fp32_issue_peak(self) -> float:
        """Measured binary32 FMA issue rate of the device, thread-level FMAs per second."""
        out = C.c_double()
        self._check(lib.wrt_fp32_issue_peak(self._h, C.byref(out)))
        return out.value

    @staticmethod
    def local_rows(params: Params) -> int:
        cnt = params.row_shard_count or 1
        if params.row_shard_index >= params.height:
            return 0
        return (params.height - params.row_shard_index + cnt - 1) // cnt

    def render(self, cam: Camera, params: Params, lanes: int = 4, out: np.ndarray | None = None) -> np.ndarray:
        """Renderer.render into a host framebuffer of shape (rows, width, lanes) f64 (lanes = Vec3 width)."""
        rows = self.local_rows(params)
        if out is None:
            out = np.zeros((rows, params.width, lanes), dtype=np.float64)
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.shape == (rows, params.width, lanes)
        self._check(lib.wrt_render(self._h, C.byref(cam), C.byref(params), _ptr(out), lanes * 8))
        return out

    def render_device(self, cam: Camera, params: Params, device_ptr: int, stride_bytes: int):
        self._check(lib.wrt_render_device(self._h, C.byref(cam), C.byref(params), C.c_void_p(device_ptr), stride_bytes))

    def encode_rgb8(self, rows: int, width: int) -> np.ndarray:
        out = np.zeros((rows, width, 3), dtype=np.uint8)
        self._check(lib.wrt_encode_rgb8(self._h, _ptr(out)))
        return out

    def primary_hits(self, cam: Camera, params: Params, n_samples: int):
        ids = np.zeros((params.height, params.width, n_samples), dtype=np.uint32)
        t = np.zeros((params.height, params.width, n_samples), dtype=np.float64)
        self._check(lib.wrt_primary_hits(self._h, C.byref(cam), C.byref(params), n_samples, _ptr(ids), _ptr(t)))
        return ids, t

    def trace_rays(self, origins: np.ndarray, directions: np.ndarray, tmin: float = 1e-4, cull_mode: int = WRT_CULL_AUTO):
        origins = np.ascontiguousarray(origins, dtype=np.float64)
        directions = np.ascontiguousarray(directions, dtype=np.float64)
        n = origins.shape[0]
        out = {
            "prim_id": np.zeros(n, np.uint32), "t": np.zeros(n, np.float64), "point": np.zeros((n, 3), np.float64),
            "normal": np.zeros((n, 3), np.float64), "uv": np.zeros((n, 2), np.float64), "front_face": np.zeros(n, np.uint32),
        }
        self._check(lib.wrt_trace_rays(self._h, _ptr(origins), _ptr(directions), n, tmin, cull_mode, _ptr(out["prim_id"]),
                                       _ptr(out["t"]), _ptr(out["point"]), _ptr(out["normal"]), _ptr(out["uv"]),
                                       _ptr(out["front_face"])))
        return out

    def sobol_pixel_samples(self, width: int, height: int, cols, rows, sample_idx):
        cols = np.ascontiguousarray(cols, np.uint32)
        rows = np.ascontiguousarray(rows, np.uint32)
        sidx = np.ascontiguousarray(sample_idx, np.uint32)
        n = cols.shape[0]
        index = np.zeros(n, np.uint64)
        offsets = np.zeros((n, 2), np.float64)
        self._check(lib.wrt_sobol_pixel_samples(self._h, width, height, _ptr(cols), _ptr(rows), _ptr(sidx), n, _ptr(index),
                                                _ptr(offsets)))
        return index, offsets

    def sobol_dimension_samples(self, sobol_index, dimension, owen_fast: bool, seed: int):
        idx = np.ascontiguousarray(sobol_index, np.uint64)
        dim = np.ascontiguousarray(dimension, np.uint32)
        out = np.zeros(idx.shape[0], np.float32)
        self._check(lib.wrt_sobol_dimension_samples(self._h, _ptr(idx), _ptr(dim), idx.shape[0], int(owen_fast), seed, _ptr(out)))
        return out


class Group:
    """A wrt_group: one process driving several devices (include/wrt.h, Multi-GPU (1))."""

    def __init__(self, devices):
        devices = list(devices)
        arr = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        rc = lib.wrt_group_create(arr, len(devices), C.byref(h))
        if rc != 0:
            raise WrtError(rc, (lib.wrt_group_last_error(None) or b"").decode())
        self._h = h
        self.devices = devices

    def close(self):
        if getattr(self, "_h", None):
            lib.wrt_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc != 0:
            raise WrtError(rc, (lib.wrt_group_last_error(self._h) or b"").decode())

    def __len__(self):
        return lib.wrt_group_size(self._h)

    def member(self, i: int) -> Context:
        return Context(self.devices[i], _borrowed=lib.wrt_group_ctx(self._h, i))

    def upload_scene(self, scene: Scene):
        self._check(lib.wrt_group_upload_scene(self._h, C.byref(scene)))

    def render(self, cam: Camera, params: Params, lanes: int = 4, out: np.ndarray | None = None, to_host: bool = True):
        if out is None and to_host:
            out = np.zeros((params.height, params.width, lanes), dtype=np.float64)
        if out is not None:
            assert out.dtype == np.float64 and out.flags.c_contiguous and out.shape == (params.height, params.width, lanes)
        self._check(lib.wrt_group_render(self._h, C.byref(cam), C.byref(params), _ptr(out), lanes * 8))
        return out

    def encode_rgb8(self, height: int, width: int) -> np.ndarray:
        out = np.zeros((height, width, 3), dtype=np.uint8)
        self._check(lib.wrt_group_encode_rgb8(self._h, _ptr(out)))
        return out

    def stats(self) -> Stats:
        s = Stats()
        self._check(lib.wrt_group_get_stats(self._h, C.byref(s)))
        return s
