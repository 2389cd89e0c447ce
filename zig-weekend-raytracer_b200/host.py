"""ctypes binding of libwrth.so — the host-side mirror of the reference interface (scene.zig / camera.zig /
render.zig / writer.zig restated in C++ under host/), i.e. the product's own scene construction, flattening,
Renderer.render call and PPM writer.  Used by bench.py and the tests; the CLI binary `weekend-raytracer` links the
same code."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from .abi import Camera, Params, Scene

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libwrth.so"
CLI_PATH = PKG_DIR / "weekend-raytracer"

SCENES = ["balls", "shrek_quads", "emissive", "cornell_box", "rtw_final", "earth", "synthetic"]


class _ImageIn(C.Structure):
    _fields_ = [("name", C.c_char_p), ("width", C.c_uint32), ("height", C.c_uint32), ("num_components", C.c_uint32),
                ("data", C.c_void_p)]


def _load() -> C.CDLL:
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} is missing: build it with `python {PKG_DIR / 'build.py'}`")
    lib = C.CDLL(str(LIB_PATH))
    vp = C.c_void_p
    lib.wrh_last_error.restype = C.c_char_p
    lib.wrh_scene_load.restype = vp
    lib.wrh_scene_load.argtypes = [C.c_char_p, C.c_uint64, C.c_uint32, C.c_char_p, C.POINTER(_ImageIn), C.c_uint32]
    lib.wrh_scene_free.argtypes = [vp]
    lib.wrh_scene_free.restype = None
    lib.wrh_scene_flat.argtypes = [vp]
    lib.wrh_scene_flat.restype = C.POINTER(Scene)
    lib.wrh_scene_input_bytes.argtypes = [vp]
    lib.wrh_scene_input_bytes.restype = C.c_uint64
    lib.wrh_scene_camera.argtypes = [vp, C.c_uint32, C.c_uint32, C.POINTER(Camera)]
    lib.wrh_scene_camera.restype = None
    lib.wrh_scene_background.argtypes = [vp, vp]
    lib.wrh_scene_background.restype = None
    lib.wrh_scene_draw.argtypes = [vp, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, vp, vp]
    lib.wrh_scene_draw.restype = C.c_int
    lib.wrh_write_ppm.argtypes = [C.c_char_p, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
    lib.wrh_write_ppm.restype = C.c_longlong
    lib.wrh_write_ppm_rgb8.argtypes = [C.c_char_p, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
    lib.wrh_write_ppm_rgb8.restype = C.c_longlong
    lib.wrh_encode_color.argtypes = [vp, vp]
    lib.wrh_encode_color.restype = None
    lib.wrh_size_of_line.argtypes = [vp]
    lib.wrh_size_of_line.restype = C.c_uint32
    lib.wrh_size_of_digit.argtypes = [C.c_uint8]
    lib.wrh_size_of_digit.restype = C.c_uint32
    return lib


lib = _load()


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class HostScene:
    """loadScene(scene_type, ctx) of the reference (scene.zig:26-34) on the host mirror."""

    def __init__(self, name: str, seed: int = 1, synthetic_prims: int = 0, images: dict | None = None,
                 asset_dir: str | None = None):
        self._keep = []
        arr, n = None, 0
        if images:
            n = len(images)
            arr = (_ImageIn * n)()
            for i, (nm, img) in enumerate(images.items()):
                img = np.ascontiguousarray(img, dtype=np.uint8)
                self._keep.append(img)
                arr[i].name = nm.encode()
                arr[i].height, arr[i].width, arr[i].num_components = img.shape
                arr[i].data = img.ctypes.data
        self._h = lib.wrh_scene_load(name.encode(), seed, synthetic_prims, asset_dir.encode() if asset_dir else None, arr, n)
        if not self._h:
            raise ValueError((lib.wrh_last_error() or b"").decode())
        self.name = name

    def close(self):
        if getattr(self, "_h", None):
            lib.wrh_scene_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def flat(self) -> Scene:
        """The wrt_scene view Renderer.render uploads (valid while this object lives)."""
        return lib.wrh_scene_flat(self._h).contents

    def input_bytes(self) -> int:
        return int(lib.wrh_scene_input_bytes(self._h))

    def camera(self, width: int, height: int) -> Camera:
        cam = Camera()
        lib.wrh_scene_camera(self._h, width, height, C.byref(cam))
        return cam

    def background(self) -> np.ndarray:
        out = np.zeros(3)
        lib.wrh_scene_background(self._h, _ptr(out))
        return out

    def params(self, width: int, height: int, spp: int, depth: int, seed: int = 1, **kw) -> Params:
        p = Params(width=width, height=height, samples_per_pixel=spp, max_ray_bounce_depth=depth, seed=seed,
                   row_shard_index=0, row_shard_count=1)
        bg = self.background()
        for k in range(3):
            p.background_color[k] = bg[k]
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    def draw(self, width: int, height: int, spp: int, depth: int, seed: int = 1, cull_mode: int = 0, device: int = 0):
        """Scene.draw -> Renderer.render (scene.zig:57-61): returns (framebuffer[h, w, 4], stats dict)."""
        fb = np.zeros((height, width, 4), dtype=np.float64)
        st = np.zeros(5)
        rc = lib.wrh_scene_draw(self._h, device, width, height, spp, depth, seed, cull_mode, _ptr(fb), _ptr(st))
        if rc != 0:
            raise RuntimeError((lib.wrh_last_error() or b"").decode())
        return fb, {"paths": int(st[0]), "rays": int(st[1]), "render_ms": st[2], "kernel_ms": st[3], "upload_ms": st[4]}


def write_ppm(path: str, fb: np.ndarray, threads: int = 8, truncate: bool = False) -> int:
    """WriterPPM.write (writer.zig:16-51) on a (rows, cols, lanes) f64 frame."""
    fb = np.ascontiguousarray(fb, dtype=np.float64)
    n = lib.wrh_write_ppm(str(path).encode(), _ptr(fb), fb.shape[2], fb.shape[1], fb.shape[0], threads, int(truncate))
    if n < 0:
        raise RuntimeError((lib.wrh_last_error() or b"").decode())
    return int(n)


def write_ppm_rgb8(path: str, rgb: np.ndarray, threads: int = 8, truncate: bool = False) -> int:
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    n = lib.wrh_write_ppm_rgb8(str(path).encode(), _ptr(rgb), rgb.shape[1], rgb.shape[0], threads, int(truncate))
    if n < 0:
        raise RuntimeError((lib.wrh_last_error() or b"").decode())
    return int(n)


def encode_color(rgb) -> np.ndarray:
    rgb = np.ascontiguousarray(rgb, np.float64)
    out = np.zeros(3, np.uint8)
    lib.wrh_encode_color(_ptr(rgb), _ptr(out))
    return out
