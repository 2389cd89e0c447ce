"""Shared fixtures.  `-m "not gpu"` runs on the CPU-only build container; `-m gpu` needs a B200."""
from __future__ import annotations

import importlib
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def wro():
    """The oracle binding (builds oracle/libwro.so on demand)."""
    import wro_py
    return wro_py


@pytest.fixture(scope="session")
def wrt():
    """The product binding; builds libwrt.so with nvcc when it is missing (cross-compiles without a GPU)."""
    pkg = ROOT / "zig-weekend-raytracer_b200"
    if not (pkg / "libwrt.so").exists():
        spec = importlib.util.spec_from_file_location("wrt_build", pkg / "build.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build_wrt()
    return importlib.import_module("zig-weekend-raytracer_b200")


@pytest.fixture(scope="session")
def ctx(wrt):
    c = wrt.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def images():
    """Texel bytes for the image-textured scenes: the reference's assets decoded by the reference's own vendored stb_image
    (tools/make_texel_fixtures.py -> zig-weekend-raytracer_b200/data/texels; earth.png and wap.jpg in full, me.jpg decimated
    4x).  Oracle, host mirror and device always consume the same bytes."""
    assets = importlib.import_module("zig-weekend-raytracer_b200.assets")
    return assets.reference_images()


def bits(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)
