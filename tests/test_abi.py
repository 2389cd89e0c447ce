"""The C-ABI boundary without a GPU: the library loads, exports exactly what include/wrt.h declares, the ctypes
mirror has the C layout, and every compute entry point fails loudly (no CPU fallback) when no device exists."""
from __future__ import annotations

import ctypes as C
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "wrt.h"


def declared_functions():
    text = HEADER.read_text()
    return sorted(set(re.findall(r"^WRT_API\s+[\w\s\*]+?\b(wrt_\w+)\s*\(", text, flags=re.M)))


def test_header_declares_the_expected_entry_points(wrt):
    names = declared_functions()
    assert names == sorted(wrt.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol(wrt):
    lib = C.CDLL(str(wrt.LIB_PATH))
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} is declared in include/wrt.h but not exported by libwrt.so"
    assert lib.wrt_abi_version() == 3


def test_library_exports_nothing_else(wrt):
    out = subprocess.check_output(["nm", "-D", "--defined-only", str(wrt.LIB_PATH)], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    extra = {s for s in exported if not s.startswith("wrt_") and not s.startswith("_")}
    assert not extra, extra
    assert {s for s in exported if s.startswith("wrt_")} == set(declared_functions())


def test_ctypes_mirror_matches_c_layout(wrt, tmp_path):
    """sizeof/offsetof of every POD struct, as gcc sees include/wrt.h, equals the ctypes mirror."""
    structs = {
        "wrt_entity": (wrt.Entity, ["kind", "a", "b", "c", "p", "bbox_min", "bbox_max"]),
        "wrt_sphere": (wrt.Sphere, ["center", "radius", "movement", "material", "is_moving"]),
        "wrt_quad": (wrt.Quad, ["start", "u", "v", "w", "normal", "offset", "area", "material"]),
        "wrt_material": (wrt.Material, ["kind", "texture", "albedo", "param"]),
        "wrt_texture": (wrt.Texture, ["kind", "even", "odd", "image", "color", "inv_scale"]),
        "wrt_image": (wrt.Image, ["width", "height", "num_components", "bytes_per_row", "texel_offset"]),
        "wrt_scene": (wrt.Scene, ["abi_version", "root", "lights", "n_entities", "n_images", "entities", "texels", "texel_bytes"]),
        "wrt_camera": (wrt.Camera, ["position", "pixel00_loc", "pixel_delta_u", "pixel_delta_v", "defocus_disk_u",
                                    "defocus_disk_v", "is_depth_of_field"]),
        "wrt_params": (wrt.Params, ["width", "height", "samples_per_pixel", "max_ray_bounce_depth", "background_color",
                                    "clear_color", "seed", "row_shard_index", "row_shard_count", "sample_begin", "sample_end",
                                    "cull_mode", "flags"]),
        "wrt_scene_info": (wrt.SceneInfo, ["n_ops", "n_ops_packet", "n_prims", "n_boxes", "n_tree_records", "tree_depth", "max_nesting",
                                           "n_lights", "ref_boxes_loose", "stack_depth", "compact_stack", "quantised_records"]),
        "wrt_stats": (wrt.Stats, ["paths", "rays", "render_ms", "kernel_ms", "upload_ms", "kernel_launches", "program_ops",
                                  "n_prims", "cull_mode_used", "traversal_steps", "ref_boxes_loose", "n_devices", "gather_ms",
                                  "kernel_ms_min", "kernel_ms_max", "tree_build_ms", "tree_build_device", "n_tree_records"]),
        "wrt_tree_info": (wrt.TreeInfo, ["n_records2", "n_records4", "stack_depth", "use_wide", "on_device", "max_nesting",
                                         "build_ms", "total_ms"]),
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for cname, (_, fields) in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for f in fields:
            lines.append(f'printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c11", "-o", str(exe), str(src)])
    got = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, (cls, fields) in structs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for f in fields:
            assert int(got[f"{cname}.{f}"]) == getattr(cls, f).offset, (cname, f)


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "c89ish.c"
    src.write_text(f'#include "{HEADER}"\nint main(void) {{ return (int)WRT_ABI_VERSION - 3; }}\n')
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-o", str(tmp_path / "a.out"), str(src)])


@pytest.mark.skipif("__import__('torch').cuda.is_available()", reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_device(wrt):
    with pytest.raises(wrt.WrtError) as ei:
        wrt.Context(0)
    assert ei.value.code == -2  # WRT_E_CUDA
    assert "no CPU fallback" in ei.value.message


def test_null_context_is_rejected(wrt):
    lib = wrt.lib
    assert lib.wrt_upload_scene(None, None) == -1
    assert lib.wrt_render(None, None, None, None, 32) == -1
    assert lib.wrt_get_stats(None, None) == -1
    assert lib.wrt_create(0, None) == -1
    lib.wrt_destroy(None)  # no-op


def test_product_does_not_link_or_import_the_oracle(wrt):
    """The product path must never route through oracle/: no dynamic dependency, no import."""
    out = subprocess.check_output(["ldd", str(wrt.LIB_PATH)], text=True)
    assert "libwro" not in out
    pkg = ROOT / "zig-weekend-raytracer_b200"
    sources = [p for ext in ("*.py", "*.cu", "*.cuh", "*.h", "*.c", "*.cpp", "*.hpp") for p in pkg.rglob(ext)]
    assert sources
    for path in sources:
        text = path.read_text()
        for needle in ("libwro", "wro_py", "import wro", '#include "wro', "wro.h"):
            assert needle not in text, (path, needle)
        # the only library bound at run time is NCCL (csrc/wrt_multi.cu: one NCCL per process, PyTorch's when it is loaded)
        if "dlopen" in text:
            assert path.name == "wrt_multi.cu", path
            import re
            names = re.findall(r'"(lib[^"]*\.so[^"]*)"', text)
            assert names and all(n.startswith("libnccl.so") for n in names), names
