"""ctypes mirror of include/wrt.h (structures and constants only; loads no library).

Shared by the product binding (``__init__.py``) and by the test-side checker's binding, because both speak the same
POD types at the boundary.
"""
from __future__ import annotations

import ctypes as C

WRT_ABI_VERSION = 3
WRT_NONE = 0xFFFFFFFF
WRT_CULL_AUTO = 0       # default: TIGHT when every reference box contains its subtree, else REFERENCE
WRT_CULL_REFERENCE = 1
WRT_CULL_TIGHT = 2
WRT_FLAG_NO_CLEAR = 1
WRT_FLAG_DISABLE_DOF = 2
WRT_FLAG_FORCE_LANE = 4
WRT_FLAG_FORCE_PACKET = 8
WRT_FLAG_ENGINE_MEGAKERNEL = 16
WRT_FLAG_ENGINE_WAVEFRONT = 32
WRT_FLAG_ENGINE_SYNC = 64
WRT_FLAG_ENGINE_REGROUP = 128
WRT_FLAG_SHARD_SAMPLES = 256
WRT_FLAG_SAMPLER_SOBOL = 512
WRT_COMM_ID_BYTES = 128


def WRT_FLAG_CHUNKS(n: int) -> int:
    """Pin the number of sample chunks (include/wrt.h): two engines given the same n produce the same bits."""
    return (int(n) & 0xFF) << 24


WRT_TRAV_FORCE_LANE = 0x100
WRT_TRAV_FORCE_PACKET = 0x200

ENT_SPHERE, ENT_QUAD, ENT_COLLECTION, ENT_BVH_NODE, ENT_TRANSLATE, ENT_ROTATE_Y = range(6)
MAT_LAMBERTIAN, MAT_ISOTROPIC, MAT_METAL, MAT_DIELECTRIC, MAT_DIFFUSE_EMISSIVE = range(5)
TEX_SOLID, TEX_CHECKER, TEX_IMAGE = range(3)

_d3 = C.c_double * 3


class Entity(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("a", C.c_uint32), ("b", C.c_uint32), ("c", C.c_uint32),
                ("p", _d3), ("bbox_min", _d3), ("bbox_max", _d3)]


class Sphere(C.Structure):
    _fields_ = [("center", _d3), ("radius", C.c_double), ("movement", _d3),
                ("material", C.c_uint32), ("is_moving", C.c_uint32)]


class Quad(C.Structure):
    _fields_ = [("start", _d3), ("u", _d3), ("v", _d3), ("w", _d3), ("normal", _d3),
                ("offset", C.c_double), ("area", C.c_double), ("material", C.c_uint32), ("_pad", C.c_uint32)]


class Material(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("texture", C.c_uint32), ("albedo", _d3), ("param", C.c_double)]


class Texture(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("even", C.c_uint32), ("odd", C.c_uint32), ("image", C.c_uint32),
                ("color", _d3), ("inv_scale", C.c_double)]


class Image(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("num_components", C.c_uint32),
                ("bytes_per_row", C.c_uint32), ("texel_offset", C.c_uint64)]


class Scene(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("root", C.c_uint32), ("lights", C.c_uint32),
                ("n_entities", C.c_uint32), ("n_children", C.c_uint32), ("n_spheres", C.c_uint32),
                ("n_quads", C.c_uint32), ("n_materials", C.c_uint32), ("n_textures", C.c_uint32),
                ("n_images", C.c_uint32),
                ("entities", C.POINTER(Entity)), ("children", C.POINTER(C.c_uint32)),
                ("spheres", C.POINTER(Sphere)), ("quads", C.POINTER(Quad)),
                ("materials", C.POINTER(Material)), ("textures", C.POINTER(Texture)),
                ("images", C.POINTER(Image)), ("texels", C.POINTER(C.c_uint8)), ("texel_bytes", C.c_uint64)]


class Camera(C.Structure):
    _fields_ = [("position", _d3), ("pixel00_loc", _d3), ("pixel_delta_u", _d3), ("pixel_delta_v", _d3),
                ("defocus_disk_u", _d3), ("defocus_disk_v", _d3), ("is_depth_of_field", C.c_uint32),
                ("_pad", C.c_uint32)]


class Params(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("samples_per_pixel", C.c_uint32),
                ("max_ray_bounce_depth", C.c_uint32), ("background_color", _d3), ("clear_color", _d3),
                ("seed", C.c_uint64), ("row_shard_index", C.c_uint32), ("row_shard_count", C.c_uint32),
                ("sample_begin", C.c_uint32), ("sample_end", C.c_uint32), ("cull_mode", C.c_uint32),
                ("flags", C.c_uint32)]


class SceneInfo(C.Structure):
    _fields_ = [("n_ops", C.c_uint32), ("n_ops_packet", C.c_uint32), ("n_prims", C.c_uint32), ("n_boxes", C.c_uint32),
                ("n_tree_records", C.c_uint32), ("tree_depth", C.c_uint32), ("max_nesting", C.c_uint32), ("n_lights", C.c_uint32),
                ("ref_boxes_loose", C.c_uint32), ("stack_depth", C.c_uint32), ("compact_stack", C.c_uint32),
                ("quantised_records", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("rays", C.c_uint64), ("render_ms", C.c_double), ("kernel_ms", C.c_double),
                ("upload_ms", C.c_double), ("kernel_launches", C.c_uint32), ("program_ops", C.c_uint32),
                ("n_prims", C.c_uint32), ("cull_mode_used", C.c_uint32), ("traversal_steps", C.c_uint64),
                ("ref_boxes_loose", C.c_uint32), ("n_devices", C.c_uint32), ("gather_ms", C.c_double),
                ("kernel_ms_min", C.c_double), ("kernel_ms_max", C.c_double), ("tree_build_ms", C.c_double),
                ("tree_build_device", C.c_uint32), ("n_tree_records", C.c_uint32)]


class TreeInfo(C.Structure):
    _fields_ = [("n_records2", C.c_uint32), ("n_records4", C.c_uint32), ("stack_depth", C.c_uint32), ("use_wide", C.c_uint32),
                ("on_device", C.c_uint32), ("max_nesting", C.c_uint32), ("build_ms", C.c_double), ("total_ms", C.c_double)]
