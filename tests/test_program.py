"""Host-side compilation of the entity tree (csrc/wrt_program.cu) without a GPU: wrt_check_scene compiles a scene exactly
as wrt_upload_scene does and checks the structure of what the kernels will read — the DFS program and its skip links,
the pruned packet program, and the SAH trees of the ordered traversal (every primitive of a BVH reachable exactly once)."""
from __future__ import annotations

import os

import pytest


SCENES = [("cornell_box", 0), ("emissive", 0), ("shrek_quads", 0), ("earth", 0), ("balls", 0), ("rtw_final", 0),
          ("synthetic", 4096), ("synthetic", 50000)]


@pytest.mark.parametrize("name,n_prims", SCENES)
def test_compiled_scene_is_structurally_sound(wrt, wro, images, name, n_prims):
    sc = wro.OracleScene(name, seed=1, n_prims=n_prims, images=images)
    info = wrt.check_scene(sc.flatten())
    assert info.n_prims == sc.n_prims
    assert info.n_ops >= info.n_prims + 1                      # every primitive has an op, plus OP_END
    assert info.n_ops_packet <= info.n_ops
    assert info.tree_depth <= info.max_nesting or info.tree_depth <= 2
    if info.n_prims >= 64:                                     # a SAH tree over n leaves stays shallow
        assert info.tree_depth <= 4 * max(1, info.n_prims).bit_length()
    sc.close()


def test_cornell_box_programs(wrt, wro):
    """The Cornell box is the headline workload: 13 primitives, a 27-op reference program (SURVEY.md A.10 topology), and the
    19-op packet program after pruning (7 of its 8 bvh_node boxes are the whole room; POP,POP fused)."""
    sc = wro.OracleScene("cornell_box")
    info = wrt.check_scene(sc.flatten())
    assert (info.n_prims, info.n_ops, info.n_ops_packet, info.n_lights) == (13, 27, 19, 2)
    sc.close()


def test_reference_tree_switch_keeps_the_reference_topology(wrt, wro):
    sc = wro.OracleScene("balls", seed=1)
    flat = sc.flatten()
    sah = wrt.check_scene(flat)
    os.environ["WRT_REFERENCE_TREE"] = "1"
    try:
        ref = wrt.check_scene(flat)
    finally:
        del os.environ["WRT_REFERENCE_TREE"]
    assert (ref.n_ops, ref.n_prims) == (sah.n_ops, sah.n_prims)
    assert ref.n_tree_records > 0 and sah.n_tree_records > 0   # four-wide records of either topology
    assert ref.tree_depth == 9                                 # balanced median split over 484 + 4 leaves (entity.zig:226-259)
    sc.close()


def test_invalid_scenes_are_rejected_with_a_reason(wrt, wro):
    sc = wro.OracleScene("emissive")
    flat = sc.flatten()
    good_root = flat.root
    flat.root = flat.n_entities + 5
    with pytest.raises(wrt.WrtError) as ei:
        wrt.check_scene(flat)
    assert "root" in ei.value.message
    flat.root = good_root
    wrt.check_scene(flat)
    sc.close()


def test_host_tree_build_is_deterministic_and_survives_coincident_centroids(wrt, wro):
    """wrt_build_trees on the host threads (cuda_device < 0): the concurrent first levels do not change the records, and a
    scene whose spheres share one centroid (no separating plane: the fallback split of wrt_treebuild.cuh) still yields a tree
    that reaches every primitive once (wrt_check_scene) with a logarithmic depth.  The device builder is compared with these
    bytes in tests/test_gpu_build.py; without a device it must fail loudly."""
    import numpy as np
    sc = wro.OracleScene("synthetic", seed=5, n_prims=40000)
    flat = sc.flatten()
    a_info, a2, a4 = wrt.build_trees(flat, -1)
    b_info, b2, b4 = wrt.build_trees(flat, -1)
    assert np.array_equal(a2, b2) and np.array_equal(a4, b4)
    assert a_info.n_records4 > 0 and a_info.n_records2 >= 40000 - 1 and a_info.on_device == 0
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(wrt.WrtError):
            wrt.build_trees(flat, 0)
    for i in range(flat.n_spheres):
        flat.spheres[i].center[0], flat.spheres[i].center[1], flat.spheres[i].center[2] = 1.0, 2.0, 3.0
        flat.spheres[i].radius = 0.25 + 1e-3 * i
        flat.spheres[i].is_moving = 0
    info = wrt.check_scene(flat)
    assert info.tree_depth <= 64
    sc.close()


def test_compact_form_is_offered_only_to_flat_single_tree_scenes(wrt, wro, images):
    """Compact stack entries + quantised four-wide records (wrt_device.cuh: TravCompactStack, Node4Q) need one tree of
    single-primitive leaves, no transforms and a primitive -> op table that inverts the program; wrt_check_scene verifies the
    tables and that every quantised box contains (by less than one step) the binary32 box it stands for."""
    sc = wro.OracleScene("synthetic", seed=2, n_prims=50000)
    info = wrt.check_scene(sc.flatten())
    assert info.compact_stack == 1 and info.quantised_records == info.n_tree_records > 0
    sc.close()
    for name in ("cornell_box", "balls", "rtw_final"):     # small (child-pair records) or instanced scenes keep the general form
        sc = wro.OracleScene(name, seed=1, images=images)
        info = wrt.check_scene(sc.flatten())
        assert info.compact_stack == 0 and info.quantised_records == 0
        sc.close()


@pytest.mark.parametrize("scale", [1e-6, 1.0, 3e7])
def test_quantised_records_stay_conservative_at_any_scene_scale(wrt, wro, scale):
    """Node4Q boxes are 8-bit offsets from the record's corner with a power-of-two step per axis: whatever the magnitude of
    the coordinates, the decoded box must contain the binary32 box and be less than one step loose (wrt_check_scene verifies
    every record), and the tree must still reach every primitive once."""
    sc = wro.OracleScene("synthetic", seed=7, n_prims=20000)
    flat = sc.flatten()
    for i in range(flat.n_spheres):
        s = flat.spheres[i]
        for k in range(3):
            s.center[k] *= scale
        s.radius *= scale
    for i in range(flat.n_quads):
        q = flat.quads[i]
        for k in range(3):
            q.start[k] *= scale
            q.u[k] *= scale
            q.v[k] *= scale
    info = wrt.check_scene(flat)
    assert info.compact_stack == 1 and info.quantised_records == info.n_tree_records
    sc.close()
