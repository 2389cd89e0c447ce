"""The ordered traversal's trees built on the device (csrc/wrt_build.cu; the reference builds its BVH in
src/entity.zig:226-259) against the host builder (csrc/wrt_program.cu): both follow csrc/wrt_treebuild.cuh and must write
the same records byte for byte — child-pair records of the binned-SAH trees and their four-wide collapse — for every scene,
including trees small enough for one warp, trees that need the big-segment path, nested trees (rtw_final) and inputs with
coincident centroids (the fallback split).  Frames rendered from a device-built scene equal the host-built ones bit for bit."""
from __future__ import annotations

import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SCENES = [("cornell_box", 0), ("shrek_quads", 0), ("balls", 0), ("rtw_final", 0), ("synthetic", 1500), ("synthetic", 40000),
          ("synthetic", 1 << 18)]


@pytest.mark.parametrize("name,n_prims", SCENES)
def test_device_build_writes_the_host_builders_bytes(wrt, wro, images, name, n_prims):
    sc = wro.OracleScene(name, seed=3, n_prims=n_prims, images=images)
    flat = sc.flatten()
    hi, h2, h4 = wrt.build_trees(flat, -1)
    di, d2, d4 = wrt.build_trees(flat, 0)
    assert di.on_device == 1 and hi.on_device == 0
    assert (di.n_records2, di.n_records4, di.stack_depth, di.use_wide, di.max_nesting) == \
           (hi.n_records2, hi.n_records4, hi.stack_depth, hi.use_wide, hi.max_nesting)
    bad2 = np.nonzero((h2 != d2).any(axis=1))[0]
    assert bad2.size == 0, f"{bad2.size} child-pair records differ, first {bad2[:5]}"
    bad4 = np.nonzero((h4 != d4).any(axis=1))[0]
    assert bad4.size == 0, f"{bad4.size} four-wide records differ, first {bad4[:5]}"
    print(f"{name} {n_prims}: {hi.n_records2} / {hi.n_records4} records, host {hi.build_ms:.2f} ms, device {di.build_ms:.2f} ms "
          f"(whole call {hi.total_ms:.1f} / {di.total_ms:.1f} ms)")
    sc.close()


def test_device_build_at_the_c5_size(wrt, wro):
    """2^20 primitives (BASELINE configs[4]): same bytes, and the device build is the faster one by a wide margin."""
    sc = wro.OracleScene("synthetic", seed=1, n_prims=1 << 20)
    flat = sc.flatten()
    hi, h2, h4 = wrt.build_trees(flat, -1)
    di, d2, d4 = wrt.build_trees(flat, 0)   # first call pays context / allocator warm-up
    di, d2, d4 = wrt.build_trees(flat, 0)
    assert np.array_equal(h2, d2) and np.array_equal(h4, d4)
    print(f"2^20 primitives: host build {hi.build_ms:.1f} ms, device build {di.build_ms:.2f} ms "
          f"(whole call incl. program compile and transfers: {hi.total_ms:.0f} / {di.total_ms:.0f} ms)")
    assert di.build_ms < hi.build_ms
    sc.close()


def coincide(flat):
    """every sphere gets one centre (radii differ): all their tight boxes share a centroid"""
    for i in range(flat.n_spheres):
        flat.spheres[i].center[0], flat.spheres[i].center[1], flat.spheres[i].center[2] = 1.0, 2.0, 3.0
        flat.spheres[i].radius = 0.25 + 1e-3 * i
        flat.spheres[i].is_moving = 0


def test_coincident_centroids_take_the_fallback_split_on_both_sides(wrt, wro):
    """No plane separates items with one centroid: both builders halve the current order (wrt_treebuild.cuh) and agree."""
    sc = wro.OracleScene("synthetic", seed=5, n_prims=6000)
    flat = sc.flatten()
    coincide(flat)
    hi, h2, h4 = wrt.build_trees(flat, -1)
    di, d2, d4 = wrt.build_trees(flat, 0)
    assert np.array_equal(h2, d2) and np.array_equal(h4, d4)
    assert (hi.stack_depth, hi.max_nesting) == (di.stack_depth, di.max_nesting)
    sc.close()


@pytest.mark.parametrize("name,n_prims,w,h,spp,depth", [("balls", 0, 96, 54, 8, 20), ("rtw_final", 0, 64, 64, 8, 20),
                                                        ("synthetic", 40000, 96, 54, 8, 12)])
def test_frames_from_device_built_trees_equal_host_built(wrt, wro, images, name, n_prims, w, h, spp, depth):
    sc = wro.OracleScene(name, seed=2, n_prims=n_prims, images=images)
    cam = sc.camera(w, h)
    params = sc.params(w, h, spp, depth, seed=11)
    frames = {}
    for mode in ("0", "1"):
        os.environ["WRT_DEVICE_BUILD"] = mode
        try:
            with wrt.Context(0) as c:
                c.upload_scene(sc.flatten())
                frames[mode] = c.render(cam, params).copy()
                st = c.stats()
                assert st.tree_build_device == int(mode)
        finally:
            del os.environ["WRT_DEVICE_BUILD"]
    assert np.array_equal(frames["0"].view(np.uint64), frames["1"].view(np.uint64))
    sc.close()
