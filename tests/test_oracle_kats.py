"""Pins the ORACLE (CPU, no GPU needed).

The reference's own tests hold no golden vector for the render path, so the oracle is pinned by
 (1) the reference's math / writer known answers restated verbatim (math.zig:230-268, writer.zig:101-123),
 (2) Sobol van-der-Corput known answers + the pixel-containment invariant,
 (3) the BVH topologies the survey derived independently (A.10),
 (4) self-consistency: BVH traversal == brute force over all leaves, furnace / analytic checks,
 (5) committed golden fixtures (tests/golden) guarding against regressions of the oracle itself.
"""
from __future__ import annotations

import math
from pathlib import Path

import numpy as np
import pytest

GOLDEN = Path(__file__).resolve().parent / "golden"


# ---- (1) reference unit tests restated ---------------------------------------------------------------------
def test_math_cross_known_answers(wro):  # math.zig:230-241
    out = np.zeros(3)
    wro.lib.wro_math_cross(wro._ptr(np.array([1.0, 0, 0])), wro._ptr(np.array([0.0, 1, 0])), wro._ptr(out))
    np.testing.assert_allclose(out, [0, 0, 1], rtol=1e-6)
    wro.lib.wro_math_cross(wro._ptr(np.array([1.0, 0, 0])), wro._ptr(np.array([0.0, -1, 0])), wro._ptr(out))
    np.testing.assert_allclose(out, [0, 0, -1], rtol=1e-6)


def test_math_dot_length_normalize(wro):  # math.zig:247-268
    u = np.array([1.0, 1, 1])
    v = np.array([2.0, 2, 2])
    assert wro.lib.wro_math_dot(wro._ptr(u), wro._ptr(v)) == pytest.approx(6.0, rel=1e-8)
    assert wro.lib.wro_math_length(wro._ptr(u)) == pytest.approx(math.sqrt(3.0), rel=1e-8)
    out = np.zeros(3)
    w = np.array([1.0, 2, 3])
    wro.lib.wro_math_normalize(wro._ptr(w), wro._ptr(out))
    assert wro.lib.wro_math_length(wro._ptr(out)) == pytest.approx(1.0, rel=1e-6)


def test_writer_size_of_line_and_digit(wro):  # writer.zig:101-123
    def line(px):
        a = np.array(px, np.uint8)
        return wro.lib.wro_size_of_line(wro._ptr(a))
    assert line([0, 0, 0]) == 6
    assert line([0, 255, 0]) == 8
    assert line([255, 255, 255]) == 12
    for digit, size in [(0, 1), (9, 1), (10, 2), (99, 2), (100, 3), (255, 3)]:
        assert wro.lib.wro_size_of_digit(digit) == size


def test_encode_color_table(wro):  # writer.zig:68-94; table from SURVEY.md §7.4
    cases = [(0.0, 0), (1.0, 255), (0.25, 128), (2.0, 255), (float("nan"), 0), (1e-12, 0), (0.999 ** 2, 255),
             (float("inf"), 255)]
    for x, want in cases:
        got = wro.encode_color([x, x, x])
        assert list(got) == [want] * 3, (x, got)
    # every channel is independent
    assert list(wro.encode_color([0.0, 0.25, 1.0])) == [0, 128, 255]


# ---- (2) Sobol ------------------------------------------------------------------------------------------------
def test_sobol_van_der_corput_known_answers(wro):
    idx = np.arange(1, 5, dtype=np.uint64)
    d0 = wro.sobol_dimension_samples(idx, np.zeros(4, np.uint32), False, 0)
    d1 = wro.sobol_dimension_samples(idx, np.ones(4, np.uint32), False, 0)
    np.testing.assert_array_equal(d0, np.array([0.5, 0.25, 0.75, 0.125], np.float32))
    np.testing.assert_array_equal(d1, np.array([0.5, 0.75, 0.25, 0.625], np.float32))


def test_sobol_pixel_known_answers_400(wro):  # SURVEY.md A.10, 400x400 (m = 9, scale 512)
    cases = [
        ((0, 0, 0), 0, (0.0, 0.0)),
        ((0, 0, 1), 394752, (0.7529296875, 0.7529296875)),
        ((0, 0, 2), 657920, (0.62744140625, 0.37646484375)),
        ((1, 0, 0), 197376, (0.505859375, 0.501953125)),
        ((0, 1, 0), 131584, (0.501953125, 0.505859375)),
        ((399, 399, 127), 33458147, (0.76171875, 0.24224853515625)),
        ((200, 100, 5), 1403942, (0.427001953125, 0.427001953125)),
    ]
    cols = [c[0][0] for c in cases]
    rows = [c[0][1] for c in cases]
    ss = [c[0][2] for c in cases]
    index, off = wro.sobol_pixel_samples(400, 400, cols, rows, ss)
    for (key, want_idx, want_off), i, o in zip(cases, index, off):
        assert int(i) == want_idx, key
        assert tuple(o) == want_off, key


@pytest.mark.parametrize("wh", [(400, 400), (1024, 1024), (1920, 1080), (3840, 2160)])
def test_sobol_pixel_containment(wro, wh):  # floor(sobolSample(idx, d) * scale) == pixel coordinate (A.10)
    w, h = wh
    rng = np.random.default_rng(5)
    n = 3000
    cols = rng.integers(0, w, n).astype(np.uint32)
    rows = rng.integers(0, h, n).astype(np.uint32)
    ss = rng.integers(0, 1024, n).astype(np.uint32)
    index, off = wro.sobol_pixel_samples(w, h, cols, rows, ss)
    scale = 1 << int(math.ceil(math.log2(max(w, h))))
    x = wro.sobol_dimension_samples(index, np.zeros(n, np.uint32), False, 0).astype(np.float64) * scale
    y = wro.sobol_dimension_samples(index, np.ones(n, np.uint32), False, 0).astype(np.float64) * scale
    # f32 rounding may land one ulp outside the pixel at scale 4096; the clamp in getPixel2D absorbs it
    assert np.sum(np.floor(x) != cols) <= 2 and np.sum(np.floor(y) != rows) <= 2
    assert np.all((off >= 0.0) & (off < 1.0))
    # distinct samples of one pixel get distinct indices
    i2, _ = wro.sobol_pixel_samples(w, h, np.full(64, cols[0]), np.full(64, rows[0]), np.arange(64))
    assert len(set(i2.tolist())) == 64


def test_sobol_owen_dimensions_wrap_and_range(wro):  # sampler.zig:203-209: dimension starts at 2, wraps at 1024
    seq = wro.sobol_get1d_sequence(400, 400, 17, 33, 5, True, 1234, 1022 + 4)
    assert np.all((seq >= 0.0) & (seq < 1.0))
    np.testing.assert_array_equal(seq[1022:1026], seq[0:4])  # wrapped back to dimension 2
    plain = wro.sobol_get1d_sequence(400, 400, 17, 33, 5, False, 1234, 8)
    assert not np.array_equal(plain, seq[:8])  # the scramble does something
    # owen_fast scrambling is a bijection on u32: 256 consecutive indices stay distinct in a scrambled dimension
    idx = np.arange(256, dtype=np.uint64)
    vals = wro.sobol_dimension_samples(idx, np.full(256, 7, np.uint32), True, 99)
    assert len(np.unique(vals)) == 256


def test_counter_rng_is_philox4x32_10(wro):  # Random123 known-answer vector, counter = key = 0
    assert wro.lib.wro_counter_rng_bits(0, 0, 0, 0) == 0xE169C58D6627E8D5
    assert wro.lib.wro_counter_rng_bits(0, 0, 0, 1) == 0x9B00DBD8BC57AC4C


# ---- (3) BVH topology ---------------------------------------------------------------------------
def test_bvh_topology_cornell_box(wro):
    sc = wro.OracleScene("cornell_box")
    kinds, mats, centers = sc.prim_table()
    # materials: 0 red, 1 white, 2 green, 3 light, 4 glass, 5 metal (scene.zig:330-346)
    # DFS leaf order: red, floor, box2T (front,right,back,left,top,bottom), green, ceil, glass, light, back
    assert mats.tolist() == [0, 1, 5, 5, 5, 5, 5, 5, 2, 1, 4, 3, 1]
    assert kinds.tolist() == [1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 1, 1]
    np.testing.assert_allclose(centers[0], [0, 277.5, 277.5])      # red wall x = 0
    np.testing.assert_allclose(centers[1], [277.5, 0, 277.5])      # floor
    np.testing.assert_allclose(centers[8], [555, 277.5, 277.5])    # green wall x = 555
    np.testing.assert_allclose(centers[9], [277.5, 555, 277.5])    # ceiling
    np.testing.assert_allclose(centers[12], [277.5, 277.5, 555])   # back wall
    # the quirky cached box of Translate(RotateY(box)) (A.9-4, A.9-5, values from A.10)
    flat = sc.flatten()
    tr = [flat.entities[i] for i in range(flat.n_entities) if flat.entities[i].kind == 4]
    assert len(tr) == 1
    np.testing.assert_allclose(list(tr[0].bbox_min), [-265.00006, -5e-5, -337.70520], rtol=0, atol=6e-5)
    np.testing.assert_allclose(list(tr[0].bbox_max), [467.08297, 165.00005, 454.37782], rtol=0, atol=6e-5)
    sc.close()


def test_bvh_topology_emissive_and_node_counts(wro):
    sc = wro.OracleScene("emissive")
    kinds, mats, centers = sc.prim_table()
    # [ground, glass | lightquad, lightsphere]; materials: 0 glass, 1 ground, 2 blue light, 3 green light
    assert mats.tolist() == [1, 0, 2, 3]
    flat = sc.flatten()
    assert sum(1 for i in range(flat.n_entities) if flat.entities[i].kind == 3) == 3  # N=4 -> 3 nodes
    sc.close()
    sc = wro.OracleScene("cornell_box")
    flat = sc.flatten()
    assert sum(1 for i in range(flat.n_entities) if flat.entities[i].kind == 3) == 7  # N=8 -> 7 nodes
    sc.close()
    sc = wro.OracleScene("rtw_final", seed=1)
    flat = sc.flatten()
    nodes = sum(1 for i in range(flat.n_entities) if flat.entities[i].kind == 3)
    assert nodes == 7 + 1023 + 511  # top level (8) + 1000 spheres + 400 ground boxes
    sc.close()


def test_rtw_final_top_level_order(wro):  # A.10: [ground_boxes, glass260, glass360, metal0 | light, me, shrek, ballsT]
    sc = wro.OracleScene("rtw_final", seed=1)
    flat = sc.flatten()
    root = flat.entities[flat.root]
    assert root.kind == 2 and root.b == 8
    order = [flat.children[root.a + k] for k in range(8)]
    kinds = [flat.entities[i].kind for i in order]
    # after the in-place sorts: left half {ground_boxes(coll), glass260, glass360, metal0}, right half {ballsT, light, me, shrek}
    assert sorted(kinds[:4]) == [0, 0, 0, 2]
    assert sorted(kinds[4:]) == [0, 0, 1, 4]
    sc.close()


# ---- (4) self-consistency ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["cornell_box", "emissive", "balls", "rtw_final", "synthetic"])
def test_bvh_equals_brute_force_on_primary_rays(wro, name):
    """The reference's (weak) culling never changes the closest hit of a primary ray: same id, same t bits."""
    sc = wro.OracleScene(name, seed=1, n_prims=2048)
    w, h = (48, 48) if name != "synthetic" else (24, 24)
    cam = sc.camera(w, h)
    p = sc.params(w, h, 4, 10)
    ids_a, t_a = sc.primary_hits(cam, p, 2)
    sc.set_no_cull(True)
    ids_b, t_b = sc.primary_hits(cam, p, 2)
    np.testing.assert_array_equal(ids_a, ids_b)
    np.testing.assert_array_equal(t_a.view(np.uint64), t_b.view(np.uint64))
    assert (ids_a != 0xFFFFFFFF).any()
    sc.close()


def test_quad_closed_and_sphere_open_interval_tie_rule(wro):
    """Box faces sharing a plane: the later coincident quad replaces the earlier one (entity.zig:485 closed interval),
    a sphere does not replace an equal-t hit (entity.zig:608 open interval)."""
    sc = wro.OracleScene("cornell_box")
    # the metal box's bottom face (prim 7) lies in the floor plane (prim 1, visited first): a ray from below the floor
    # straight up through the box footprint hits both at the same t; the last visited wins.
    o = np.array([[265.0 + 40.0, -10.0, 295.0 + 60.0]])
    d = np.array([[0.0, 1.0, 0.0]])
    hit = sc.trace_rays(o, d)
    assert hit["t"][0] == 10.0
    assert hit["prim_id"][0] == 7
    sc.close()


def test_white_furnace_emissive_background(wro):
    """A Lambertian scene under a constant white background with albedo 1 keeps radiance 1 (energy conservation of the
    cosine estimator: attenuation * scatteringPdf / pdf == albedo)."""
    abi = wro.abi
    import ctypes as C
    # one big white sphere, no lights, flat arrays by hand
    ents = (abi.Entity * 2)()
    ents[0].kind = 2; ents[0].a = 0; ents[0].b = 1; ents[0].c = 0xFFFFFFFF
    ents[1].kind = 0; ents[1].a = 0
    for k in range(3):
        ents[1].bbox_min[k] = -1.0; ents[1].bbox_max[k] = 1.0
    children = (C.c_uint32 * 1)(1)
    sph = (abi.Sphere * 1)()
    sph[0].radius = 1.0
    mats = (abi.Material * 1)()
    mats[0].kind = 0; mats[0].texture = 0
    tex = (abi.Texture * 1)()
    tex[0].kind = 0
    for k in range(3):
        tex[0].color[k] = 1.0
    flat = abi.Scene(abi_version=abi.WRT_ABI_VERSION, root=0, lights=0xFFFFFFFF, n_entities=2, n_children=1, n_spheres=1, n_quads=0,
                     n_materials=1, n_textures=1, n_images=0, entities=ents, children=children, spheres=sph,
                     materials=mats, textures=tex)
    sc = wro.OracleScene(flat=flat)
    cam = abi.Camera()
    cam.position[2] = -3.0
    cam.pixel00_loc[0] = -0.4; cam.pixel00_loc[1] = 0.4; cam.pixel00_loc[2] = -2.0
    cam.pixel_delta_u[0] = 0.1
    cam.pixel_delta_v[1] = -0.1
    p = abi.Params(width=8, height=8, samples_per_pixel=32, max_ray_bounce_depth=50, seed=3, row_shard_count=1)
    for k in range(3):
        p.background_color[k] = 1.0
    for mode in (wro.RNG_REFERENCE, wro.RNG_COUNTER):
        fb, st = sc.render(cam, p, mode)
        np.testing.assert_allclose(fb[..., :3], 1.0, rtol=1e-12)
        assert st.rays > st.paths  # bounces happened
    sc.close()


def test_mixture_and_cosine_estimators_agree_in_the_mean(wro):
    """Light importance sampling (pdf.zig mixture) must not change the expected image: render cornell_box with and
    without the light list and compare the mean radiance (statistical; loose bound)."""
    sc = wro.OracleScene("cornell_box")
    cam = sc.camera(32, 32)
    p = sc.params(32, 32, 256, 12, seed=11)
    with_lights, _ = sc.render(cam, p, wro.RNG_COUNTER)
    flat = sc.flatten()
    flat_nolights = wro.abi.Scene.from_buffer_copy(flat)
    flat_nolights.lights = 0xFFFFFFFF
    sc2 = wro.OracleScene(flat=flat_nolights)
    without, _ = sc2.render(cam, p, wro.RNG_COUNTER)
    m1 = np.nanmean(with_lights[..., :3])
    m2 = np.nanmean(without[..., :3])
    assert abs(m1 - m2) / m2 < 0.15, (m1, m2)
    sc.close(); sc2.close()


def test_reference_and_counter_rng_agree_statistically(wro):
    """The counter stream + direct sphere/circle sampling is distribution-equivalent to the restated std.Random path."""
    sc = wro.OracleScene("emissive")
    cam = sc.camera(40, 40)
    p = sc.params(40, 40, 256, 10, seed=5)
    a, _ = sc.render(cam, p, wro.RNG_REFERENCE)
    b, _ = sc.render(cam, p, wro.RNG_COUNTER)
    assert abs(np.nanmean(a[..., :3]) - np.nanmean(b[..., :3])) / np.nanmean(a[..., :3]) < 0.03
    # metal fuzz path (unit-sphere sampler) on balls
    sc.close()
    sc = wro.OracleScene("balls", seed=1)
    cam = sc.camera(48, 27)
    p = sc.params(48, 27, 128, 20, seed=5)
    a, _ = sc.render(cam, p, wro.RNG_REFERENCE)
    b, _ = sc.render(cam, p, wro.RNG_COUNTER)
    assert abs(np.nanmean(a[..., :3]) - np.nanmean(b[..., :3])) / np.nanmean(a[..., :3]) < 0.02
    sc.close()


def test_render_flags_sample_range_and_row_shards(wro):
    """sample ranges accumulate to the full frame; row shards tile it (the oracle mirrors the ABI's sharding fields)."""
    sc = wro.OracleScene("emissive")
    w, h, spp = 33, 20, 8
    cam = sc.camera(w, h)
    p = sc.params(w, h, spp, 10, seed=9)
    full, _ = sc.render(cam, p, wro.RNG_COUNTER)
    acc = np.zeros_like(full)
    p1 = sc.params(w, h, spp, 10, seed=9, sample_begin=0, sample_end=3)
    sc.render_into(acc, cam, p1, wro.RNG_COUNTER)
    p2 = sc.params(w, h, spp, 10, seed=9, sample_begin=3, sample_end=8, flags=wro.abi.WRT_FLAG_NO_CLEAR)
    sc.render_into(acc, cam, p2, wro.RNG_COUNTER)
    np.testing.assert_allclose(acc, full, rtol=1e-12, atol=1e-15)
    tiles = np.zeros_like(full)
    for r in range(3):
        ps = sc.params(w, h, spp, 10, seed=9, row_shard_index=r, row_shard_count=3)
        part, _ = sc.render(cam, ps, wro.RNG_COUNTER)
        tiles[r::3] = part
    np.testing.assert_array_equal(tiles, full)
    sc.close()


def test_flatten_round_trip_is_exact(wro, images):
    for name in ["cornell_box", "rtw_final", "balls", "shrek_quads", "earth"]:
        sc = wro.OracleScene(name, seed=1, images=images)
        sc2 = wro.OracleScene(flat=sc.flatten())
        cam = sc.camera(24, 24)
        p = sc.params(24, 24, 2, 12, seed=3)
        a, _ = sc.render(cam, p, wro.RNG_COUNTER)
        b, _ = sc2.render(cam, p, wro.RNG_COUNTER)
        np.testing.assert_array_equal(a, b)
        assert sc.n_prims == sc2.n_prims
        sc.close(); sc2.close()


def test_camera_viewport_invariants(wro):  # camera.zig:117-157
    sc = wro.OracleScene("cornell_box")
    for w, h in [(400, 400), (1920, 1080)]:
        cam = sc.camera(w, h)
        p00 = np.array(cam.pixel00_loc); du = np.array(cam.pixel_delta_u); dv = np.array(cam.pixel_delta_v)
        pos = np.array(cam.position)
        centre = p00 + du * (w - 1) / 2 + dv * (h - 1) / 2
        # the viewport centre lies focus_dist (10) in front of the camera along -w = +z
        np.testing.assert_allclose(centre - pos, [0, 0, 10.0], atol=1e-9)
        vh = 2 * math.tan(math.radians(40.0) / 2) * 10.0
        np.testing.assert_allclose(np.linalg.norm(dv) * h, vh, rtol=1e-12)
        np.testing.assert_allclose(np.linalg.norm(du) * w, vh * w / h, rtol=1e-12)
        assert cam.is_depth_of_field == 0
    sc.close()
    sc = wro.OracleScene("balls", seed=1)
    assert sc.camera(64, 36).is_depth_of_field == 1
    sc.close()


# ---- (5) golden fixtures ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["cornell_box", "emissive", "balls", "rtw_final", "shrek_quads", "earth", "synthetic"])
def test_oracle_matches_golden_fixture(wro, images, name):
    g = np.load(GOLDEN / f"{name}.npz")
    sc = wro.OracleScene(name, seed=int(g["scene_seed"]), n_prims=int(g["n_prims_arg"]), images=images)
    w, h, spp, depth = (int(g[k]) for k in ("width", "height", "spp", "depth"))
    cam = sc.camera(w, h)
    p = sc.params(w, h, spp, depth, seed=int(g["seed"]))
    ids, t = sc.primary_hits(cam, p, int(g["n_primary"]))
    np.testing.assert_array_equal(ids, g["prim_ids"])
    np.testing.assert_array_equal(t.view(np.uint64), g["t_bits"])
    fb, st = sc.render(cam, p, wro.RNG_COUNTER)
    # libm (sin/cos/acos/atan2/pow) may differ between hosts by an ulp: compare the radiance loosely, the
    # integer outputs exactly
    np.testing.assert_allclose(fb[..., :3], g["radiance"], rtol=1e-9, atol=1e-12, equal_nan=True)
    assert abs(int(st.rays) - int(g["rays"])) <= max(2, int(g["rays"]) // 10000)
    sc.close()
