"""The product's host side (zig-weekend-raytracer_b200/host: scene construction, flattening, PPM writer, CLI) checked on
the CPU against the oracle and against the reference's own unit-test known answers.  Nothing here needs a GPU: the
host code only builds trees and writes files; rays are never traced on the host."""
from __future__ import annotations

import importlib
import subprocess

import numpy as np
import pytest


@pytest.fixture(scope="module")
def host(wrt):
    return importlib.import_module("zig-weekend-raytracer_b200.host")


NONE = 0xFFFFFFFF


@pytest.mark.parametrize("name,w,h", [("cornell_box", 40, 40), ("emissive", 40, 40), ("balls", 48, 27), ("rtw_final", 40, 40),
                                      ("shrek_quads", 32, 32), ("earth", 32, 18), ("synthetic", 32, 18)])
def test_host_built_scene_equals_oracle_built_scene(host, wro, images, name, w, h):
    """Two independent restatements of scene.zig + entity.zig (C++ host mirror, C oracle) must produce the same tree:
    tracing the host-flattened scene with the oracle's reference traversal gives bit-identical primary hits and an
    identical render to the oracle's own scene.  Cameras must agree to the bit as well."""
    hs = host.HostScene(name, seed=1, synthetic_prims=4096, images=images)
    os_ = wro.OracleScene(name, seed=1, n_prims=4096, images=images)
    flat = hs.flat()
    from_host = wro.OracleScene(flat=flat)
    assert from_host.n_prims == os_.n_prims
    cam_h, cam_o = hs.camera(w, h), os_.camera(w, h)
    assert bytes(cam_h) == bytes(cam_o)
    np.testing.assert_array_equal(hs.background(), os_.background())
    p = os_.params(w, h, 2, 12, seed=8)
    ids_a, t_a = os_.primary_hits(cam_o, p, 2)
    ids_b, t_b = from_host.primary_hits(cam_h, p, 2)
    np.testing.assert_array_equal(ids_a, ids_b)
    np.testing.assert_array_equal(t_a.view(np.uint64), t_b.view(np.uint64))
    fa, sa = os_.render(cam_o, p, wro.RNG_COUNTER)
    fb, sb = from_host.render(cam_h, p, wro.RNG_COUNTER)
    np.testing.assert_array_equal(fa, fb)
    assert sa.rays == sb.rays
    # same primitive table in DFS order (kinds and box centres; material numbering may differ between flatteners)
    ka, _, ca = os_.prim_table()
    kb, _, cb = from_host.prim_table()
    np.testing.assert_array_equal(ka, kb)
    np.testing.assert_array_equal(ca, cb)
    for s in (hs, os_, from_host):
        s.close()


def test_flat_scene_is_well_formed(host, images):
    hs = host.HostScene("rtw_final", seed=1, images=images)
    f = hs.flat()
    assert f.abi_version == 3 and f.root < f.n_entities and f.lights < f.n_entities
    kinds = [f.entities[i].kind for i in range(f.n_entities)]
    assert kinds.count(4) == 1 and kinds.count(5) == 1            # one Translate(RotateY(...))
    assert kinds.count(3) == 7 + 1023 + 511                        # BVH node counts
    assert f.n_spheres == 1005 and f.n_quads == 2401
    assert f.n_images == 2 and f.texel_bytes == 300 * 292 * 3 + 579 * 772 * 3
    assert hs.input_bytes() > f.texel_bytes
    # every collection child index is in range
    for i in range(f.n_entities):
        e = f.entities[i]
        if e.kind == 2:
            assert e.a + e.b <= f.n_children
            for k in range(e.b):
                assert f.children[e.a + k] < f.n_entities
    hs.close()


def test_unknown_scene_is_an_error(host):
    with pytest.raises(ValueError):
        host.HostScene("no_such_scene")


# ---- PPM writer (writer.zig) ------------------------------------------------------------------------------------------
def test_writer_known_answers(host):  # writer.zig:101-123
    def line(px):
        a = np.array(px, np.uint8)
        return host.lib.wrh_size_of_line(host._ptr(a))
    assert line([0, 0, 0]) == 6 and line([0, 255, 0]) == 8 and line([255, 255, 255]) == 12
    for digit, size in [(0, 1), (9, 1), (10, 2), (99, 2), (100, 3), (255, 3)]:
        assert host.lib.wrh_size_of_digit(digit) == size
    for x, want in [(0.0, 0), (1.0, 255), (0.25, 128), (2.0, 255), (float("nan"), 0)]:
        assert list(host.encode_color([x, x, x])) == [want] * 3


def test_writer_encode_matches_oracle_on_random_colors(host, wro):
    rng = np.random.default_rng(1)
    vals = np.concatenate([rng.uniform(-0.1, 1.5, (2000, 3)), rng.uniform(0, 1e-3, (200, 3))])
    vals[::97, 1] = np.nan
    for v in vals:
        assert list(host.encode_color(v)) == list(wro.encode_color(v))


@pytest.mark.parametrize("threads", [1, 8])
def test_ppm_file_bytes(host, wro, tmp_path, threads):
    """P3 header, one "{r} {g} {b}\\n" line per pixel in row-major order, file sized for 12 bytes per pixel and left
    NUL-padded exactly like the reference (writer.zig:20-23); truncate_to_content cuts the tail."""
    rng = np.random.default_rng(7)
    h, w = 37, 53  # 1961 pixels: two 1024-pixel chunks, the second ragged
    fb = np.zeros((h, w, 4))
    fb[..., :3] = rng.uniform(0, 1.2, (h, w, 3)) ** 2
    fb[3, 5, 0] = np.nan
    path = tmp_path / "out.ppm"
    n = host.write_ppm(path, fb, threads=threads)
    raw = path.read_bytes()
    header = f"P3\n{w} {h}\n255\n".encode()
    assert raw.startswith(header)
    assert len(raw) == h * w * 12 + len(header)
    body = raw[len(header):n]
    assert set(raw[n:]) <= {0}
    rgb = wro.encode_image(fb)
    want = "".join(f"{r} {g} {b}\n" for r, g, b in rgb.reshape(-1, 3)).encode()
    assert body == want
    n2 = host.write_ppm(tmp_path / "cut.ppm", fb, threads=threads, truncate=True)
    assert n2 == n and (tmp_path / "cut.ppm").read_bytes() == raw[:n]
    # the quantised entry point (device-side encode) writes the same file
    host.write_ppm_rgb8(tmp_path / "q.ppm", rgb, threads=threads)
    assert (tmp_path / "q.ppm").read_bytes() == raw


# ---- CLI (main.zig + argparser.zig) --------------------------------------------------------------------------------------
def run_cli(host, *args):
    return subprocess.run([str(host.CLI_PATH), *args], capture_output=True, text=True)


def test_cli_help_lists_flags_and_scenes(host):  # argparser.zig:94-113,124 ; main.zig:60-64
    for flag in ("--help", "-h", "help"):
        r = run_cli(host, flag)
        assert r.returncode == 0
        for needle in ("--image_width", "--image_height", "--image_out_path", "--thread_pool_size", "--scene",
                       "--samples_per_pixel", "--ray_bounce_max_depth", "cornell_box", "rtw_final", "emissive"):
            assert needle in r.stderr


def test_cli_argument_errors(host):  # argparser.zig:211-408 cases
    r = run_cli(host, "--image_width=4")
    assert r.returncode == 1 and "RequiredArgumentMissing" in r.stderr and "Usage:" in r.stderr
    r = run_cli(host, "--image_width=4", "--image_height=4", "--bogus=1")
    assert r.returncode == 1 and "UnrecognizedArgument" in r.stderr
    r = run_cli(host, "--image_width=4", "--image_height")
    assert r.returncode == 1 and "ArgumentMissingValue" in r.stderr
    r = run_cli(host, "--image_width=abc", "--image_height=4")
    assert r.returncode == 1 and "ParseIntFailed" in r.stderr
    r = run_cli(host, "--image_width=4", "--image_height=4", "--scene=nope")
    assert r.returncode == 1 and "ParseEnumFailed" in r.stderr


@pytest.mark.skipif("__import__('torch').cuda.is_available()", reason="only meaningful on a box without a GPU")
def test_cli_fails_loudly_without_a_gpu(host, tmp_path):
    r = run_cli(host, "-image_width=8", "---image_height=8", f"--image_out_path={tmp_path / 'x.ppm'}")  # any number of dashes
    assert r.returncode == 1
    assert "no CPU fallback" in r.stderr
    assert not (tmp_path / "x.ppm").exists()


def test_host_loader_decodes_the_reference_assets_like_zstbi(host, images):
    """Image.initFromFile (image.zig:12-17): the host library carries the reference's vendored stb_image, compiled where it
    lies, so `--asset_dir=<reference>/assets/` loads the JPEG / PNG files themselves.  Where the reference checkout is
    mounted (the build container) the texels the loader hands to wrt_upload_scene must equal the committed reference decode;
    a missing image is an error, as in the reference (no stand-in)."""
    import ctypes as C
    from pathlib import Path
    with pytest.raises(ValueError, match="ImageInitFailed"):
        host.HostScene("earth", seed=1, asset_dir="/nonexistent/")
    assets = Path("/root/reference/assets")
    if not assets.exists():
        pytest.skip("reference checkout not mounted")
    hs = host.HostScene("earth", seed=1, asset_dir=str(assets) + "/")
    f = hs.flat()
    assert f.n_images == 1 and f.images[0].width == 2048 and f.images[0].height == 1024 and f.images[0].num_components == 3
    texels = np.ctypeslib.as_array(C.cast(f.texels, C.POINTER(C.c_uint8)), shape=(f.texel_bytes,))
    np.testing.assert_array_equal(texels.reshape(1024, 2048, 3), images["earth.png"])
    hs.close()
    hs = host.HostScene("shrek_quads", seed=1, asset_dir=str(assets) + "/")
    f = hs.flat()
    texels = np.ctypeslib.as_array(C.cast(f.texels, C.POINTER(C.c_uint8)), shape=(f.texel_bytes,))
    np.testing.assert_array_equal(texels.reshape(292, 300, 3), images["wap.jpg"])
    hs.close()
